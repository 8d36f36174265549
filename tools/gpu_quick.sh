#!/bin/bash
# short GPU session: selected tests + per-layer timings + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-q}
shift
for f in "$@"; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=20 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?"; tail -3 gpurun_out/${TAG}_$n.log
done
B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers.log 2>&1
cat gpurun_out/${TAG}_layers.log | tail -60

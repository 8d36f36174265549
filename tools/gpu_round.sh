#!/bin/bash
# one GPU session: per-file GPU tests (a crashing file must not take the others down), micro-bench, bench.py
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python tools/bench_conv_tc.py > gpurun_out/${TAG}_convbench.log 2>&1
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=20 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?" >> gpurun_out/${TAG}_summary.txt
  tail -3 gpurun_out/${TAG}_$n.log >> gpurun_out/${TAG}_summary.txt
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_summary.txt
cat gpurun_out/${TAG}_summary.txt
cat gpurun_out/${TAG}_convbench.log | tail -4

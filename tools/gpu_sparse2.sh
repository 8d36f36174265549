#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
timeout 600 python -m pytest tests/test_gpu_sparse_tc.py -q -m gpu -p no:cacheprovider -x -k "tile_plan" > gpurun_out/${TAG}_plan_test.log 2>&1; echo "plan test rc=$?"; tail -3 gpurun_out/${TAG}_plan_test.log
for pm in 0 3; do
  B2S_SP_PLAN=$pm B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers_plan$pm.log 2>&1
  echo "== plan mode $pm"; grep -E "issuer|gather|epilogue" gpurun_out/${TAG}_layers_plan$pm.log | grep -E "<64,64>|<32,32>" | head -12
  grep -E "^(rulebook)" gpurun_out/${TAG}_layers_plan$pm.log | tr '\n' ';'; echo
done

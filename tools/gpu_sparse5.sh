#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
for f in tests/test_gpu_sparse_tc.py tests/test_gpu_properties.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider -x > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?"; tail -3 gpurun_out/${TAG}_$n.log
done
for pm in 0 3; do
  B2S_SP_PLAN=$pm B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers_plan$pm.log 2>&1
  echo "== plan mode $pm"; grep -E "issuer|epilogue" gpurun_out/${TAG}_layers_plan$pm.log | grep -E "<64,64>|<32,32>|<16,16>" | head -8
done
for pm in 0 3; do
  B2S_SP_PLAN=$pm timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_plan$pm.json 2> gpurun_out/${TAG}_bench_plan$pm.err
  echo "bench plan $pm rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_plan$pm.json
done
(cd tests/cuda && make mma_probe2 > /dev/null 2>&1 && timeout 120 ./mma_probe2) > gpurun_out/${TAG}_mma_probe2.txt 2>&1
grep -E "grid=148" gpurun_out/${TAG}_mma_probe2.txt

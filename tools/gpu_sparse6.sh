#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
timeout 600 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_properties.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${TAG}_test16.log 2>&1; echo "gw16 tests rc=$?"; tail -2 gpurun_out/${TAG}_test16.log
B2S_SP_GW=8 timeout 600 python -m pytest tests/test_gpu_sparse_tc.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${TAG}_test8.log 2>&1; echo "gw8 tests rc=$?"; tail -2 gpurun_out/${TAG}_test8.log
for gw in 8 16; do
for pm in 0 3; do
B2S_SP_GW=$gw B2S_SP_PLAN=$pm B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_g${gw}_p$pm.log 2>&1
echo "== gw $gw plan $pm"; grep -E "issuer" gpurun_out/${TAG}_g${gw}_p$pm.log | grep -E "<64,64>|<32,32>|<16,16>" | head -4
done
done
for gw in 8 16; do
for pm in 0 3; do
  B2S_SP_GW=$gw B2S_SP_PLAN=$pm timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_g${gw}_p$pm.json 2> gpurun_out/${TAG}_bench_g${gw}_p$pm.err
  echo "bench gw $gw plan $pm rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_g${gw}_p$pm.json
done
done

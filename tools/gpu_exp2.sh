#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-x}
for f in tests/test_gpu_sparse_tc.py tests/test_gpu_sparse.py tests/test_gpu_e2e.py tests/test_gpu_properties.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=10 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?"; tail -2 gpurun_out/${TAG}_$n.log
done
B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers.log 2>&1
grep -E "sparse_tc<64,64>|sparse_tc<32,32>|sparse_tc<8,16>" gpurun_out/${TAG}_layers.log | head -9; grep -E "^sparse_conv|^rpn|^rulebook|^nms|^vox|^to_bev|^deco" gpurun_out/${TAG}_layers.log
echo "=== GW=16"; B2S_SP_GW=16 B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers16.log 2>&1
grep -E "sparse_tc<64,64>|sparse_tc<32,32>" gpurun_out/${TAG}_layers16.log | head -6; grep -E "^sparse_conv" gpurun_out/${TAG}_layers16.log
B2S_SP_GW=16 timeout 600 python -m pytest tests/test_gpu_sparse_tc.py -q -m gpu -p no:cacheprovider --maxfail=5 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -30

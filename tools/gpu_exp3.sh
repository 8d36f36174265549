#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-x}
for f in tests/test_gpu_conv_tc.py tests/test_gpu_sparse.py tests/test_gpu_sparse_tc.py tests/test_gpu_e2e.py tests/test_gpu_properties.py tests/test_gpu_kernels.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=10 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?"; tail -2 gpurun_out/${TAG}_$n.log
done
for z in 17; do
  echo "=== flags $z"
  B2S_SP_ZSKIP=$z timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers_z$z.log 2>&1
  grep -E "sparse_tc<64,64>" gpurun_out/${TAG}_layers_z$z.log | head -3; grep -E "^sparse_conv|^rulebook" gpurun_out/${TAG}_layers_z$z.log
done
timeout 300 python tools/bench_conv_tc.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -30

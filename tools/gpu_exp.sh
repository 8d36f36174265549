#!/bin/bash
# experiment session: quick parity subset + sparse-conv timing with the diagnostic library
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-x}
for f in tests/test_gpu_conv_tc.py tests/test_gpu_sparse_tc.py tests/test_gpu_e2e.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=10 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?"; tail -2 gpurun_out/${TAG}_$n.log
done
echo "=== normal build"; B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers.log 2>&1
grep -E "sparse_tc<64,64>|sparse_tc<32,32>" gpurun_out/${TAG}_layers.log | head -8; grep -E "^sparse_conv|^rpn|^rulebook" gpurun_out/${TAG}_layers.log
if [ -f second.pytorch_b200/csrc/libb2second_diag.so ]; then
  for z in 19 21 23 25; do
    echo "=== DIAG flags $z"
    B2S_LIB=$PWD/second.pytorch_b200/csrc/libb2second_diag.so B2S_SP_ZSKIP=$z timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers_z$z.log 2>&1
    grep -E "sparse_tc<64,64>" gpurun_out/${TAG}_layers_z$z.log | head -3; grep -E "^sparse_conv(6|3|0|10) " gpurun_out/${TAG}_layers_z$z.log
  done
fi
timeout 300 python tools/bench_conv_tc.py 2>&1 | tail -2
B2S_CONV_B=32 B2S_CONV_ACC=0 timeout 300 python tools/bench_conv_tc.py 2>&1 | tail -1

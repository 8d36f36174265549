#!/bin/bash
# ncu --set full of the sparse tensor-core kernels of one steady-state step (source counters on)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-spncu}
B2S_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_sparse_conv_tc' -c 14 -o gpurun_out/${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --batch 32 > gpurun_out/${TAG}.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}*

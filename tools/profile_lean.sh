#!/bin/bash
# lean ncu evidence (keeps gpurun_out far below the 64 MiB copy-back limit):
#   <tag>_launches.csv   launch list of bench.py steps (gpu__time_duration per launch; cold cache, serialised: compare SHARES)
#   <tag>_small_raw.csv  --set full raw page of the rulebook / tile-plan kernels of ONE step (exported on the box, rep deleted)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2lean}
ARGS="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --batch 32"
B2S_PROFILE=2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python $ARGS > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
B2S_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none \
    -k regex:'k_tile_plan|k_subm_nbr|k_subm_ranked|k_conv_scatter_nbr|k_conv_mark|k_conv_emit' -o gpurun_out/${TAG}_small -f \
    python $ARGS > gpurun_out/${TAG}_small.log 2>&1
echo "small full rc=$?"
ncu -i gpurun_out/${TAG}_small.ncu-rep --page raw --csv > gpurun_out/${TAG}_small_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_small.ncu-rep
ls -la gpurun_out/${TAG}_*

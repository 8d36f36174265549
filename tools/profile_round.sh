#!/bin/bash
# ncu evidence for one steady-state step of bench.py (B200, one GPU):  bash tools/profile_round.sh <tag> [batch]
#   <tag>_launches.csv        every launch with its device time (cold cache, serialised: compare SHARES)
#   <tag>_full_raw.csv        `--set full` metrics of every kernel of ONE step (raw page, exported on the box)
#   <tag>_tc.ncu-rep          `--set full --import-source on` of the tcgen05 kernels (source page readable offline)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2prof}
B=${2:-32}
ARGS="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --batch $B"
B2S_PROFILE=2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python $ARGS > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
B2S_PROFILE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -o gpurun_out/${TAG}_full -f \
    python $ARGS > gpurun_out/${TAG}_full.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_full.ncu-rep
B2S_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'k_conv3x3_tc2|k_sparse_conv_tc|k_conv_tc' -c 24 -o gpurun_out/${TAG}_tc -f python $ARGS > gpurun_out/${TAG}_tc.log 2>&1
echo "tc rc=$?"
rm -f gpurun_out/${TAG}_tc.ncu-rep   # (65 MB: gpurun copies back at most 64 MiB in total; keep the rep only when run by hand)
ls -la gpurun_out/${TAG}_*

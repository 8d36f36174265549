#!/bin/bash
# round-end validation in ONE call: every GPU test file, the default bench line (configs block + CPU baseline), smoke(),
# and the ncu launch list of one steady-state step (gpu__time_duration per launch)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-fin}
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=20 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$? $(tail -1 gpurun_out/${TAG}_$n.log)"
done
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
B2S_PROFILE=2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --batch 32 > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"

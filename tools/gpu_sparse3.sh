#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
for fl in 17 49; do
for pm in 0 3; do
  B2S_SP_PLAN=$pm B2S_SP_ZSKIP=$fl timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers_f${fl}_plan$pm.log 2>&1
  echo "== flags $fl plan mode $pm"; grep -E "issuer|gather|epilogue" gpurun_out/${TAG}_layers_f${fl}_plan$pm.log | grep -E "<64,64>|<32,32>" | head -6
done
done
for fl in 1 33; do
  B2S_SP_ZSKIP=$fl B2S_SP_PLAN=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_f$fl.json 2> gpurun_out/${TAG}_bench_f$fl.err
  echo "bench flags $fl rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_f$fl.json
done
B2S_SP_ZSKIP=1 B2S_SP_PLAN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_p0.json 2> gpurun_out/${TAG}_bench_p0.err
python tools/show_bench.py gpurun_out/${TAG}_bench_p0.json

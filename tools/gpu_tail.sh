#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-tail}
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -p no:cacheprovider -x -k "fused_rpn_tail" > gpurun_out/${TAG}_t1.log 2>&1; rc=$?; echo "tail test rc=$rc $(tail -1 gpurun_out/${TAG}_t1.log)"
if [ $rc -ne 0 ]; then grep -E "^E |Error" gpurun_out/${TAG}_t1.log | head -12; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_conv_tc.py -q -m gpu -p no:cacheprovider --maxfail=5 > gpurun_out/${TAG}_t2.log 2>&1; echo "e2e rc=$? $(tail -1 gpurun_out/${TAG}_t2.log)"
for t in 1; do
B2S_RPN_TAIL=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_$t.json 2> gpurun_out/${TAG}_bench_$t.err
echo "bench tail=$t rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_$t.json
done

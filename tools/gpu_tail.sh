#!/bin/bash
# A/B of the fused RPN tail: y as the A operand of GEMM 2 from shared memory (smem) or from tensor memory (tmem)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-tail}
for a in tmem smem; do
B2S_RPN_TAIL_A=$a timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -p no:cacheprovider -x -k "fused_rpn_tail" > gpurun_out/${TAG}_${a}_t1.log 2>&1; rc=$?; echo "$a tail test rc=$rc $(tail -1 gpurun_out/${TAG}_${a}_t1.log)"
if [ $rc -ne 0 ]; then grep -E "^E |Error" gpurun_out/${TAG}_${a}_t1.log | head -8; continue; fi
B2S_RPN_TAIL_A=$a timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --maxfail=5 > gpurun_out/${TAG}_${a}_t2.log 2>&1; echo "$a e2e rc=$? $(tail -1 gpurun_out/${TAG}_${a}_t2.log)"
B2S_RPN_TAIL_A=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_$a.json 2> gpurun_out/${TAG}_bench_$a.err
echo "bench $a rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_$a.json
done

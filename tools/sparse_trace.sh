#!/bin/bash
# Per-role timeline of the sparse tensor-core kernel (the measurements quoted in DESIGN.md section 7).
#   build:   make -C second.pytorch_b200/csrc clean && make -C second.pytorch_b200/csrc DIAG=1      (then rebuild without DIAG!)
#   run:     bash tools/sparse_trace.sh [flags ...]            on a B200 (gpurun -- 'bash tools/sparse_trace.sh 17 81 87')
# B2S_SP_ZSKIP bits: 1 zero-slot skip (default), 16 print the issuer's cycles per K block, 32 lane-0 polling with a sleeping
# wait, and in DIAG builds (results WRONG): 2 no gather copies, 4 no weight loads, 8 no zero fills, 64 clock64 stamps of
# every role of CTA 0 per K block (printed at kernel exit), 128 plain arrive instead of cp.async.mbarrier.arrive.noinc.
# B2S_SP_PLAN: 0 no tile plan, 1 masks only, 2/3 rows grouped (SubM / all rulebooks), 4 SubM rulebooks only (default).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for fl in "${@:-17}"; do
  for pm in 0 4; do
    B2S_SP_PLAN=$pm B2S_SP_ZSKIP=$fl timeout 300 python tools/layer_times.py 32 > gpurun_out/sparse_trace_p${pm}_f$fl.log 2>&1
    echo "== B2S_SP_PLAN=$pm B2S_SP_ZSKIP=$fl"
    grep -E "issuer" gpurun_out/sparse_trace_p${pm}_f$fl.log | head -14
    grep -A 24 "trace<64,64>" gpurun_out/sparse_trace_p${pm}_f$fl.log | sed -n 1,26p
  done
done

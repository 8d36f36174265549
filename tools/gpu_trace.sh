#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B2S_SP_PLAN=0 B2S_SP_ZSKIP=113 timeout 300 python tools/layer_times.py 32 > gpurun_out/diag6_trace.log 2>&1
grep -A 60 "trace<64,64>" gpurun_out/diag6_trace.log | sed -n 12,36p
B2S_SP_PLAN=0 B2S_SP_ZSKIP=119 timeout 300 python tools/layer_times.py 32 > gpurun_out/diag6_trace_nomem.log 2>&1
grep -A 60 "trace<64,64>" gpurun_out/diag6_trace_nomem.log | sed -n 12,28p

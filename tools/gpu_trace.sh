#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for fl in 21 19 23; do
B2S_SP_OWN=4 B2S_SP_PLAN=0 B2S_SP_ZSKIP=$fl timeout 300 python tools/layer_times.py 32 > gpurun_out/trace_o4_f$fl.log 2>&1
echo "== own 4 flags $fl"; grep -E "issuer" gpurun_out/trace_o4_f$fl.log | grep -E "<64,64>|<32,32>|<16,16>" | head -4
done
B2S_SP_OWN=4 B2S_SP_PLAN=0 B2S_SP_ZSKIP=85 timeout 300 python tools/layer_times.py 32 > gpurun_out/trace_o4_f85.log 2>&1
grep -A 40 "trace<64,64>" gpurun_out/trace_o4_f85.log | sed -n 12,22p

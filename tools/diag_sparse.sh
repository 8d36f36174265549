#!/bin/bash
# diagnostic (make DIAG=1 build): issuer / gather-warp cycle split of the sparse tensor-core kernel under B2S_SP_ZSKIP variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for fl in "$@"; do
  B2S_SP_ZSKIP=$fl timeout 300 python tools/layer_times.py 32 > gpurun_out/diag_sp_$fl.log 2>&1
  echo "== flags $fl"; grep -E "sparse_tc<64,64>\] issuer|sparse_tc<32,32>\] issuer|sparse_tc<64,64>\] gather" gpurun_out/diag_sp_$fl.log | head -6
  grep -E "^sparse_conv(7|4) " gpurun_out/diag_sp_$fl.log
done

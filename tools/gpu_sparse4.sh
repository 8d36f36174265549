#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
for fl in 17 21 19 23 25; do
  B2S_SP_PLAN=0 B2S_SP_ZSKIP=$fl timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_diag_f${fl}.log 2>&1
  echo "== flags $fl"; grep -E "issuer" gpurun_out/${TAG}_diag_f${fl}.log | grep -E "<64,64>|<32,32>|<16,16>" | head -4
done
python - <<'PY'
import sys, os, torch, ctypes
sys.path.insert(0, "second.pytorch_b200")
from b2second import loader
sp = loader.product_spconv(); L = sp._lib; lib = L.load()
for n in (100000, 500000):
    nbr = (torch.randint(0, 100, (n, 27), device="cuda") < 30).int() * torch.arange(n, device="cuda", dtype=torch.int32)[:, None] - 1
    nbr = nbr.int().contiguous()
    nd = torch.tensor([n], dtype=torch.int32, device="cuda")
    perm = torch.zeros(n, dtype=torch.int32, device="cuda"); tm = torch.zeros((n + 127) // 128, dtype=torch.int32, device="cuda")
    for sort in (0, 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(3):
            L.check(lib.b2s_sparse_tile_plan(L.ptr(nbr), 27, L.i3([3, 3, 3]), L.ptr(nd), n, sort, L.ptr(perm), L.ptr(tm), L.stream()), "plan")
        torch.cuda.synchronize(); ev[0].record()
        for it in range(20):
            L.check(lib.b2s_sparse_tile_plan(L.ptr(nbr), 27, L.i3([3, 3, 3]), L.ptr(nd), n, sort, L.ptr(perm), L.ptr(tm), L.stream()), "plan")
        ev[1].record(); torch.cuda.synchronize()
        print("plan n=%d sort=%d: %.1f us per call" % (n, sort, ev[0].elapsed_time(ev[1]) * 1000 / 20))
PY

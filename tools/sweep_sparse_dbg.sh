# diagnostic sweep of k_sparse_conv_tc (B2S_SP_ZSKIP bits: 1 zero-slot skip, 2 no gather copies, 4 no weight loads,
# 8 no neighbour-table staging; bits >= 2 give wrong results) -- prints the sparse-conv stage time per setting
for z in 1 3 5 7 9 15; do
  B2S_SP_ZSKIP=$z timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 16 > gpurun_out/z.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/z.log").read().strip().splitlines()[-1])
    print("zskip=$z sparse_conv %.3f ms  step %.3f" % (d["stage_ms_eager"]["sparse_conv"], d["ms_per_step"]))
except Exception as e:
    print("zskip=$z failed:", open("gpurun_out/z.log").read().strip().splitlines()[-1][:200])
PY
done

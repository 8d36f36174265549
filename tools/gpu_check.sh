#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-x}
for f in tests/test_gpu_kernels.py tests/test_gpu_conv_tc.py tests/test_gpu_e2e.py tests/test_gpu_properties.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -p no:cacheprovider --maxfail=10 > gpurun_out/${TAG}_$n.log 2>&1
  echo "$n rc=$?"; tail -2 gpurun_out/${TAG}_$n.log
done
B2S_RPN_BG_FILL=separate timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k background 2>&1 | tail -1
B2S_RPN_BG_FILL=separate timeout 300 python tools/layer_times.py 32 2>&1 | grep -E "^rpn "
timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers.log 2>&1
grep -E "^rpn|^to_bev|^nms|^deco" gpurun_out/${TAG}_layers.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -30
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]; print({k:r[k] for k in ("achieved","frac","ms_per_launch","frac_of_pipe_issued","tiles_computed_frac_per_layer") if k in r})
for c in d["configs"]: print(c.get("config"), c.get("frames_per_gpu_per_step"), round(c.get("ms_per_step",0),3), round(c.get("clouds_per_s",0)), round(c.get("e2e_clouds_per_s",0)), c.get("error"))
PY

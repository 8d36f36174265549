#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
timeout 600 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_properties.py tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
B2S_SP_ZSKIP=17 timeout 300 python tools/layer_times.py 32 > gpurun_out/${TAG}_layers.log 2>&1
grep -E "issuer" gpurun_out/${TAG}_layers.log | head -14
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench.json

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-sp}
timeout 900 python -m pytest tests/test_gpu_sparse_tc.py tests/test_gpu_sparse.py tests/test_gpu_properties.py tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
for pm in 0 2 4; do
  B2S_SP_PLAN=$pm timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_bench_p$pm.json 2> gpurun_out/${TAG}_bench_p$pm.err
  echo "bench plan $pm rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_p$pm.json
done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/last_bench.json
timeout 60 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -p no:cacheprovider -x -k "fused_rpn_tail" 2>&1 | tail -1

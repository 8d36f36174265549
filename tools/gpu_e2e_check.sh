#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-e2e}
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_properties.py -q -m gpu -p no:cacheprovider --maxfail=10 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/'+__import__('sys').argv[1] if False else 'gpurun_out/TAG_bench.json'.replace('TAG','%s')).read().strip().splitlines()[-1]) if False else None
PY
python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
for c in d.get('configs',[]): print(c['config'], round(c['clouds_per_s']), round(c['e2e_clouds_per_s']))
"

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1; echo "pytest gpu rc=$?"
tail -5 gpurun_out/full_pytest.log
timeout 900 python bench.py > gpurun_out/full_bench_n1.json 2> gpurun_out/full_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/full_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/full_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','unit','ms_per_step','gpu_launches') if k in d})
print('e2e', d.get('e2e')); print('roofline', d.get('roofline')); print('cpu_baseline', d.get('cpu_baseline'))
enc=d.get('encode') or {}
for k,v in enc.items():
    print('  enc', k, v)
for k in ('train_first','decode','checks'):
    if k in d: print(k, d[k])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --selftest > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/n2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','unit','ms_per_step','gpu_launches','scaling','n_gpus') if k in d})
print('e2e', d.get('e2e'))
for k in ('replicas','selftest','train_first','decode','checks','sharded'):
    if k in d: print(k, d[k])
enc=d.get('encode') or {}
for k in ('value','ms_per_step','ms_per_gib','scaling','e2e','roundtrip_ok','warm_caches'):
    print('  enc', k, enc.get(k))
PY

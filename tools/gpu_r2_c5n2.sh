#!/bin/bash
# 2 GPUs: config 5 leg (text sharded over the ranks) at reduced size + the sharded parity tests
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 1 --corpus-mib 256 --encode-gib 1 --skip-first --skip-cpu-baseline --config5 --config5-gib 2 --config5-vocab 20000 > gpurun_out/c5n2.json 2> gpurun_out/c5n2.err; echo "bench c5 n2 rc=$?"
tail -c 1200 gpurun_out/c5n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c5n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('metric','value','unit','ms_per_step','n_gpus') if k in d})
    print('config5', json.dumps(d.get('config5'))[:1500])
except Exception as e: print('no json', e)
PY
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/c5n2_pytest.log 2>&1; echo "pytest sharded rc=$?"
tail -3 gpurun_out/c5n2_pytest.log

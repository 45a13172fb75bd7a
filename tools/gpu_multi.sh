#!/bin/bash
# N GPUs: the sharded trainer with the resident exchange (train leg only, 1 step) + config 5 at reduced size
N=${1:-8}
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 1 --warmup 1 --skip-encode --skip-first --skip-cpu-baseline --config5 --config5-gib 2 --config5-vocab 20000 > gpurun_out/multi_n${N}.json 2> gpurun_out/multi_n${N}.err; echo "bench n$N rc=$?"
tail -c 800 gpurun_out/multi_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_n${N}.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('metric','value','unit','ms_per_step','n_gpus','gpu_launches') if k in d})
    t=d['train']; print({k:t.get(k) for k in ('what','merges','same_merges_on_every_rank')}, t.get('oracle_check'), t.get('replicas'))
    print('config5', json.dumps(d.get('config5'))[:900])
except Exception as e: print('no json', e)
PY

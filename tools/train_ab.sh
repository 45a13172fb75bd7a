#!/bin/bash
# A/B of train-loop variants on the GPU box: tools/train_ab.sh "ENV1=.." "ENV2=.." (each run: bench train leg only, merge list checked against the oracle)
mkdir -p gpurun_out
for v in "$@"; do
  env $v timeout 200 python bench.py --steps 3 --warmup 3 --skip-encode --skip-cpu-baseline --skip-first 2>gpurun_out/train_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); rc=d['train']['stats']['resident_cycles']
print('$v', round(d['ms_per_step'],1), 'ms', d['train']['oracle_check']['merges_equal'], {k: round(rc[k]/rc['steps']) for k in ('select','hits','mutate_alloc','seg_fill','fin')}, 'e2e', d['e2e']['wall_s_steps'])" 2>&1 | tail -2
done

#!/bin/bash
# A/B of train-loop variants on the GPU box: tools/train_ab.sh "ENV1=.." "ENV2=.." (each run: bench train leg only)
for v in "$@"; do
  env $v timeout 400 python bench.py --steps 3 --warmup 3 --skip-encode --skip-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); rc=d['train']['stats']['resident_cycles']
print('$v', round(d['ms_per_step'],1), 'ms', d['train']['oracle_check']['merges_equal'], {k: round(rc[k]/rc['steps']) for k in ('select','hits','mutate_alloc','seg_fill','fin')}, 'e2e', d['e2e']['wall_s_steps'])"
done

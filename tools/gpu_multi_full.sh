#!/bin/bash
# N GPUs of one box: the default bench through torchrun, with the sharded trainer diffed against the oracle on the goldens first
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --selftest > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -c 400 gpurun_out/bench_n$N.err

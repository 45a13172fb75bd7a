#!/bin/bash
# 2 GPUs: sharded trainer (resident exchange through peer-mapped memory) -- parity test, then timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/s2_env.txt 2>&1
nvidia-smi topo -m >> gpurun_out/s2_env.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -s > gpurun_out/s2_pytest.log 2>&1; echo "pytest sharded rc=$?"
tail -15 gpurun_out/s2_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py 64 4096 > gpurun_out/s2_check64.json 2> gpurun_out/s2_check64.err; echo "check64 rc=$?"
tail -c 1500 gpurun_out/s2_check64.json; tail -5 gpurun_out/s2_check64.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/sharded_check.py 1024 32768 > gpurun_out/s2_check1g.json 2> gpurun_out/s2_check1g.err; echo "check1g rc=$?"
tail -c 1500 gpurun_out/s2_check1g.json; tail -5 gpurun_out/s2_check1g.err | cut -c1-300

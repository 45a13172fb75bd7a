#!/bin/bash
# 2 GPUs: sharded trainer against the oracle (both modes, 64 MiB), then the 1 GiB timing with and without the candidate mirror
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/sh_pytest.log 2>&1; echo "pytest sharded rc=$?"; tail -2 gpurun_out/sh_pytest.log
for v in "CHECK_SKIP_STEPWISE=1" "CHECK_SKIP_STEPWISE=1 MBPE_NO_SHARD_MIRROR=1"; do
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/sharded_check.py 1024 32768 > "gpurun_out/sh_check1g_${v// /_}.json" 2> gpurun_out/sh_check1g.err; echo "check1g [$v] rc=$?"
  tail -c 700 "gpurun_out/sh_check1g_${v// /_}.json"; echo
done

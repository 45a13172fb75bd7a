#!/bin/bash
# One gpurun call: environment facts, smoke, GPU parity tests, short bench. Every step under its own timeout.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  echo "nproc=$(nproc)"; free -g | head -2
} > gpurun_out/env.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/env.txt
timeout ${TEST_TIMEOUT:-1200} python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/env.txt
tail -5 gpurun_out/pytest_gpu.log
timeout ${BENCH_TIMEOUT:-900} python bench.py ${BENCH_ARGS:---steps 2 --warmup 1 --corpus-mib 256 --encode-mib 256} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/env.txt
tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err

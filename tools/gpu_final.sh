#!/bin/bash
# end of round: the default bench (N = 1), the reference arm, and the whole GPU test suite
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/final_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm rc=$?"
tail -c 600 gpurun_out/final_bench_ref.json
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -3 gpurun_out/final_pytest_gpu.log

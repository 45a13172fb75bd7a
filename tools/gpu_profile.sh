#!/bin/bash
# One gpurun call: full-size bench (plain), the ncu launch list of the same command, and ncu --set full captures
# of the two dominant kernels on a smaller run. A run under ncu is never a bench value.
mkdir -p gpurun_out
FULL="python bench.py --steps 2 --warmup 1"
SMALL="python bench.py --steps 1 --warmup 0 --corpus-mib 256 --encode-mib 256 --skip-cpu-baseline"
timeout 1500 $FULL > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; rc=$?; echo "bench_full rc=$rc"
tail -c 600 gpurun_out/bench_full.err
if [ $rc -eq 0 ]; then
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv $FULL > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
timeout 600 $SMALL > gpurun_out/bench_small.log 2> gpurun_out/bench_small.err; rc=$?; echo "bench_small rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode_tiles -c 1 -o gpurun_out/prof_encode $SMALL > gpurun_out/ncu_encode.log 2>&1; echo "ncu encode rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 20 -c 2 -o gpurun_out/prof_persistent $SMALL > gpurun_out/ncu_persistent.log 2>&1; echo "ncu persistent rc=$?"
fi
ls -la gpurun_out

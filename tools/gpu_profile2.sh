#!/bin/bash
# final-version evidence for profiles/: launch list (first 2500 launches) + --set full of the two dominant kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --no-check"
timeout 400 $CMD > gpurun_out/p2_plain.log 2> gpurun_out/p2_plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/p2_ncu_launches.log 2>&1; echo "launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 40 -c 1 -o gpurun_out/prof_persistent_final $CMD > gpurun_out/p2_ncu_persistent.log 2>&1; echo "persistent rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_encode_tiles -s 60 -c 1 -o gpurun_out/prof_encode_final $CMD > gpurun_out/p2_ncu_encode.log 2>&1; echo "encode rc=$?"
fi
ls -la gpurun_out | tail -8

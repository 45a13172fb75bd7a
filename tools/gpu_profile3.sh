#!/bin/bash
# evidence for the pre-tokeniser / dedup kernels: launch list + --set full of the three kernels (run under gpurun)
mkdir -p gpurun_out
CMD="python tools/pretok_bench.py 512"
timeout 300 $CMD > gpurun_out/p3_plain.log 2> gpurun_out/p3_plain.err; rc=$?; echo "plain rc=$rc"; cat gpurun_out/p3_plain.log
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pretok.csv $CMD > gpurun_out/p3_ncu_launches.log 2>&1; echo "launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pretok_mark|k_bits_compact" -s 2 -c 2 -o gpurun_out/prof_pretok $CMD > gpurun_out/p3_ncu_pretok.log 2>&1; echo "pretok rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dedup_insert -c 1 -o gpurun_out/prof_dedup $CMD > gpurun_out/p3_ncu_dedup.log 2>&1; echo "dedup rc=$?"
fi
ls -la gpurun_out | tail -6

#!/bin/bash
# end-of-round evidence: launch list of the default bench + --set full of the dominant kernels (run under gpurun)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --no-check"
timeout 400 $CMD > gpurun_out/p4_plain.log 2> gpurun_out/p4_plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_end.csv $CMD > gpurun_out/p4_ncu_launches.log 2>&1; echo "launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 40 -c 1 -o gpurun_out/prof_persistent_end $CMD > gpurun_out/p4_ncu_persistent.log 2>&1; echo "persistent rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_encode_tiles|k_decode_tiles" -s 12 -c 2 -o gpurun_out/prof_encdec_end $CMD > gpurun_out/p4_ncu_encdec.log 2>&1; echo "encode/decode rc=$?"
fi
ls -la gpurun_out | tail -6

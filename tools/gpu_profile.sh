#!/bin/bash
# One gpurun call (one GPU): evidence for profiles/. A number printed by a run under ncu is never a bench value.
#   1. the bench command, plain (must exit 0), then its ncu launch list (per-launch device time, serialised and cold)
#   2. ncu --set full of one launch of each dominant kernel: k_persistent (train), k_encode_tiles (warm), k_decode_lean
# Outputs in gpurun_out/: prof_*.csv / prof_*.ncu-rep / prof_*.log. tools/profile_summary.py folds them into profiles/.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --encode-gib 2 --skip-first --skip-cpu-baseline --no-check"
timeout 600 $CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/prof_launches.csv $CMD > gpurun_out/prof_launches.log 2>&1; echo "launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 300 -c 1 -o gpurun_out/prof_persistent $CMD > gpurun_out/prof_persistent.log 2>&1; echo "persistent rc=$?"
fi
timeout 300 python tools/enc_ab.py 512 0 > gpurun_out/prof_enc_plain.log 2>&1; rc=$?; echo "enc plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode_tiles --launch-skip 11 --launch-count 1 -o gpurun_out/prof_encode python tools/enc_ab.py 512 0 > gpurun_out/prof_encode.log 2>&1; echo "encode rc=$?"
fi
timeout 300 python tools/dec_ab.py 1024 > gpurun_out/prof_dec_plain.log 2>&1; rc=$?; echo "dec plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode --launch-skip 3 --launch-count 1 -o gpurun_out/prof_decode python tools/dec_ab.py 1024 > gpurun_out/prof_decode.log 2>&1; echo "decode rc=$?"
fi
grep -E "^cfg|best" gpurun_out/prof_enc_plain.log gpurun_out/prof_dec_plain.log
ls -la gpurun_out | grep prof_

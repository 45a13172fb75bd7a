#!/bin/bash
# decode: ncu --set full of one launch (MBPE_DEC_CFG=$1, default configuration if empty) -> gpurun_out/prof_decode.ncu-rep
mkdir -p gpurun_out
[ -n "$1" ] && export MBPE_DEC_CFG=$1
timeout 120 python tools/dec_ab.py 1024 > gpurun_out/prof_dec_plain.log 2>&1; echo "plain rc=$?"; grep best gpurun_out/prof_dec_plain.log | head -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_decode --launch-skip 3 --launch-count 1 -o gpurun_out/prof_decode python tools/dec_ab.py 1024 > gpurun_out/prof_decode.log 2>&1; echo "ncu rc=$?"

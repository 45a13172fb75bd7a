#!/bin/bash
# decode: ncu --set full of one launch of the lane-per-token variant
mkdir -p gpurun_out
MBPE_DEC_CFG=${1:-6} timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_decode --launch-skip 3 --launch-count 1 -o gpurun_out/prof_decode_lanes python tools/dec_ab.py 1024 > gpurun_out/d2_ncu.log 2>&1; echo "ncu rc=$?"

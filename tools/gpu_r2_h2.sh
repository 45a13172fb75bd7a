#!/bin/bash
# k_encode_hot: phase statistics (MBPE_DEBUG) and one ncu --set full capture of a warm launch
mkdir -p gpurun_out
CFG=${1:-8}
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 $CFG > gpurun_out/h2_stats.log 2>&1; echo "stats rc=$?"
grep -E "^cfg|k_encode_hot" gpurun_out/h2_stats.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode_hot --launch-skip 11 --launch-count 1 -o gpurun_out/prof_hot python tools/enc_ab.py 512 $CFG > gpurun_out/h2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_hot.ncu-rep

#!/bin/bash
# k_encode_hot: rebuild with per-step statistics, print the breakdown, capture one warm launch with ncu --set full
mkdir -p gpurun_out
CFG=${1:-8}
MBPE_DEFS=-DMBPE_HOT_STATS timeout 400 python minbpe-cc_b200/build.py --force > gpurun_out/h4_build.log 2>&1; echo "stats build rc=$?"
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 $CFG > gpurun_out/h4_stats.log 2>&1; echo "stats rc=$?"
grep -E "^cfg|k_encode_hot" gpurun_out/h4_stats.log | tail -3 | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode_hot --launch-skip 11 --launch-count 1 -o gpurun_out/prof_hot python tools/enc_ab.py 512 $CFG > gpurun_out/h4_ncu.log 2>&1; echo "ncu rc=$?"

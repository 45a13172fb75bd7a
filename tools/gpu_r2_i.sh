#!/bin/bash
mkdir -p gpurun_out
MBPE_ENC_CFG=0 timeout 300 python tools/enc_ab.py 256 0 > gpurun_out/i_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode_tiles -s 14 -c 2 -o gpurun_out/prof_enc_r2i python tools/enc_ab.py 256 0 > gpurun_out/i_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/i_ncu.log

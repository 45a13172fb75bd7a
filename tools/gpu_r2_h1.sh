#!/bin/bash
# k_encode_hot: parity (whole encode test file under two hot configurations), then A/B against k_encode_tiles
mkdir -p gpurun_out
for c in 9 8; do
MBPE_ENC_CFG=$c timeout 900 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/h1_pytest_$c.log 2>&1; echo "pytest cfg $c rc=$?"
tail -5 gpurun_out/h1_pytest_$c.log | cut -c1-400
done
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 8 9 10 11 > gpurun_out/h1_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg|cache misses" gpurun_out/h1_enc_ab.log | cut -c1-250

#!/bin/bash
# k_encode_hot: parity under one configuration, A/B timing, then a rebuild with per-step statistics for the breakdown
mkdir -p gpurun_out
MBPE_ENC_CFG=9 timeout 900 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/h3_pytest.log 2>&1; echo "pytest cfg 9 rc=$?"
tail -3 gpurun_out/h3_pytest.log | cut -c1-400
timeout 600 python tools/enc_ab.py 512 0 8 9 10 11 > gpurun_out/h3_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/h3_enc_ab.log | cut -c1-250
MBPE_DEFS=-DMBPE_HOT_STATS timeout 400 python minbpe-cc_b200/build.py > gpurun_out/h3_build.log 2>&1; echo "stats build rc=$?"
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 ${1:-9} > gpurun_out/h3_stats.log 2>&1; echo "stats rc=$?"
grep -E "^cfg|k_encode_hot" gpurun_out/h3_stats.log | tail -3 | cut -c1-600

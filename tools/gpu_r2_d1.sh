#!/bin/bash
# decode: lane-per-token assembly (MBPE_DEC_CFG 5..7) -- A/B against the thread-per-8-ids kernel (equality with the text is checked), then parity tests
mkdir -p gpurun_out
AB_ENV="MBPE_DEC_CFG=0;MBPE_DEC_CFG=5;MBPE_DEC_CFG=6;MBPE_DEC_CFG=7" timeout 150 python tools/dec_ab.py 1024 > gpurun_out/d1_dec_ab.log 2>&1; echo "dec ab rc=$?"
grep best gpurun_out/d1_dec_ab.log
MBPE_DEC_CFG=5 timeout 200 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/d1_pytest.log 2>&1; echo "pytest dec cfg 5 rc=$?"
tail -3 gpurun_out/d1_pytest.log | cut -c1-300

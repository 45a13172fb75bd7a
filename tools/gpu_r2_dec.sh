#!/bin/bash
mkdir -p gpurun_out
for c in 0 3; do MBPE_DEC_CFG=$c timeout 300 python -m pytest tests/test_gpu_encode.py -x -q -m gpu -k "decode or roundtrip or round_trip" > gpurun_out/dec_pytest_$c.log 2>&1; echo "pytest decode cfg $c rc=$?"; tail -1 gpurun_out/dec_pytest_$c.log; done
AB_ENV="MBPE_DEC_CFG=0;MBPE_DEC_CFG=1;MBPE_DEC_CFG=2;MBPE_DEC_CFG=3;MBPE_DEC_CFG=4;MBPE_DEC_CFG=5" timeout 600 python tools/dec_ab.py 1024 > gpurun_out/dec_ab.log 2>&1; echo "dec ab rc=$?"
grep -E "best" gpurun_out/dec_ab.log | head -6

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/y8_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -3 gpurun_out/y8_pytest_encode.log
timeout 600 python tools/enc_ab.py 512 > gpurun_out/y8_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/y8_enc_ab.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 1 2 > gpurun_out/y8_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|thread 0|^cfg" gpurun_out/y8_enc_prof.log | grep -B2 "^cfg" | cut -c1-420

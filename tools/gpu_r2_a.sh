#!/bin/bash
# round 2, call A: new encode kernel -- parity tests first, then the configuration A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/env.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/a_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -15 gpurun_out/a_pytest_encode.log
timeout 900 python tools/enc_ab.py 512 > gpurun_out/a_enc_ab.log 2>&1; echo "enc_ab rc=$?"
cat gpurun_out/a_enc_ab.log | tail -20
AB_ENV="MBPE_ENC_NO_BULK=1;MBPE_ENC_NO_BULK=-,MBPE_ENC_ABLATE=1;MBPE_ENC_ABLATE=2;MBPE_ENC_ABLATE=4;MBPE_ENC_ABLATE=7;MBPE_ENC_ABLATE=-,MBPE_NO_L2_PERSIST=1" timeout 900 python tools/enc_ab.py 512 0 1 > gpurun_out/a_enc_ablate.log 2>&1; echo "ablate rc=$?"
tail -20 gpurun_out/a_enc_ablate.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -8 gpurun_out/a_pytest_all.log

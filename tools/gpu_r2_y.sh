#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/y_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -3 gpurun_out/y_pytest_encode.log
timeout 300 python tools/enc_dbg.py 8 0 4 > gpurun_out/y_dbg.log 2>&1; echo "dbg rc=$?"; grep -v "same=True" gpurun_out/y_dbg.log | head -5 | cut -c1-300
timeout 600 python tools/enc_ab.py 512 > gpurun_out/y_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/y_enc_ab.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 1 > gpurun_out/y_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|^cfg" gpurun_out/y_enc_prof.log | grep -B1 "^cfg" | cut -c1-420
AB_ENV="MBPE_ENC_ABLATE=1;MBPE_ENC_ABLATE=2;MBPE_ENC_ABLATE=4;MBPE_ENC_ABLATE=7;MBPE_ENC_ABLATE=-,MBPE_ENC_NO_BULK=1;MBPE_ENC_NO_BULK=-,MBPE_NO_L2_PERSIST=1" timeout 600 python tools/enc_ab.py 512 0 > gpurun_out/y_enc_ablate.log 2>&1; echo "ablate rc=$?"
grep -E "^cfg" gpurun_out/y_enc_ablate.log

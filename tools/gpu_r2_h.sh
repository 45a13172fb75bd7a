#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/h_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -3 gpurun_out/h_pytest_encode.log
timeout 300 python tools/enc_dbg.py 8 3 6 > gpurun_out/h_dbg.log 2>&1; echo "dbg rc=$?"; grep -v "same=True" gpurun_out/h_dbg.log | head -5 | cut -c1-300
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 1 > gpurun_out/h_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|^cfg" gpurun_out/h_enc_prof.log | awk 'NR%7==6 || /^cfg/' | cut -c1-420
timeout 600 python tools/enc_ab.py 512 2 3 4 5 6 7 > gpurun_out/h_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/h_enc_ab.log
AB_ENV="MBPE_ENC_ABLATE=8;MBPE_ENC_ABLATE=4;MBPE_ENC_ABLATE=7" timeout 600 python tools/enc_ab.py 512 0 > gpurun_out/h_enc_ablate.log 2>&1; echo "ablate rc=$?"
grep -E "^cfg" gpurun_out/h_enc_ablate.log

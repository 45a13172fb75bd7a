#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/y2_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -3 gpurun_out/y2_pytest_encode.log
timeout 600 python tools/enc_ab.py 512 > gpurun_out/y2_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/y2_enc_ab.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 > gpurun_out/y2_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|^cfg" gpurun_out/y2_enc_prof.log | grep -B1 "^cfg" | cut -c1-420
AB_ENV="MBPE_ENC_ABLATE=1;MBPE_ENC_ABLATE=2;MBPE_ENC_ABLATE=4;MBPE_ENC_ABLATE=7" timeout 600 python tools/enc_ab.py 512 0 > gpurun_out/y2_enc_ablate.log 2>&1; echo "ablate rc=$?"
grep -E "^cfg" gpurun_out/y2_enc_ablate.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_encode_tiles --launch-skip 11 --launch-count 1 -o gpurun_out/y2_encode_tiles python tools/enc_ab.py 512 0 > gpurun_out/y2_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/y2_ncu.log

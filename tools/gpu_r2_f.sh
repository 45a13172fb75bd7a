#!/bin/bash
# round 2, call F: tile kernel with the parallel open-chunk phase -- parity, phase counters, config sweep, ablation
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/f_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -3 gpurun_out/f_pytest_encode.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 1 2 4 > gpurun_out/f_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|^cfg" gpurun_out/f_enc_prof.log | cat | tail -12
timeout 600 python tools/enc_ab.py 512 3 5 6 7 > gpurun_out/f_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/f_enc_ab.log
AB_ENV="MBPE_ENC_ABLATE=8;MBPE_ENC_ABLATE=4;MBPE_ENC_ABLATE=7;MBPE_ENC_ABLATE=-,MBPE_ENC_NO_BULK=1" timeout 600 python tools/enc_ab.py 512 0 > gpurun_out/f_enc_ablate.log 2>&1; echo "ablate rc=$?"
grep -E "^cfg" gpurun_out/f_enc_ablate.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/y9_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/y9_pytest.log
timeout 600 python tools/enc_ab.py 512 0 1 3 6 4 7 > gpurun_out/y9_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/y9_enc_ab.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 1 > gpurun_out/y9_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|thread 0|^cfg" gpurun_out/y9_enc_prof.log | grep -B2 "^cfg" | cut -c1-420
timeout 300 python tools/dec_ab.py 1024 > gpurun_out/y9_dec.log 2>&1; grep best gpurun_out/y9_dec.log | head -1
timeout 600 python tools/ref16_gpu.py > gpurun_out/y9_ref16.json 2> gpurun_out/y9_ref16.err; echo "ref16 rc=$?"; tail -c 600 gpurun_out/y9_ref16.json

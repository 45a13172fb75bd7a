#!/bin/bash
# round 2, call B: encode parity again (bounded), phase cycle counters, finer ablation, decode timing, one ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu -s > gpurun_out/b_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -6 gpurun_out/b_pytest_encode.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 2 6 > gpurun_out/b_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|^cfg" gpurun_out/b_enc_prof.log | tail -40
AB_ENV="MBPE_ENC_ABLATE=8;MBPE_ENC_ABLATE=2;MBPE_ENC_ABLATE=10;MBPE_ENC_ABLATE=-,MBPE_ENCODE_CACHE=23;MBPE_ENCODE_CACHE=21;MBPE_ENCODE_CACHE=-" timeout 600 python tools/enc_ab.py 512 6 > gpurun_out/b_enc_ablate.log 2>&1; echo "ablate rc=$?"
grep -E "^cfg" gpurun_out/b_enc_ablate.log
timeout 600 python tools/dec_ab.py 512 > gpurun_out/b_dec.log 2>&1; echo "dec rc=$?"
tail -3 gpurun_out/b_dec.log
MBPE_ENC_CFG=6 timeout 300 python tools/enc_ab.py 256 6 > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode_tiles -s 7 -c 1 -o gpurun_out/prof_enc_r2b python tools/enc_ab.py 256 6 > gpurun_out/b_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5

#!/bin/bash
# round 2, call C: reworked tile kernel -- parity, phase counters, config sweep; then the whole GPU suite and the new bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/c_pytest_encode.log 2>&1; echo "pytest encode rc=$?"
tail -4 gpurun_out/c_pytest_encode.log
MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 2 6 > gpurun_out/c_enc_prof.log 2>&1; echo "enc prof rc=$?"
grep -E "cycles per tile|^cfg" gpurun_out/c_enc_prof.log | awk 'NR%6==0 || /^cfg/' | tail -12
timeout 600 python tools/enc_ab.py 512 1 3 4 7 > gpurun_out/c_enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/c_enc_ab.log
AB_ENV="MBPE_ENC_ABLATE=8;MBPE_ENC_ABLATE=1;MBPE_ENC_ABLATE=4;MBPE_ENC_ABLATE=7;MBPE_ENC_ABLATE=-,MBPE_ENC_NO_BULK=1" timeout 600 python tools/enc_ab.py 512 6 > gpurun_out/c_enc_ablate.log 2>&1; echo "ablate rc=$?"
grep -E "^cfg" gpurun_out/c_enc_ablate.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -6 gpurun_out/c_pytest_all.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c_bench.err; head -c 3000 gpurun_out/c_bench.json

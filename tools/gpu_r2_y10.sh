#!/bin/bash
mkdir -p gpurun_out
AB_ENV="MBPE_ENCODE_CACHE=22;MBPE_ENCODE_CACHE=21;MBPE_ENCODE_CACHE=20;MBPE_ENCODE_CACHE=23;MBPE_ENCODE_CACHE=24" MBPE_DEBUG=1 timeout 600 python tools/enc_ab.py 512 0 > gpurun_out/y10_cache.log 2>&1; echo "rc=$?"
grep -E "^cfg|cache entries" gpurun_out/y10_cache.log | awk '/^cfg/ {print last; print} {last=$0}' | cut -c1-300

#!/bin/bash
# decode variants: A/B (output compared with the text). usage: tools/gpu_dec_ab.sh "ENV1;ENV2;..." (AB_ENV syntax of tools/dec_ab.py)
mkdir -p gpurun_out
AB_ENV="$1" timeout 200 python tools/dec_ab.py 1024 > gpurun_out/dec_ab.log 2>&1; echo "dec ab rc=$?"
grep best gpurun_out/dec_ab.log

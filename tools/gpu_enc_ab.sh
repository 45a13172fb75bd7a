#!/bin/bash
# encode variants: tools/gpu_enc_ab.sh "ENV1;ENV2;..." [cfg ...] (AB_ENV syntax of tools/enc_ab.py; ids compared between the variants)
mkdir -p gpurun_out
ENVS="$1"; shift
AB_ENV="$ENVS" timeout 280 python tools/enc_ab.py 512 ${@:-0} > gpurun_out/enc_ab.log 2>&1; echo "enc ab rc=$?"
grep -E "^cfg" gpurun_out/enc_ab.log | cut -c1-220

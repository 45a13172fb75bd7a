#!/bin/bash
# decode variants: A/B (output compared with the text), then the encode/decode test file under the first variant given
mkdir -p gpurun_out
CFGS=${1:-"0 8 9 10"}
ENVS=$(for c in $CFGS; do printf "MBPE_DEC_CFG=%s;" $c; done)
AB_ENV="${ENVS%;}" timeout 150 python tools/dec_ab.py 1024 > gpurun_out/dec_ab.log 2>&1; echo "dec ab rc=$?"
grep best gpurun_out/dec_ab.log | head -${2:-8}
T=$(echo $CFGS | awk '{print $2}')
MBPE_DEC_CFG=${T:-0} timeout 200 python -m pytest tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/dec_pytest.log 2>&1; echo "pytest MBPE_DEC_CFG=${T:-0} rc=$?"
tail -3 gpurun_out/dec_pytest.log | cut -c1-300

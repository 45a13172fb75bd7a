#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/enc_dbg.py 8 3 2 > gpurun_out/g_dbg.log 2>&1; echo "dbg rc=$?"; grep -v "same=True" gpurun_out/g_dbg.log | sort | uniq -c | sort -rn | head -20 | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
AB_ENV="MBPE_ENC_CARVEOUT=-;MBPE_ENC_CARVEOUT=100;MBPE_ENC_CARVEOUT=85;MBPE_ENC_CARVEOUT=70;MBPE_ENC_CARVEOUT=55" timeout 600 python tools/enc_ab.py 512 0 1 2 > gpurun_out/y3_carve.log 2>&1; echo "carve rc=$?"
grep -E "^cfg" gpurun_out/y3_carve.log
for c in 0 1 2; do
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,dram__bytes_read.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,launch__shared_mem_config_size,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_encode_tiles --launch-skip 11 --launch-count 1 --csv --log-file gpurun_out/y3_ncu_cfg$c.csv python tools/enc_ab.py 512 $c > gpurun_out/y3_ncu_$c.log 2>&1; echo "ncu cfg $c rc=$?"
grep -E "k_encode_tiles" gpurun_out/y3_ncu_cfg$c.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done

#!/bin/bash
mkdir -p gpurun_out
G="python scripts/bench_gemm_i8.py 4096 65536"
$G > gpurun_out/gemm_plain.log 2>&1 || exit 1
for sj in 2 4 8 16; do
MAF_OZ_SJ=$sj timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:ozaki_i8_gemm_v3 -c 6 --csv --log-file gpurun_out/ncu_sj$sj.csv $G > gpurun_out/ncu_sj.log 2>&1
echo "sj=$sj rc=$?"
done

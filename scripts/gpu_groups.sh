#!/bin/bash
# column-group sweep of the balanced schedule: DRAM bytes and time per launch of the INT8 contraction (65536 x 4096 x 4096)
mkdir -p gpurun_out
OUT=gpurun_out/i8gemm_v6_groups_sweep.txt; : > $OUT
for cfg in "1 4" "2 1" "2 2" "2 4" "3 1" "3 2" "3 4" "4 1" "4 2" "2 8" "1 2"; do
  set -- $cfg
  echo "== groups=$1 sj=$2" >> $OUT
  MAF_OZ_GROUPS=$1 MAF_OZ_SJ=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:ozaki_i8_gemm --csv --log-file gpurun_out/groups_tmp.csv python scripts/bench_gemm_i8.py 4096 65536 > /dev/null 2>&1
  python - >> $OUT <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/groups_tmp.csv') if not l.startswith('=='))]
by={}
for r in rows:
    by.setdefault(int(r['ID']),{})[r['Metric Name']]=(float(r['Metric Value'].replace(',','')), r['Metric Unit'])
def gb(v,u): return v*{'byte':1e-9,'Kbyte':1e-6,'Mbyte':1e-3,'Gbyte':1.0}[u]
def ms(v,u): return v*{'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3}.get(u, 1e-6)
for i,k in enumerate(sorted(by)):
    m=by[k]
    print("   launch %d: read %.2f GB write %.2f GB time %.2f ms (under ncu)" % (i, gb(*m['dram__bytes_read.sum']), gb(*m['dram__bytes_write.sum']), ms(*m['gpu__time_duration.sum'])))
PY
done
cat $OUT

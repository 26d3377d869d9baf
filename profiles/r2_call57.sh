#!/bin/bash
# Round 2 (session 2): ncu of the K8 word-vector kernels (the row the verdict asked to close with evidence).
O=gpurun_out/r2c57
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"wordvec" -c 6 --csv --log-file $O/wordvec.csv python bench.py --steps 2 --warmup 3 --no-first-stage --no-cpu-baseline --no-kernel-table > $O/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c57/wordvec.csv')) if len(r)>10]
hdr=rows[0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:28]),{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k, v)
PY

#!/bin/bash
O=gpurun_out/r2c41
mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"roi_" -c 24 --csv --log-file $O/roi_kernels.csv python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c41/roi_kernels.csv')) if len(r)>10]
hdr=rows[0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:50]),{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k, v)
PY

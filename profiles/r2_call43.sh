#!/bin/bash
O=gpurun_out/r2c43
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio --clock-control none -k regex:"nms_" -c 12 --csv --log-file $O/nms.csv python bench.py --steps 2 --warmup 3 --no-first-stage --no-cpu-baseline > $O/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c43/nms.csv')) if len(r)>10]
hdr=rows[0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:24]),{})[r[mi].split('.')[0][-28:]]=r[vi]
for k,v in cur.items(): print(k, v)
PY

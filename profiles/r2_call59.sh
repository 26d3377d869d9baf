#!/bin/bash
O=gpurun_out/r2c59
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_crop_maxpool_fwd_rows -c 1 -f -o /tmp/roi_fwd python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
ncu -i /tmp/roi_fwd.ncu-rep --page raw --csv > $O/roi_fwd_raw.csv 2>/dev/null
ncu -i /tmp/roi_fwd.ncu-rep --page source --csv --print-source sass > $O/roi_fwd_sass.csv 2>/dev/null
ls -la $O

#!/bin/bash
# Round 2 (session 2): ncu of the tile-owner ROI backward (first version).
O=gpurun_out/r2c25
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_tiles -c 2 -f -o $O/roi_tiles_v1 python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ncu -i $O/roi_tiles_v1.ncu-rep --page raw --csv > $O/roi_tiles_v1_raw.csv 2>/dev/null
ls -la $O

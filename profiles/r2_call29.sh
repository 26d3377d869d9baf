#!/bin/bash
O=gpurun_out/r2c29
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_tiles_bwd -s 2 -c 1 -f -o $O/roi_tiles_v4_fold python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
ncu -i $O/roi_tiles_v4_fold.ncu-rep --page raw --csv > $O/roi_tiles_v4_fold_raw.csv 2>/dev/null

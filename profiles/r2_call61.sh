#!/bin/bash
O=gpurun_out/r2c61
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_tiles_bwd -s 4 -c 1 -f -o /tmp/tiles_routed python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
ncu -i /tmp/tiles_routed.ncu-rep --page raw --csv > $O/tiles_routed_raw.csv 2>/dev/null
ncu -i /tmp/tiles_routed.ncu-rep --page source --csv --print-source sass > $O/tiles_routed_sass.csv 2>/dev/null
ls -la $O

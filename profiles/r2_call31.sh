#!/bin/bash
# Round 2 (session 2): tile-owner ROI backward v6 (four channels per lane) -- parity, timing, ncu.
O=gpurun_out/r2c31
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "roi" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python profiles/run_roi.py > $O/roi.json 2>&1; tail -c 1100 $O/roi.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_tiles_bwd -s 1 -c 2 -f -o $O/roi_tiles_v6 python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
ncu -i $O/roi_tiles_v6.ncu-rep --page raw --csv > $O/roi_tiles_v6_raw.csv 2>/dev/null

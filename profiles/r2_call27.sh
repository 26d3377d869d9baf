#!/bin/bash
# Round 2 (session 2): tile-owner ROI backward v3 (row-wise operand loads, depth as a template constant) -- parity, timing
# with and without the L2 prefetch of the next proposal, ncu.
O=gpurun_out/r2c27
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "roi" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python profiles/run_roi.py > $O/roi.json 2>&1; tail -c 700 $O/roi.json
C2D_ROI_TILES_PREFETCH=0 timeout 300 python profiles/run_roi.py > $O/roi_nopf.json 2>&1; tail -c 700 $O/roi_nopf.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_tiles_bwd -c 1 -f -o $O/roi_tiles_v3 python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
ncu -i $O/roi_tiles_v3.ncu-rep --page raw --csv > $O/roi_tiles_v3_raw.csv 2>/dev/null

#!/bin/bash
# Round 2, GPU call 5: forward v4 (plan padding, predicate codes) -- parity, timing, ncu.
O=gpurun_out/r2c5
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "roi" > $O/pytest_roi.log 2>&1; echo "rc=$?" >> $O/pytest_roi.log
C2D_ROI_FWD=old python profiles/run_roi.py --reps 7 --check --dump /tmp/old.pt > $O/roi_old.json 2> $O/roi_old.err
python profiles/run_roi.py --reps 7 --check --compare /tmp/old.pt > $O/roi_new.json 2> $O/roi_new.err
python profiles/run_roi.py --reps 7 --dtype f32 --check > $O/roi_new_f32.json 2>> $O/roi_new.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_crop -c 2 -o $O/roi_new \
    python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu_roi.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -n 4 $O/pytest_roi.log $O/pytest_all.log
tail -n 5 $O/roi_new.err

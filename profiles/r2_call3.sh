#!/bin/bash
# Round 2, GPU call 3: new ROI kernels -- parity, A/B against the round-1 kernels, ncu --set full.
O=gpurun_out/r2c3
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "roi" > $O/pytest_roi.log 2>&1; echo "rc=$?" >> $O/pytest_roi.log
C2D_ROI_FWD=old C2D_ROI_BWD=old python profiles/run_roi.py --reps 7 --check --dump /tmp/old.pt > $O/roi_old.json 2> $O/roi_old.err
python profiles/run_roi.py --reps 7 --check --compare /tmp/old.pt > $O/roi_new.json 2> $O/roi_new.err
python profiles/run_roi.py --reps 7 --dtype f32 --check > $O/roi_new_f32.json 2>> $O/roi_new.err
python profiles/run_roi.py --reps 7 --images 1 --check > $O/roi_new_b1.json 2>> $O/roi_new.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_crop -c 2 -o $O/roi_new \
    python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu_roi.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-first-stage --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -n 4 $O/pytest_roi.log $O/pytest_all.log
cat $O/roi_old.json $O/roi_new.json $O/roi_new_f32.json $O/roi_new_b1.json
tail -n 5 $O/roi_new.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c3/bench.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'hbm_group', d.get('hbm_group'))
PY

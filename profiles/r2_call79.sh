#!/bin/bash
# Round 2 (session 2): final check of HEAD: full GPU suite, full bench line, smoke.
O=gpurun_out/r2c79
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c79/bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['peak'], d['hbm_group']['frac'], d['eval_sweep']['images_per_sec'], d['eval_sweep']['e2e']['images_per_sec'], d['voc07_step']['ms_per_step'], d['clocks'])
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1

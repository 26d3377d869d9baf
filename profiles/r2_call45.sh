#!/bin/bash
# Round 2 (session 2): all evaluation scales through ONE head pass.
O=gpurun_out/r2c45
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-first-stage --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err | cut -c1-200
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
e=d['eval_sweep']
print('eval graphed', e['images_per_sec'], e['ms_per_image'], 'eager', e['eager_images_per_sec'], 'e2e', e['e2e']['images_per_sec'], e['k7_nms']['ms_per_pass'])"

#!/bin/bash
# Round 2 (session 2): eval sweep e2e with the next image's inputs staged on a copy stream.
O=gpurun_out/r2c48
mkdir -p $O
timeout 600 python bench.py --steps 5 --warmup 3 --no-first-stage --no-cpu-baseline --no-kernel-table > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err | cut -c1-200
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
e=d['eval_sweep']
print('eval graphed', e['images_per_sec'], e['ms_per_image'], 'eager', e['eager_images_per_sec'], 'e2e', e['e2e'])"

#!/bin/bash
# Round 2 (session 2): K7 per-class NMS without global-memory operations inside the serial tile walk.
O=gpurun_out/r2c42
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_reference_outputs.py -x -q -k "nms or postprocess or eval or full_size or predict" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-first-stage --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('eval', d['eval_sweep']['images_per_sec'], d['eval_sweep']['ms_per_image'], d['eval_sweep']['e2e']['images_per_sec'])
print([ (k['kernel'], round(k['ms'],4)) for k in d['kernels'] if 'nms' in k['kernel']])"

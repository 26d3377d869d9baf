#!/bin/bash
O=gpurun_out/r2c21
mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -n 5 $O/pytest_all.log
C2D_PROFILE_PER_LAUNCH=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $O/bench_1gpu.json 2> $O/bench_1gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c21/bench_1gpu.json'))
print('ms/step', round(d['ms_per_step'], 4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches'] / d['steps'])
r = d['roofline']
print({k: r['per_kernel'][k] for k in r['per_kernel']})
for i, e in enumerate(r.get('per_launch', [])): print(i, e['kind'], round(e['us'], 1), round(e['tflops']))
PY

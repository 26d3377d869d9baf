#!/bin/bash
O=gpurun_out/r2c19
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -n 6 $O/pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c19/bench_1gpu.json'))
print('ms/step', round(d['ms_per_step'], 4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches'] / d['steps'])
print(json.dumps({k: v for k, v in d['roofline'].items() if k != 'per_launch'})[:700])
print(json.dumps(d['hbm_group']), json.dumps(d['wordvec_extract'])[:300])
for k in d['kernels']: print('   ', k['kernel'], round(k['ms'], 4), round(k['frac'], 3))
PY

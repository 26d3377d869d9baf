#!/bin/bash
# Round 2 (session 2): the Mixed_5a pool term routed inside the head backward (side stream) instead of as a pre-pass of K1'.
O=gpurun_out/r2c54
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline --no-kernel-table"
for r in 1 2; do
  timeout 300 python bench.py $ARGS > $O/bench_$r.json 2> $O/bench_$r.err
  python -c "
import json
d=json.loads(open('$O/bench_$r.json').read().strip().splitlines()[-1])
print('step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']/20)"
done

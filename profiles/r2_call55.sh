#!/bin/bash
# Round 2 (session 2): A/B on one box -- pool term routed inside the head backward (1) or as a pre-pass of K1' (0).
O=gpurun_out/r2c55
mkdir -p $O
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline --no-kernel-table"
for m in 0 1 0 1; do
  C2D_HEAD_ROUTE=$m timeout 300 python bench.py $ARGS > $O/bench_$m.json 2> $O/bench_$m.err
  python -c "
import json
d=json.loads(open('$O/bench_$m.json').read().strip().splitlines()[-1])
print('route in head', $m, 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']/20)"
done

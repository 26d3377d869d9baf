#!/bin/bash
# Round 2 (session 2): weight-gradient launches on a side stream (fork / join inside c2d_head_mixed5_bwd).
O=gpurun_out/r2c51
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py -x -q > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline --no-kernel-table"
for m in 0 1 0 1; do
  C2D_WGRAD_STREAM=$m timeout 300 python bench.py $ARGS > $O/bench_side$m.json 2> $O/bench_side$m.err
  python -c "
import json
d=json.loads(open('$O/bench_side$m.json').read().strip().splitlines()[-1])
print('side stream', $m, 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['config']['step_launch'][:30])"
done

#!/bin/bash
# Round 2 (session 2): the pool branches of the head forward on the side stream (64-thread CTAs) beside the block's GEMMs.
O=gpurun_out/r2c65
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -k "bf16 or full_size or model or predictor or eval" > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
ARGS="--steps 20 --warmup 5 --no-first-stage --no-cpu-baseline --no-kernel-table"
for m in 0 1 0 1; do
  C2D_FWD_POOL_STREAM=$m timeout 300 python bench.py $ARGS > $O/bench_$m.json 2> $O/bench_$m.err
  python -c "
import json
d=json.loads(open('$O/bench_$m.json').read().strip().splitlines()[-1])
print('fwd pools on side', $m, 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval_sweep']['images_per_sec'],1), 'voc', round(d['voc07_step']['ms_per_step'],4))"
done

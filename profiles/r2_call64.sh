#!/bin/bash
# Round 2 (session 2): optimizer update beside the ROI backward (fork / join inside the step graph).
O=gpurun_out/r2c64
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -x -q -k "graphed or train_step or trainer or checkpoint or resume" > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline --no-kernel-table"
for m in 0 1 0 1; do
  C2D_OVERLAP_UPDATE=$m timeout 300 python bench.py $ARGS > $O/bench_$m.json 2> $O/bench_$m.err
  python -c "
import json
d=json.loads(open('$O/bench_$m.json').read().strip().splitlines()[-1])
print('overlap update', $m, 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'first stage step', round(d['with_first_stage']['ms_per_step'],4))" || tail -3 $O/bench_$m.err
done

#!/bin/bash
# Round 2 (session 2): wgrad operands staged with grouped 5-d TMA boxes (2-3 instead of 6-8 TMA instructions per k-step).
O=gpurun_out/r2c39
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_bf16.py -x -q > $O/pytest_bf16.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_bf16.log
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline"
for gr in 0 1; do
  C2D_WGRAD_GROUPED=$gr timeout 300 python bench.py $ARGS > $O/bench_gr$gr.json 2> $O/bench_gr$gr.err
  python -c "
import json
d=json.loads(open('$O/bench_gr$gr.json').read().strip().splitlines()[-1])
pk=d['roofline']['per_kernel']
print('grouped', $gr, 'step', round(d['ms_per_step'],4), 'wgrad', round(pk['wgrad_tc_kernel']['ms_per_step'],4), round(pk['wgrad_tc_kernel']['frac_of_peak'],3), 'conv', round(pk['conv_gemm_tc_kernel']['ms_per_step'],4))"
done

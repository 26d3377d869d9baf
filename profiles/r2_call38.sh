#!/bin/bash
# Round 2 (session 2): wgrad with TMA L2 prefetch k-steps ahead (0 = off, 3, 6, 12).
O=gpurun_out/r2c38
mkdir -p $O
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline --no-kernel-table"
for pf in 0 3 6 12; do
  C2D_WGRAD_PREFETCH=$pf timeout 600 python bench.py $ARGS > $O/bench_pf$pf.json 2> $O/bench_pf$pf.err
  python -c "
import json
d=json.loads(open('$O/bench_pf$pf.json').read().strip().splitlines()[-1])
pk=d['roofline']['per_kernel']
print('pf', $pf, 'step', round(d['ms_per_step'],4), 'wgrad', round(pk['wgrad_tc_kernel']['ms_per_step'],4), 'conv', round(pk['conv_gemm_tc_kernel']['ms_per_step'],4))"
done
C2D_WGRAD_PREFETCH=6 timeout 900 python -m pytest tests/test_gpu_bf16.py -x -q > $O/pytest_pf6.log 2>&1; tail -2 $O/pytest_pf6.log

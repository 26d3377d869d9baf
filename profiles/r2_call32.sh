#!/bin/bash
# Round 2 (session 2): the training step with (A) tile-owner K1' + folded pool backward, (B) tile-owner K1' + the head's
# own pool backward kernel, (C) per-proposal scatter + folded pool backward (the state before this session).
O=gpurun_out/r2c32
mkdir -p $O
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline --no-kernel-table"
timeout 600 python bench.py $ARGS > $O/bench_A.json 2> $O/bench_A.err
timeout 600 python -c "
import sys, runpy
import cap2det_b200.cap2det_model as m
m.Model.fold_pool_backward = False
sys.argv = ['bench.py'] + '$ARGS'.split()
runpy.run_path('bench.py', run_name='__main__')" > $O/bench_B.json 2> $O/bench_B.err
C2D_ROI_TILES=0 timeout 600 python bench.py $ARGS > $O/bench_C.json 2> $O/bench_C.err
for f in A B C; do python -c "
import json,sys
d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['value'], d['gpu_launches'], d['e2e']['value'])"; tail -2 $O/bench_$f.err; done

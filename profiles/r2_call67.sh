#!/bin/bash
# Round 2 (session 2): fused K3 epilogue A/B on one box (C2D_FUSE_K3=0 = separate avgpool_dropout_fwd kernel).
O=gpurun_out/r2c67
mkdir -p $O
for rep in 1 2; do
for f in 1 0; do
  C2D_FUSE_K3=$f timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench_f${f}_$rep.json 2> $O/bench.err
  python - $O/bench_f${f}_$rep.json $f <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('fuse', sys.argv[2], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['eval_sweep']['images_per_sec'], d['roofline']['frac'],
      [k['ms'] for k in d['kernels'] if 'fwd' in k['kernel'] and 'K2' in k['kernel']])
PY
done
done

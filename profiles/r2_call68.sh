#!/bin/bash
# Round 2 (session 2): fused K3 epilogue (keep factors prefetched, two TMEM loads in flight) A/B on one box.
O=gpurun_out/r2c68
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -1
for rep in 1 2 3; do
for f in 1 0; do
  C2D_FUSE_K3=$f timeout 900 python bench.py --steps 40 --warmup 5 --no-first-stage --no-cpu-baseline --no-extra-configs > $O/bench_f${f}_$rep.json 2> $O/bench.err
  python - $O/bench_f${f}_$rep.json $f <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('fuse', sys.argv[2], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'],
      [k['ms'] for k in d['kernels'] if 'fwd' in k['kernel'] and 'K2' in k['kernel']], d['clocks'])
PY
done
done

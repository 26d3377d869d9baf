#!/bin/bash
# Round 2 (session 2): per-launch times of the forward tensor-core launches with / without the fused K3 epilogue.
O=gpurun_out/r2c69
mkdir -p $O
for f in 1 0 1 0; do
  C2D_PROFILE_PER_LAUNCH=1 C2D_FUSE_K3=$f timeout 900 python bench.py --steps 10 --warmup 3 --no-first-stage --no-cpu-baseline --no-extra-configs > $O/bench_f$f.json 2> $O/bench.err
  python - $O/bench_f$f.json $f <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
pl = d['roofline']['per_launch']
print('fuse', sys.argv[2], round(d['ms_per_step'], 3), ' '.join('%d:%.1f/%.0f' % (i, e['us'], e['gflop']) for i, e in enumerate(pl[:14])))
PY
done

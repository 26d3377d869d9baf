#!/bin/bash
# Round 2 (session 2): the Mixed_5a pool backward as a dense pre-pass (roi_pool5a_route_kernel) + tile-owner K1' with two
# gradient inputs, against the per-bin fold.
O=gpurun_out/r2c40
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "roi or tile" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python profiles/run_roi.py > $O/roi.json 2>&1; python -c "
import json
d=json.loads(open('$O/roi.json').read().strip().splitlines()[-1])
for k,v in d.items():
  if isinstance(v,dict) and 'ms' in v: print(k, round(v['ms'],4))
print(d.get('fold_tiles_vs_scatter'))"
ARGS="--steps 20 --warmup 5 --no-extra-configs --no-first-stage --no-cpu-baseline --no-kernel-table"
for m in 1 0; do
  C2D_ROI_FOLD_ROUTED=$m timeout 300 python bench.py $ARGS > $O/bench_routed$m.json 2> $O/bench_routed$m.err
  python -c "
import json
d=json.loads(open('$O/bench_routed$m.json').read().strip().splitlines()[-1])
print('routed', $m, 'step', round(d['ms_per_step'],4), 'launches/step', d['gpu_launches']/20, 'e2e', round(d['e2e']['value']))"
done

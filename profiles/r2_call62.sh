#!/bin/bash
O=gpurun_out/r2c62
mkdir -p $O
for pf in 1 0 1 0; do
C2D_ROI_TILES_PREFETCH=$pf timeout 300 python profiles/run_roi.py > $O/roi_pf$pf.json 2>&1; python -c "
import json
d=json.loads(open('$O/roi_pf$pf.json').read().strip().splitlines()[-1])
print('prefetch', $pf, {k: round(v['ms'],4) for k,v in d.items() if isinstance(v,dict) and 'ms' in v and 'bwd' in k and 'scatter' not in k})"
done

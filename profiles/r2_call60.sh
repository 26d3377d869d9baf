#!/bin/bash
# Round 2 (session 2): K1 forward with the final lerp addition as fma(a, run-time 1.0, p): three packed instructions per lerp.
O=gpurun_out/r2c60
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_reference_outputs.py -x -q -k "roi or full_size or prediction" > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
timeout 300 python profiles/run_roi.py --check > $O/roi.json 2>&1; python -c "
import json
d=json.loads(open('$O/roi.json').read().strip().splitlines()[-1])
print({k: round(v['ms'],4) for k,v in d.items() if isinstance(v,dict) and 'ms' in v}, d.get('forward_bit_exact_vs_oracle'))"

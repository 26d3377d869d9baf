#!/bin/bash
# Round 2 (session 2): K3 (spatial mean + dropout) fused into the epilogue of the four launches that write Mixed_5c's
# output -- bf16 head tests, then the bench line.
O=gpurun_out/r2c66
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c66/bench.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['eval_sweep']['images_per_sec'], d['roofline']['frac'])
PY

#!/bin/bash
# Round 2 (session 2): K8 word-vector similarity, one CTA per token.
O=gpurun_out/r2c49
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_reference_outputs.py -x -q -k "word or vector or extractor or label" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-first-stage --no-cpu-baseline --no-kernel-table > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print(d['wordvec_extract'])"

#!/bin/bash
# Round 2 (session 2): state check at HEAD -- GPU suite, bench line, ROI kernels alone.
O=gpurun_out/r2c23
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 3000 $O/bench.json
timeout 300 python profiles/run_roi.py > $O/roi.json 2>&1; tail -c 1500 $O/roi.json

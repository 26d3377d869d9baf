#!/bin/bash
# Round 2 (session 2): tile-owner ROI backward -- parity, then timing against the per-proposal scatter.
O=gpurun_out/r2c24
mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "tile_owner" > $O/sanitizer.log 2>&1; tail -5 $O/sanitizer.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py -x -q -k "roi or bf16 or model" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 300 python profiles/run_roi.py > $O/roi.json 2>&1; tail -c 2500 $O/roi.json

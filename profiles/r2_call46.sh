#!/bin/bash
# Round 2 (session 2): compute-sanitizer memcheck over the kernels added in this session.
O=gpurun_out/r2c46
mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "tile_owner or nms or postprocess or predictor or optimizers or feature_map_dropout or roi" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "tile_owner and shape0 or nms_bit" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck.log

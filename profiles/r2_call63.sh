#!/bin/bash
# Round 2 (session 2): memcheck over the ROI tests (K1 forward changes) and the bf16 head tests (side stream).
O=gpurun_out/r2c63
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py -q -x -k "roi or head or bf16" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck.log

#!/bin/bash
O=gpurun_out/r2c58
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_scale_roi or evicts or predictor" > $O/pytest.log 2>&1; tail -5 $O/pytest.log

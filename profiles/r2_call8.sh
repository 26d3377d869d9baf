#!/bin/bash
O=gpurun_out/r2c8
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_bf16.py -q -s > $O/pytest_new.log 2>&1; echo "rc=$?" >> $O/pytest_new.log
grep -E "passed|failed|rc=|OICR stages|relu_decisions|Error|assert" $O/pytest_new.log | head -40

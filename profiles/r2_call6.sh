#!/bin/bash
# Round 2, GPU call 6: executed-reference parity tests on the GPU; bf16 gradient error exploration.
O=gpurun_out/r2c6
mkdir -p $O
timeout 900 python -m pytest tests/test_reference_outputs.py -q -m gpu > $O/pytest_ref.log 2>&1; echo "rc=$?" >> $O/pytest_ref.log
timeout 1200 python profiles/explore_bf16_grad.py > $O/bf16_grad.txt 2> $O/bf16_grad.err
tail -n 30 $O/pytest_ref.log
cat $O/bf16_grad.txt; tail -3 $O/bf16_grad.err

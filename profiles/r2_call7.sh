#!/bin/bash
O=gpurun_out/r2c7
mkdir -p $O
timeout 1200 python profiles/explore_bf16_grad.py > $O/bf16_grad.txt 2> $O/bf16_grad.err
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
cat $O/bf16_grad.txt; tail -3 $O/bf16_grad.err; tail -5 $O/pytest_all.log

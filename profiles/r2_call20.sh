#!/bin/bash
O=gpurun_out/r2c20
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -n 8 $O/pytest_all.log

#!/bin/bash
# Round 2 (session 2): new optimizer / dropout / moving-average tests + full GPU suite.
O=gpurun_out/r2c37
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -15 $O/pytest.log

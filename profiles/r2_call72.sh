#!/bin/bash
# Round 2 (session 2): the fused-vs-separate K3 test.
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -k "fused_k3" 2>&1 | tail -5

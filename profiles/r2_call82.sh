#!/bin/bash
# Round 2 (session 2): store-only microbenchmark of the epilogue's write pattern.
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_pattern profiles/microbench_store_pattern.cu && timeout 120 /tmp/store_pattern

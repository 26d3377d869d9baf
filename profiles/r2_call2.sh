#!/bin/bash
# Round 2, GPU call 2: L2-side microbenchmark (gather bandwidth, TMA bulk reduce) + launch list of eager steps.
O=gpurun_out/r2c2
mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbl2 profiles/microbench_l2.cu && timeout 300 /tmp/mbl2 > $O/microbench_l2.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-first-stage --no-cpu-baseline --no-kernel-table --no-cuda-graph > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
cat $O/microbench_l2.txt

#!/bin/bash
# Round 2 (session 2): launch list (durations) of eager steps at HEAD after K3 moved into the GEMM epilogue.
O=gpurun_out/r2c75
mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-extra-configs --no-kernel-table --no-first-stage --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv $CMD > $O/ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1
head -60 $O/launches.txt

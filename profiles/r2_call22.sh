#!/bin/bash
# Round 2: launch list (durations only) of eager steps after the epilogue rebuild.
O=gpurun_out/r2c22
mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-extra-configs --no-kernel-table --no-first-stage --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pool|roi_|oicr|midn" -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1
head -50 $O/launches.txt

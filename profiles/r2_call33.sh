#!/bin/bash
# Round 2 (session 2): launch list (durations) of eager steps at HEAD + ncu --set full of the tensor-core launches of one step.
O=gpurun_out/r2c33
mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-extra-configs --no-kernel-table --no-first-stage --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv $CMD > $O/ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1
head -70 $O/launches.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_tc|wgrad_tc" -s 210 -c 42 -f -o $O/tc_full $CMD > $O/ncu_tc.log 2>&1
ncu -i $O/tc_full.ncu-rep --page raw --csv > $O/tc_full_raw.csv 2>/dev/null
ls -la $O

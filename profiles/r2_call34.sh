#!/bin/bash
# Round 2 (session 2): ncu --set full of the tensor-core launches of one eager step at HEAD (raw page only comes back).
O=gpurun_out/r2c34
mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-extra-configs --no-kernel-table --no-first-stage --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none -k regex:"conv_gemm_tc|wgrad_tc" -s 210 -c 42 -f -o /tmp/tc_full $CMD > $O/ncu_tc.log 2>&1
ncu -i /tmp/tc_full.ncu-rep --page raw --csv > $O/tc_full_raw.csv 2>/dev/null
ls -la $O

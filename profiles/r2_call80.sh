#!/bin/bash
# Round 2 (session 2): PC sampling of the shortest-K tensor-core launch (128 -> 1024 data gradient): where its warps wait.
O=gpurun_out/r2c80
mkdir -p $O
timeout 300 python profiles/run_short_k_dgrad.py 3 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc2 --launch-skip 2 -c 1 -f -o /tmp/shortk python profiles/run_short_k_dgrad.py 2 > $O/ncu.log 2>&1
ncu -i /tmp/shortk.ncu-rep --page source --csv --print-source sass > $O/shortk_source.csv 2>/dev/null
ncu -i /tmp/shortk.ncu-rep --page raw --csv > $O/shortk_raw.csv 2>/dev/null
python profiles/analyze_ncu_source.py $O/shortk_source.csv | head -70
ls -la $O

#!/bin/bash
# Round 2: ncu --set full of the 42 tensor-core launches of one eager step (uniform-issue build) + source-level
# captures of two slow launches (#0 Mixed_5a sibling forward, #41 merged Mixed_5a data gradient); word-vector timing.
O=gpurun_out/r2c17
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-extra-configs --no-kernel-table --no-first-stage --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none -k regex:"conv_gemm|wgrad_tc" -s 84 -c 42 -o /tmp/tc_full $CMD > $O/ncu_tc.log 2>&1
ncu -i /tmp/tc_full.ncu-rep --page raw --csv > $O/r2_ncu_full_tc_kernels.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|wgrad_tc" -s 84 -c 1 -o $O/tc_launch0 $CMD > $O/ncu_tc0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|wgrad_tc" -s 125 -c 1 -o $O/tc_launch41 $CMD > $O/ncu_tc41.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc" -s 30 -c 1 -o $O/tc_wgrad $CMD > $O/ncu_wg.log 2>&1
python - <<'PY' > $O/wordvec.txt 2>&1
import sys, tempfile, json
sys.path.insert(0, '.')
import torch, bench
torch.cuda.set_device(0)
print(json.dumps(bench.wordvec_extract(torch.device('cuda', 0), tempfile.mkdtemp())))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_outputs.py tests/test_gpu_fullsize.py -q -k "word or label" > $O/pytest_labels.log 2>&1; tail -3 $O/pytest_labels.log
cat $O/wordvec.txt | tail -2
ls -la $O

#!/bin/bash
# Round 2, GPU call 1: (a) baseline ncu --set full of the two ROI kernels, (b) pipe-rate microbenchmark,
# (c) parity + bench of the three prepared tensor-core switches, one by one.
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import torch" 2>/dev/null
# (a) ROI kernels: timing, then one full capture (fwd + bwd launch each)
python profiles/run_roi.py --reps 7 --check > $O/roi_baseline.json 2> $O/roi_baseline.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_crop -c 2 -o $O/roi_baseline \
    python profiles/run_roi.py --reps 1 --warm 0 > $O/ncu_roi.log 2>&1
# (b) microbenchmark
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb profiles/microbench_pipes.cu && timeout 300 /tmp/mb > $O/microbench.txt 2>&1
# (c) tensor-core switches
run_variant() {   # name, env assignments...
  local name=$1; shift
  env "$@" timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_backbone.py -x -q > $O/pytest_$name.log 2>&1
  echo "pytest rc=$?" >> $O/pytest_$name.log
  env "$@" C2D_PROFILE_PER_LAUNCH=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-first-stage --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err
}
run_variant base C2D_UNIFORM_ISSUE=0
run_variant U C2D_UNIFORM_ISSUE=1
run_variant UE C2D_UNIFORM_ISSUE=1 C2D_EPILOGUE_EARLY_SHIFT=1
run_variant UEP C2D_UNIFORM_ISSUE=1 C2D_EPILOGUE_EARLY_SHIFT=1 C2D_EPILOGUE_PREFETCH=1
run_variant P C2D_EPILOGUE_PREFETCH=1
# leave the default library in place
python cap2det_b200/build.py > /dev/null 2>&1
tail -n 3 $O/pytest_*.log
for f in $O/bench_*.json; do python - "$f" <<'EOF'
import json, sys
try:
  d = json.load(open(sys.argv[1]))
  r = d['roofline']
  print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 4), 'conv', {k: round(v['ms_per_step'], 4) for k, v in r['per_kernel'].items()})
except Exception as e:
  print(sys.argv[1], 'unreadable', e)
EOF
done
cat $O/microbench.txt
cat $O/roi_baseline.json

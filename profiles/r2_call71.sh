#!/bin/bash
# Round 2 (session 2): 2-GPU check at HEAD (fused K3 epilogue): the multi-GPU test and the 2-rank bench line.
O=gpurun_out/r2c71
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2c71/bench_2gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['config'].get('step_launch'), d['eval_sweep']['images_per_sec'])
PY

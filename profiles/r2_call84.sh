#!/bin/bash
# Round 2 (session 2): 4-GPU bench line at HEAD (completes the 1 / 2 / 4 / 8 set under profiles/r2).
O=gpurun_out/r2c84
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_4gpu.json 2> $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2c84/bench_4gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['eval_sweep']['images_per_sec'])
PY

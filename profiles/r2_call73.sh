#!/bin/bash
# Round 2 (session 2): 8-GPU bench line at HEAD (fused K3 epilogue).
O=gpurun_out/r2c73
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_8gpu.json 2> $O/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2c73/bench_8gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['eval_sweep']['images_per_sec'])
PY

#!/bin/bash
# Round 2 (session 2): 8-GPU data-parallel bench (three graphs + all-reduce, sharded eval sweep).
O=gpurun_out/r2c47
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_8gpu.json 2> $O/bench_8gpu.err; tail -2 $O/bench_8gpu.err | cut -c1-300
python -c "
import json
d=json.loads(open('$O/bench_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['eval_sweep']['images_per_sec'], d['config']['step_launch'])"

#!/bin/bash
# Round 2, GPU call 9 (2 GPUs): data-parallel CUDA-graph step -- replica test, 2-GPU bench, 1-GPU bench with all configs.
O=gpurun_out/r2c9
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
tail -n 15 $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-first-stage > $O/bench_2gpu.json 2> $O/bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-first-stage --no-cuda-graph --no-extra-configs > $O/bench_2gpu_eager.json 2> $O/bench_2gpu_eager.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
tail -n 5 $O/bench_2gpu.err $O/bench_1gpu.err
python - <<'PY'
import json
for f in ('bench_2gpu', 'bench_2gpu_eager', 'bench_1gpu'):
  try:
    d = json.load(open('gpurun_out/r2c9/%s.json' % f))
    print(f, 'ms/step', round(d['ms_per_step'], 4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('host_busy_ms_per_step'),
          d['config']['step_launch'], 'replicas', d.get('replicas_identical'), 'launches', d['gpu_launches'])
    for k in ('eval_sweep', 'voc07_step', 'wordvec_extract', 'hbm_group', 'cpu_baseline'):
      if k in d: print('   ', k, json.dumps(d[k])[:600])
    print('    roofline', json.dumps({k: v for k, v in d['roofline'].items() if k != 'per_launch'})[:900])
  except Exception as e:
    print(f, 'unreadable', e)
PY

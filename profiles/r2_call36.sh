#!/bin/bash
# Round 2 (session 2): 2-GPU data-parallel step (three graphs + all-reduce) with the tile-owner K1', and the 2-GPU test.
O=gpurun_out/r2c36
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 1500 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log

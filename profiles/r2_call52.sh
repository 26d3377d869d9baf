#!/bin/bash
# Round 2 (session 2): 2-rank worker (eager vs graphed steps) with the weight-gradient side stream off / on, three runs each.
O=gpurun_out/r2c52
mkdir -p $O
for m in 0 1; do for r in 1 2 3; do
  C2D_WGRAD_STREAM=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$r tests/ddp_worker.py 2>/dev/null | grep DDP_ | sed "s/^/side=$m run=$r /"
done; done

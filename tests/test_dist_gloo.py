"""CPU, world_size 2 over gloo: the host logic of the N>1 path (image sharding, gradient all-reduce +
1/G averaging, max-over-ranks timing).  The CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cap2det_b200 import dist as c2d_dist


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    # each rank owns a different local batch => different local gradients
    g = [torch.full((5,), float(rank + 1)), torch.arange(4, dtype=torch.float32) * (rank + 1)]
    c2d_dist.allreduce_sum(g)
    avg = [x / world for x in g]          # the Adagrad kernel applies grad_scale = 1/G
    t = c2d_dist.max_over_ranks(10.0 + rank)
    shard = c2d_dist.shard_indices(11, rank, world, 'strided')
    shard_c = c2d_dist.shard_indices(11, rank, world, 'contiguous')
    out[rank] = (avg[0].tolist(), avg[1].tolist(), t, shard, shard_c)
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_sharding_and_timing():
  world = 2
  port = _free_port()
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  assert set(out.keys()) == {0, 1}
  for r in range(world):
    a0, a1, t, shard, shard_c = out[r]
    assert a0 == [1.5] * 5                      # mean of 1 and 2
    assert a1 == [0.0, 1.5, 3.0, 4.5]           # mean of k and 2k
    assert t == 11.0                            # slowest rank
  assert sorted(out[0][3] + out[1][3]) == list(range(11))      # every image exactly once
  assert sorted(out[0][4] + out[1][4]) == list(range(11))
  assert out[0][4] == [0, 1, 2, 3, 4, 5] and out[1][4] == [6, 7, 8, 9, 10]


def test_single_process_is_a_noop():
  g = [torch.ones(3)]
  c2d_dist.allreduce_sum(g)
  assert g[0].tolist() == [1.0, 1.0, 1.0]
  assert c2d_dist.max_over_ranks(3.5) == 3.5
  assert c2d_dist.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]

"""Data-parallel plumbing (one process per GPU, torch.distributed).  Images are independent units of the
path: eval/predict shards images with no communication (the reference's analogue is the hash-bucket
`shard_indicator "i/n"` filter, readers/cap2det_reader.py:201-211); training adds ONE all-reduce of the
gradient buffers per step."""
import torch
import torch.distributed as dist


def world():
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def shard_indices(num_items, rank, world_size, mode='strided'):
  """Indices of the images rank `rank` owns.  'strided': r, r+G, ... ; 'contiguous': equal blocks (the
  first num_items % G ranks get one more)."""
  if not 0 <= rank < world_size:
    raise ValueError('rank %d outside world of size %d' % (rank, world_size))
  if mode == 'strided':
    return list(range(rank, num_items, world_size))
  if mode == 'contiguous':
    base, extra = divmod(num_items, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))
  raise ValueError('unknown shard mode %r' % mode)


def allreduce_sum(tensors):
  """In-place sum over ranks of every tensor (no-op in a single process)."""
  if world()[1] == 1:
    return
  for t in tensors:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)


def max_over_ranks(value, device='cpu'):
  """Max over ranks of a python float (bench timing rule: the slowest rank defines the step time)."""
  if world()[1] == 1:
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())

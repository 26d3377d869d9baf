"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run with 2+ ranks, one GPU each.  Every rank
trains on DIFFERENT data for a few steps; with a correct gradient all-reduce the replicas stay bit-identical."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap2det_b200 import builder, config, synthetic, trainer  # noqa: E402
from cap2det_b200.standard_fields import InputDataFields as F  # noqa: E402


def main():
  rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=True, head_dtype=torch.bfloat16, first_stage=True)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
  for v in model.get_variables_to_train():
    dist.broadcast(v.data, src=0)
  step = trainer.TrainStep(model, learning_rate=0.01, world_size=world)
  rng = np.random.default_rng(100 + rank)                      # different data per rank
  B, P, H, W = 1, 32, 128, 160
  for it in range(4):
    ex = {F.image: torch.from_numpy(rng.integers(0, 256, size=(B, H, W, 3)).astype(np.uint8)).cuda(),
          F.proposals: torch.from_numpy(synthetic.make_proposals(rng, B, P, H, W)).cuda(),
          F.num_proposals: torch.full((B,), P, dtype=torch.int32, device='cuda'),
          F.object_texts: synthetic.make_object_texts(rng, B, classes)}
    total = step(ex)
  model.raise_if_assert_failed()
  ok = bool(torch.isfinite(total))
  for v in model.get_variables_to_train():
    ref = v.detach().clone()
    dist.broadcast(ref, src=0)
    ok = ok and bool(torch.equal(ref, v.detach()))
  flag = torch.tensor([1 if ok else 0], device='cuda')
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print('DDP_REPLICAS_IDENTICAL' if int(flag.item()) == 1 else 'DDP_REPLICAS_DIVERGED')
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()

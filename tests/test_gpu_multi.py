"""Multi-GPU check (skipped on boxes with one GPU): replicas trained on different data stay bit-identical, i.e.
the overlapped gradient all-reduce of trainer.TrainStep reduces every gradient buffer exactly once per step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_ranks_stay_identical():
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
         '--master-port', '29561', os.path.join(root, 'tests', 'ddp_worker.py')]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=root)
  assert 'DDP_REPLICAS_IDENTICAL' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
  assert 'DDP_GRAPHED_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]

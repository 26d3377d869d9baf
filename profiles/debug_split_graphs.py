import os, sys, tempfile
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import numpy as np, torch
from cap2det_b200 import builder, config, synthetic, trainer
from cap2det_b200.standard_fields import InputDataFields as F
torch.cuda.set_device(0)
torch.cuda.set_stream(torch.cuda.Stream())
d = tempfile.mkdtemp()
classes = synthetic.VOC_CLASSES
text = synthetic.model_options_text(extractor='groundtruth_extractor', extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
m = config.Model(); m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
res = []
for split in (False, True):
  model = builder.build(m, is_training=True, head_dtype=torch.bfloat16, seed=3)
  with torch.no_grad(): model.fc_weights.mul_(8.0)
  step = trainer.TrainStep(model, learning_rate=0.01)
  rng = np.random.default_rng(200)
  B, P = 1, 48
  def batch():
    return {F.features_to_crop: torch.from_numpy(synthetic.make_feature_map(rng, B, 160, 208)).cuda().requires_grad_(True),
            F.proposals: torch.from_numpy(synthetic.make_proposals(rng, B, P, 160, 208)).cuda(),
            F.num_proposals: torch.full((B,), P, dtype=torch.int32, device='cuda'),
            F.object_texts: synthetic.make_object_texts(rng, B, classes),
            F.dropout_keep_mask: torch.from_numpy((rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)).cuda()}
  batches = [batch() for _ in range(3)]
  run = trainer.GraphedTrainStep(step, batches[0], split_graphs=split)
  out = [float(run(ex)) for ex in batches]
  torch.cuda.synchronize()
  print('split', split, out)
  res.append(out)
print('OK' if np.allclose(res[0], res[1], rtol=2e-3) else 'MISMATCH')

"""One eager training step at the BASELINE configuration (dropout on), for compute-sanitizer / debugging:
    compute-sanitizer --tool memcheck --print-limit 5 python tools/one_step.py [batch] [npoints]"""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from coda_neurips2023_b200 import synthetic  # noqa: E402
from coda_neurips2023_b200.criterion import build_criterion  # noqa: E402
from coda_neurips2023_b200.engine import TrainStep  # noqa: E402
from coda_neurips2023_b200.models import build_model  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
npoints = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
args = synthetic.make_args()
cfg = synthetic.SyntheticDatasetConfig(args)
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model, _ = build_model(args, cfg)
model = model.cuda().train()
crit = build_criterion(args, cfg).cuda()
step = TrainStep(args, model, crit, torch.device("cuda", 0))
b = synthetic.to_device(synthetic.make_batch(batch, npoints, seed=0), "cuda")
np.random.seed(0)
for i in range(2):
    loss, _ = step(b, 0.0)
    torch.cuda.synchronize()
    print("step", i, "loss", float(loss))

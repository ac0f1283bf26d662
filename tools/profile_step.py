"""Per-kernel time breakdown of one training step with torch.profiler (CUPTI).
Not a benchmark: numbers under the profiler are only used as SHARES."""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from coda_neurips2023_b200 import synthetic  # noqa: E402
from coda_neurips2023_b200.criterion import build_criterion  # noqa: E402
from coda_neurips2023_b200.engine import TrainStep  # noqa: E402
from coda_neurips2023_b200.models import build_model  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
args = synthetic.make_args()
cfg = synthetic.SyntheticDatasetConfig(args)
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model, _ = build_model(args, cfg)
model = model.cuda().train()
criterion = build_criterion(args, cfg).cuda()
step = TrainStep(args, model, criterion, torch.device("cuda"))
batch = synthetic.to_device(synthetic.make_batch(8, 20000, seed=0), "cuda")
for _ in range(3):
    step(batch, 0.0)
torch.cuda.synchronize()

# coarse phase timing with CUDA events (outside the profiler)
def timed(fn, reps=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

ms_step, _ = timed(lambda: step(batch, 0.0))
with torch.no_grad():
    ms_fwd, out = timed(lambda: model(batch, curr_epoch=0))
    ms_enc, _ = timed(lambda: model.run_encoder(batch["point_clouds"]))
    ms_pre, _ = timed(lambda: model.pre_encoder(batch["point_clouds"][..., :3].contiguous()))
    ms_crit, _ = timed(lambda: criterion(out, dict(batch)))
    ms_clip, _ = timed(lambda: model.get_predicted_box_clip_embedding(batch, dict(out["outputs"]), curr_epoch=0))
print(f"PHASES ms: step {ms_step:.2f} | fwd(no grad) {ms_fwd:.2f} | pre_encoder {ms_pre:.2f} | "
      f"pre_encoder+encoder {ms_enc:.2f} | clip branch {ms_clip:.2f} | criterion(no grad) {ms_crit:.2f}")

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(batch, 0.0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=90, max_name_column_width=70))


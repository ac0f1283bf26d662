"""Micro-benchmark of the crop / resize kernel (include/coda_image.h) at the BASELINE shape: 8 images 530 x 730,
256 crops -> 224 x 224, for each tile height.  CUDA events, L2 flushed between launches by the 77 MB output itself."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from coda_neurips2023_b200 import ops  # noqa: E402

rng = np.random.default_rng(0)
h, w, n = 530, 730, 256
imgs = torch.from_numpy(rng.integers(0, 256, size=(8, h, w, 3), dtype=np.uint8)).cuda()
for name, lo_frac in (("large boxes (random-init predictions)", 0.6), ("mixed boxes", 0.1)):
    bw = rng.integers(int(lo_frac * w), w + 1, n); bh = rng.integers(int(lo_frac * h), h + 1, n)
    x0 = rng.integers(0, w - bw + 1); y0 = rng.integers(0, h - bh + 1)
    boxes = torch.from_numpy(np.stack([x0, y0, x0 + bw, y0 + bh], 1).astype(np.int32)).cuda()
    scene = torch.arange(8, dtype=torch.int32).repeat_interleave(32).cuda()
    valid = torch.ones(n, dtype=torch.bool).cuda()
    ref = None
    for tr in (0, 16, 8, 4, 2):
        for patch in (0, 32):
            f = lambda: ops.crop_resize_normalize(imgs, scene, boxes, valid, 224, dtype=torch.float16, patch=patch,  # noqa: E731
                                                  tile_rows=tr)
            out = f()
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                f()
            e1.record()
            torch.cuda.synchronize()
            if patch == 0:
                if ref is None:
                    ref = out
                same = torch.equal(out, ref)
            else:
                same = torch.equal(out, ref.view(n, 3, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5))
            print(f"{name}: tile_rows {tr:2d} patch {patch:2d}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us  same={same}")

"""Launch the step's representative GEMM shapes a few times each (an ncu target:
    ncu --set full --import-source on -k regex:gemm_nt -s 8 -c 4 -o out python tools/gemm_once.py)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from coda_neurips2023_b200 import ops  # noqa: E402

SHAPES = [  # m, n, k, nsplit, fp16
    (16384, 768, 256, 3, False),     # encoder qkv projection
    (1048576, 128, 64, 3, False),    # SA shared MLP layer
    (12800, 3072, 768, 1, True),     # CLIP MLP c_fc
    (12800, 768, 3072, 1, True),     # CLIP MLP c_proj
]
REPS = 3

for rep in range(REPS):
    for (m, n, k, ns, fp16) in SHAPES:
        if fp16:
            a = (torch.randn(1, 1, m, k, device="cuda") * 0.1).half()
            b = (torch.randn(1, 1, n, k, device="cuda") * 0.1).half()
            out = torch.empty(1, m, n, device="cuda", dtype=torch.float16)
            ops.gemm_nt(a, b, m, n, out=out, bias=torch.zeros(n, device="cuda"))
        else:
            a = torch.randn(m, k, device="cuda")
            b = torch.randn(n, k, device="cuda")
            ap, bp = ops.pack_split(a, m, k, k, 1, ns), ops.pack_split(b, n, k, k, 1, ns)
            out = torch.empty(1, m, n, device="cuda")
            ops.gemm_nt(ap, bp, m, n, out=out)
torch.cuda.synchronize()
print("done")

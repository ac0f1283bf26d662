import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from coda_neurips2023_b200 import ops  # noqa: E402

def timeit(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for (m, n, k) in [(12800, 2304, 768), (12800, 768, 768), (12800, 3072, 768), (12800, 768, 3072), (12544, 768, 3072)]:
    a = (torch.randn(1, 1, m, k, device="cuda") * 0.1).half()
    b = (torch.randn(1, 1, n, k, device="cuda") * 0.1).half()
    bias = torch.randn(n, device="cuda")
    o32 = torch.empty(1, m, n, device="cuda")
    o16 = torch.empty(1, m, n, device="cuda", dtype=torch.float16)
    row = {"m": m, "n": n, "k": k}
    row["out_fp32_ms"] = round(timeit(lambda: ops.gemm_nt(a, b, m, n, bias=bias, out=o32)), 4)
    row["out_fp16_ms"] = round(timeit(lambda: ops.gemm_nt(a, b, m, n, bias=bias, out=o16)), 4)
    row["out_fp16_gelu_ms"] = round(timeit(lambda: ops.gemm_nt(a, b, m, n, bias=bias, out=o16, act=2)), 4)
    row["out_fp16_nobias_ms"] = round(timeit(lambda: ops.gemm_nt(a, b, m, n, out=o16)), 4)
    a2, b2 = a[0, 0], b[0, 0]
    row["cublas_fp16_ms"] = round(timeit(lambda: torch.nn.functional.linear(a2, b2)), 4)
    row["tflops_fp16_out16"] = round(2.0 * m * n * k / row["out_fp16_ms"] / 1e9, 1)
    print(json.dumps(row), flush=True)

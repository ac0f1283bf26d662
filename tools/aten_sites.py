"""Which lines of the package launch the ATen kernels that remain in the step?

One eager training step under a TorchDispatchMode: every aten op that reaches the CUDA backend is logged with its
operand shapes, the innermost package frames of the Python stack (forward) or the autograd node that is running
(backward).  Output: sites sorted by element traffic.  Attribution tool, not a benchmark.

    python tools/aten_sites.py [--top 70]
"""
import collections
import sys
import traceback
import warnings
from pathlib import Path

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from coda_neurips2023_b200 import synthetic  # noqa: E402
from coda_neurips2023_b200.criterion import build_criterion  # noqa: E402
from coda_neurips2023_b200.engine import TrainStep  # noqa: E402
from coda_neurips2023_b200.models import build_model  # noqa: E402

SKIP = {"aten::view", "aten::_unsafe_view", "aten::reshape", "aten::t", "aten::transpose", "aten::permute",
        "aten::expand", "aten::slice", "aten::select", "aten::unsqueeze", "aten::squeeze", "aten::detach",
        "aten::alias", "aten::as_strided", "aten::empty", "aten::empty_like", "aten::empty_strided",
        "aten::unbind", "aten::split", "aten::split_with_sizes", "aten::narrow", "aten::unflatten",
        "aten::flatten", "aten::view_as", "aten::lift_fresh", "aten::_local_scalar_dense", "aten::chunk",
        "aten::diagonal", "aten::movedim", "aten::new_empty", "aten::new_empty_strided", "aten::is_same_size",
        "aten::record_stream", "aten::result_type", "aten::size", "aten::stride"}


def numel(x):
    n = 0
    if isinstance(x, torch.Tensor):
        n = x.numel() if x.is_cuda else 0
    elif isinstance(x, (list, tuple)):
        n = sum(numel(v) for v in x)
    return n


def shapes(args):
    out = []
    for a in args:
        if isinstance(a, torch.Tensor):
            out.append("x".join(map(str, a.shape)) + ("" if a.is_contiguous() else "s"))
        elif isinstance(a, (list, tuple)) and a and isinstance(a[0], torch.Tensor):
            out.append("[" + ",".join("x".join(map(str, t.shape)) for t in a[:3]) + (",.." if len(a) > 3 else "") + "]")
    return " ".join(out)


class Sites(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.log = collections.defaultdict(lambda: [0, 0])

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = func._schema.name
        if name in SKIP:
            return out
        elems = numel(args) + numel(out)
        if elems == 0:
            return out
        node = torch._C._current_autograd_node()
        frames = [f for f in traceback.extract_stack(limit=40)
                  if "coda_neurips2023_b200" in f.filename and "tools/" not in f.filename]
        where = " < ".join(f"{Path(f.filename).name}:{f.lineno}" for f in reversed(frames[-3:]))
        if node is not None:
            where = f"[bwd {node.name()}] " + where
        key = (name.replace("aten::", ""), shapes(args), where)
        self.log[key][0] += 1
        self.log[key][1] += elems
        return out


def main():
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 70
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
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
    for _ in range(2):
        step(batch, 0.0)
    torch.cuda.synchronize()
    mode = Sites()
    with mode:
        step(batch, 0.0)
    torch.cuda.synchronize()
    rows = sorted(mode.log.items(), key=lambda kv: -kv[1][1])
    total_calls = sum(v[0] for v in mode.log.values())
    total_elems = sum(v[1] for v in mode.log.values())
    print(f"# {total_calls} ATen calls on CUDA tensors in one eager step, {total_elems / 1e6:.1f} M elements touched")
    by_op = collections.Counter()
    for (op, _, _), (c, e) in mode.log.items():
        by_op[op] += c
    print("# calls by op: " + ", ".join(f"{k}={v}" for k, v in by_op.most_common(25)))
    print(f"{'calls':>5} {'Melem':>8}  op | shapes | site")
    for (op, shp, where), (c, e) in rows[:top]:
        print(f"{c:5d} {e / 1e6:8.2f}  {op} | {shp} | {where}")
    # and by call count: the small launches
    print("\n# by call count")
    for (op, shp, where), (c, e) in sorted(mode.log.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{c:5d} {e / 1e6:8.2f}  {op} | {shp} | {where}")


if __name__ == "__main__":
    main()

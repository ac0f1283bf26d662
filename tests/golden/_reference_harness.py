"""Imports the REFERENCE (read-only at /root/reference) in this container so that
golden vectors can be generated from its own Python code on CPU.

Only the scripts in tests/golden/ (and tests that are explicitly skipped when
/root/reference is absent) use this; it never travels to the GPU box.

What is stubbed, and why:
  * `pointnet2._ext` -- the reference's extension is CUDA-only; here it is backed by
    the CPU oracle (oracle/pointnet2_oracle.c), which is itself pinned to the real
    extension by tests/golden/pointnet2_ref_*.npz.
  * `timm`, `plyfile`, `trimesh`, `ftfy`, `tensorboardX`, `models.vision_transformer`,
    `models.resnet`: imported by reference modules but unused on the path; not installed.
  * `torch.Tensor.cuda` / `.to('cuda')` style calls are redirected to CPU.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[2]


def available() -> bool:
    return REF.exists()


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _oracle_ext():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_pointnet2 as orc

    def t(a, dtype=None):
        return torch.from_numpy(np.ascontiguousarray(a))

    ext = types.ModuleType("pointnet2._ext")
    ext.furthest_point_sampling = lambda p, n: t(orc.furthest_point_sampling(p.detach().numpy(), int(n)))
    ext.gather_points = lambda p, i: t(orc.gather_points(p.detach().numpy(), i.numpy()))
    ext.gather_points_grad = lambda g, i, n: t(orc.gather_points_grad(g.detach().numpy(), i.numpy(), int(n)))
    ext.ball_query = lambda nx, x, r, ns: t(orc.ball_query(nx.detach().numpy(), x.detach().numpy(), float(r), int(ns)))
    ext.group_points = lambda p, i: t(orc.group_points(p.detach().numpy(), i.numpy()))
    ext.group_points_grad = lambda g, i, n: t(orc.group_points_grad(g.detach().numpy(), i.numpy(), int(n)))
    ext.three_nn = lambda u, k: [t(a) for a in orc.three_nn(u.detach().numpy(), k.detach().numpy())]
    ext.three_interpolate = lambda p, i, w: t(orc.three_interpolate(p.detach().numpy(), i.numpy(), w.detach().numpy()))
    ext.three_interpolate_grad = lambda g, i, w, m: t(
        orc.three_interpolate_grad(g.detach().numpy(), i.numpy(), w.detach().numpy(), int(m)))
    return ext


_INSTALLED = False


def install():
    """Puts the reference on sys.path with the stubs above.  Idempotent."""
    global _INSTALLED
    if _INSTALLED:
        return
    assert available(), "/root/reference is not mounted"
    os.chdir(REF)  # the reference opens datasets/*.npy by relative path
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REF / "third_party_pointnet2" / "pointnet2"))
    pkg = _stub("pointnet2")
    pkg.__path__ = []
    ext = _oracle_ext()
    sys.modules["pointnet2._ext"] = ext
    pkg._ext = ext
    _stub("plyfile", PlyData=None, PlyElement=None)
    _stub("trimesh")
    mpl = _stub("matplotlib")
    mpl.__path__ = []
    mpl.use = lambda *a, **k: None
    _stub("matplotlib.pyplot")
    _stub("matplotlib.cm")
    _stub("tensorboardX", SummaryWriter=None)
    # `datasets` is a package whose __init__ pulls every dataset module (matplotlib, ...); the
    # model only needs datasets.sunrgbd_utils / scannet_utils, so root the package without __init__
    ds = _stub("datasets")
    ds.__path__ = [str(REF / "datasets")]
    _stub("ftfy", fix_text=lambda s: s)
    timm = _stub("timm")
    timm.__path__ = []
    _stub("timm.data").__path__ = []
    _stub("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
          IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225), DEFAULT_CROP_PCT=0.875)
    # `models` must be a package rooted at the reference, without running its __init__
    models = _stub("models")
    models.__path__ = [str(REF / "models")]
    _stub("models.vision_transformer", _create_vision_transformer=None,
          _create_multi_modal_vision_transformer=None, _create_two_modal_vision_transformer=None)
    _stub("models.resnet", resnet50=None)
    # CPU redirection of explicit .cuda() / device='cuda'
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_to = torch.Tensor.to

    def _to(self, *args, **kwargs):
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
        if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
            kwargs["device"] = "cpu"
        return _orig_to(self, *args, **kwargs)

    torch.Tensor.to = _to
    _INSTALLED = True


def load(module: str):
    """import_module inside the reference tree, e.g. load('models.transformer')."""
    install()
    import importlib

    return importlib.import_module(module)

"""Deterministic, name-keyed parameter fill shared by the golden generator (reference
model, CPU) and the parity tests (our model, GPU): no weights need to be stored."""
import zlib

import torch


@torch.no_grad()
def fill_by_name(module: torch.nn.Module, seed: int = 0) -> None:
    state = dict(module.named_parameters())
    state.update({k: v for k, v in module.named_buffers()})
    for name in sorted(state):
        t = state[name]
        if not t.is_floating_point() or name.endswith(("running_mean", "running_var")):
            continue
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        if name.endswith("gauss_B"):
            v = r
        elif name.endswith("logit_scale"):
            v = torch.full(t.shape, 2.6592)
        elif t.dim() >= 2:
            fan_in = t[0].numel() if leaf != "proj" and leaf != "text_projection" else t.shape[0]
            v = r * (1.0 / max(fan_in, 1) ** 0.5)
        elif leaf == "weight":      # norm scales
            v = 1.0 + 0.1 * r
        else:                       # biases, class / positional embeddings (1-D)
            v = 0.1 * r
        t.copy_(v.to(t.dtype))

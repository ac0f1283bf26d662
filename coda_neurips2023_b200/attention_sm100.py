"""Fused multi-head attention on tcgen05 tensor cores (include/coda_attention.h).

forward : one launch per attention call; S = Q K^T, online softmax and O = P V are
          tiled through TMEM, the (Lq x Lk) probabilities never reach HBM.
backward: two tcgen05 kernels (dQ; dK / dV) that recompute P tile by tile from the saved
          log-sum-exp (csrc/attention_bwd_sm100.cu); nothing of size Lq x Lk is stored.
`_math` below is the plain-tensor formulation used by the tests as the numerics reference and by
the causal CLIP text tower, which runs once at model construction (never inside the step).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_of


def _math(q, k, v, nhead, dropout_p, training, causal, keep=None):
    """softmax(q k^T / sqrt(hd)) v per head in plain tensor ops: numerics reference in tests and the
    init-time CLIP text tower.  `keep` is an explicit (B*H, Lq, Lk) dropout keep-mask."""
    lq, b, e = q.shape
    lk = k.shape[0]
    hd = e // nhead
    qh = (q * (float(hd) ** -0.5)).reshape(lq, b * nhead, hd).transpose(0, 1)
    kh = k.reshape(lk, b * nhead, hd).transpose(0, 1)
    vh = v.reshape(lk, b * nhead, hd).transpose(0, 1)
    s = torch.bmm(qh, kh.transpose(1, 2))
    if causal:
        s = s + torch.full((lq, lk), float("-inf"), device=s.device, dtype=s.dtype).triu_(1)
    p = torch.softmax(s, dim=-1)
    if keep is not None:
        p = p * (keep if keep.dtype == p.dtype else keep.to(p.dtype) * (1.0 / (1.0 - float(np.float32(dropout_p)))))
    elif training and dropout_p > 0.0:
        p = torch.nn.functional.dropout(p, dropout_p)
    return torch.bmm(p, vh).transpose(0, 1).reshape(lq, b, e)


def kernel_available() -> bool:
    return hasattr(lib(), "coda_attention_fwd")


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, nhead, dropout_p, salt):
        from . import attention_launch

        out, lse = attention_launch.forward(q, k, v, nhead, dropout_p, salt)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.nhead, ctx.dropout_p, ctx.salt = nhead, dropout_p, salt
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import attention_launch

        q, k, v, out, lse = ctx.saved_tensors
        # fused tcgen05 backward (dQ kernel + dK/dV kernel); P is recomputed tile by tile, never stored
        dq, dk, dv = attention_launch.backward(q, k, v, out, dout, lse, ctx.nhead, ctx.dropout_p, ctx.salt)
        return dq, dk, dv, None, None, None


def attention(q, k, v, nhead, dropout_p=0.0, training=False, causal=False):
    """q (Lq, B, E), k / v (Lk, B, E) -> (Lq, B, E); see ops.attention."""
    if not q.is_cuda:
        raise RuntimeError("attention: CPU not supported")
    hd = q.shape[-1] // nhead
    if causal or q.dtype != torch.float32 or hd not in (64, 128):
        # CLIP text tower only (causal, fp16/fp32, runs ONCE at model construction, never inside the step)
        return _math(q, k, v, nhead, dropout_p, training, causal)
    from . import attention_launch

    p = float(dropout_p) if training else 0.0
    return _Attention.apply(q, k, v, nhead, p, attention_launch.next_salt() if p > 0.0 else 0)

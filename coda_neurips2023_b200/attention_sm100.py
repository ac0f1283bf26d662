"""Fused multi-head attention on tcgen05 tensor cores (include/coda_attention.h).

forward : one launch per attention call; S = Q K^T, online softmax and O = P V are
          tiled through TMEM, the (Lq x Lk) probabilities never reach HBM.
backward: (this round) re-derives P with cuBLAS batched GEMMs under autograd; the
          tcgen05 backward kernel is the next step.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import check, lib, ptr, stream_of


def _math(q, k, v, nhead, dropout_p, training, causal):
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
    if training and dropout_p > 0.0:
        p = torch.nn.functional.dropout(p, dropout_p)
    return torch.bmm(p, vh).transpose(0, 1).reshape(lq, b, e)


def kernel_available() -> bool:
    return hasattr(lib(), "coda_attention_fwd")


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, nhead, causal):
        from . import attention_launch

        out = attention_launch.forward(q, k, v, nhead, causal)
        ctx.save_for_backward(q, k, v)
        ctx.nhead, ctx.causal = nhead, causal
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v = ctx.saved_tensors
        with torch.enable_grad():
            qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
            out = _math(qq, kk, vv, ctx.nhead, 0.0, False, ctx.causal)
        dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), dout)
        return dq, dk, dv, None, None


def attention(q, k, v, nhead, dropout_p=0.0, training=False, causal=False):
    """q (Lq, B, E), k / v (Lk, B, E) -> (Lq, B, E); see ops.attention."""
    if not q.is_cuda:
        raise RuntimeError("attention: CPU not supported")
    if training and dropout_p > 0.0 or not kernel_available():
        # attention-probability dropout (p = 0.1 in training) is applied on the materialised
        # probabilities until the in-kernel Philox mask lands; eval / parity runs use the kernel
        return _math(q, k, v, nhead, dropout_p, training, causal)
    return _Attention.apply(q, k, v, nhead, causal)

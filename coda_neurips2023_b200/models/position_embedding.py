"""Coordinate positional encodings (mirror of reference models/position_embedding.py).

The Fourier variant -- the one 3DETR / CoDA use -- is a single fused kernel
(ops.fourier_pos_embed).  The sine variant is dead on the path and not built.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops


class PositionEmbeddingCoordsSine(nn.Module):
    def __init__(self, temperature=10000, normalize=False, scale=None, pos_type="fourier", d_pos=None,
                 d_in=3, gauss_scale=1.0):
        super().__init__()
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if scale is None:
            scale = 2 * math.pi
        assert pos_type in ["sine", "fourier"]
        self.pos_type = pos_type
        self.scale = scale
        if pos_type == "fourier":
            assert d_pos is not None and d_pos % 2 == 0
            # gaussian projection d_in -> d_pos/2; a registered buffer, saved in checkpoints
            B = torch.empty((d_in, d_pos // 2)).normal_()
            B *= gauss_scale
            self.register_buffer("gauss_B", B)
            self.d_pos = d_pos

    def get_sine_embeddings(self, xyz, num_channels, input_range):
        raise NotImplementedError("pos_type='sine' is never selected on the CoDA path (every model builds "
                                  "PositionEmbeddingCoordsSine(pos_type='fourier')); only the Fourier kernel exists")

    def get_fourier_embeddings(self, xyz, num_channels=None, input_range=None):
        """xyz (B, N, 3) -> (B, num_channels, N); reference position_embedding.py:89-118."""
        if num_channels is None:
            num_channels = self.gauss_B.shape[1] * 2
        assert num_channels > 0 and num_channels % 2 == 0
        d_in, max_d_out = self.gauss_B.shape
        d_out = num_channels // 2
        assert d_out <= max_d_out and d_in == xyz.shape[-1] == 3
        rng = input_range if self.normalize else None
        return ops.fourier_pos_embed(xyz, self.gauss_B, d_out, rng)

    def forward(self, xyz, num_channels=None, input_range=None):
        assert isinstance(xyz, torch.Tensor) and xyz.ndim == 3
        with torch.no_grad():
            if self.pos_type == "sine":
                return self.get_sine_embeddings(xyz, num_channels, input_range)
            return self.get_fourier_embeddings(xyz, num_channels, input_range)

    def extra_repr(self):
        st = f"type={self.pos_type}, scale={self.scale}, normalize={self.normalize}"
        if hasattr(self, "gauss_B"):
            st += f", gaussB={self.gauss_B.shape}, gaussBsum={self.gauss_B.sum().item()}"
        return st

"""Coordinate positional encodings (mirror of reference models/position_embedding.py).

The Fourier variant -- the one 3DETR uses -- is a single fused kernel
(ops.fourier_pos_embed); the sine variant is kept as tensor code.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops
from ..utils.pc_util import shift_scale_points


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
        xyz = xyz.clone()
        if self.normalize:
            xyz = shift_scale_points(xyz, src_range=input_range)
        ndim = num_channels // xyz.shape[2]
        if ndim % 2 != 0:
            ndim -= 1
        rems = num_channels - (ndim * xyz.shape[2])  # remainder goes to the first dims, two at a time
        assert ndim % 2 == 0
        embeds, prev_dim, dim_t = [], 0, None
        for d in range(xyz.shape[2]):
            cdim = ndim
            if rems > 0:
                cdim += 2
                rems -= 2
            if cdim != prev_dim:
                dim_t = torch.arange(cdim, dtype=torch.float32, device=xyz.device)
                dim_t = self.temperature ** (2 * (dim_t // 2) / cdim)
            raw_pos = xyz[:, :, d]
            if self.scale:
                raw_pos *= self.scale
            pos = raw_pos[:, :, None] / dim_t
            pos = torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=3).flatten(2)
            embeds.append(pos)
            prev_dim = cdim
        return torch.cat(embeds, dim=2).permute(0, 2, 1)

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

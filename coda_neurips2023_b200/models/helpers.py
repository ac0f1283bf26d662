"""Building blocks shared by the 3DETR modules (mirror of reference models/helpers.py).

`NORM_DICT["ln"]` is the warp-per-row LayerNorm kernel (same parameters as
nn.LayerNorm); everything else is plain nn.Modules whose GEMMs run in cuBLAS.
"""
from __future__ import annotations

import copy
from functools import partial

import torch.nn as nn

from ..ops import LayerNorm


class BatchNormDim1Swap(nn.BatchNorm1d):
    """BatchNorm1d for sequence-first (HW, N, C) tensors (reference helpers.py:8-24)."""

    def forward(self, x):
        return super().forward(x.permute(1, 2, 0)).permute(2, 0, 1)


NORM_DICT = {
    "bn": BatchNormDim1Swap,
    "bn1d": nn.BatchNorm1d,
    "id": nn.Identity,
    "ln": LayerNorm,
}

ACTIVATION_DICT = {
    "relu": nn.ReLU,
    "gelu": nn.GELU,
    "leakyrelu": partial(nn.LeakyReLU, negative_slope=0.1),
}

WEIGHT_INIT_DICT = {
    "xavier_uniform": nn.init.xavier_uniform_,
}


class GenericMLP(nn.Module):
    """[Conv1d|Linear -> norm -> act -> dropout] x len(hidden_dims) -> output layer.

    Same constructor arguments and `layers.{i}` parameter paths as reference
    models/helpers.py:45-112.
    """

    def __init__(self, input_dim, hidden_dims, output_dim, norm_fn_name=None, activation="relu",
                 use_conv=False, dropout=None, hidden_use_bias=False, output_use_bias=True,
                 output_use_activation=False, output_use_norm=False, weight_init_name=None):
        super().__init__()
        act = ACTIVATION_DICT[activation]
        norm = NORM_DICT[norm_fn_name] if norm_fn_name is not None else None
        if norm_fn_name == "ln" and use_conv:
            norm = lambda x: nn.GroupNorm(1, x)  # noqa: E731  LayerNorm over channels of a conv map
        if dropout is not None and not isinstance(dropout, list):
            dropout = [dropout for _ in range(len(hidden_dims))]

        def dense(i, o, bias):
            return nn.Conv1d(i, o, 1, bias=bias) if use_conv else nn.Linear(i, o, bias=bias)

        layers = []
        prev = input_dim
        for idx, width in enumerate(hidden_dims):
            layers.append(dense(prev, width, hidden_use_bias))
            if norm:
                layers.append(norm(width))
            layers.append(act())
            if dropout is not None:
                layers.append(nn.Dropout(p=dropout[idx]))
            prev = width
        layers.append(dense(prev, output_dim, output_use_bias))
        if output_use_norm:
            layers.append(norm(output_dim))
        if output_use_activation:
            layers.append(act())
        self.layers = nn.Sequential(*layers)
        if weight_init_name is not None:
            self.do_weight_init(weight_init_name)

    def do_weight_init(self, weight_init_name):
        func = WEIGHT_INIT_DICT[weight_init_name]
        for _, param in self.named_parameters():
            if param.dim() > 1:  # skips the norm layers
                func(param)

    def forward(self, x):
        return self.layers(x)


def get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])

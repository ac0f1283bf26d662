"""3DETR transformer encoder / decoder (mirror of the LIVE classes of reference
models/transformer.py: TransformerEncoder, TransformerEncoderLayer,
MaskedTransformerEncoder, TransformerDecoder, TransformerDecoderLayer).

Differences from the reference are implementation-only: attention runs in the
fused tcgen05 kernel (no (B*H, L, L) probability tensor is written, no head
averaging), LayerNorm is the warp-per-row kernel, and the K/V input projections
of one packed weight are issued as single GEMMs.  Parameter names and shapes
match `nn.MultiheadAttention` / the reference layers so checkpoints interchange.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import Tensor, nn

from .. import ops
from .helpers import ACTIVATION_DICT, NORM_DICT, WEIGHT_INIT_DICT, get_clones


class MultiheadAttention(nn.Module):
    """Drop-in for `nn.MultiheadAttention(embed_dim, num_heads, dropout)` as the
    reference uses it (sequence-first, packed in_proj, bias, no kdim/vdim).

    forward(query, key, value, attn_mask=None, key_padding_mask=None) -> (out, None);
    attention weights are not returned (the reference discards them everywhere
    except under `return_attn_weights`, which no shipped configuration sets).
    """

    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = ops.Linear(embed_dim, embed_dim, bias=True)
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, attn_mask: Optional[Tensor] = None,
                key_padding_mask: Optional[Tensor] = None, need_weights: bool = False):
        if attn_mask is not None or key_padding_mask is not None:
            raise NotImplementedError("masked attention (enc_type=masked) is not on the B200 hot path yet")
        e = self.embed_dim
        w, b = self.in_proj_weight, self.in_proj_bias
        if query is key and key is value:      # encoder self-attention: one packed GEMM
            q, k, v = ops.linear(query, w, b).split(e, dim=-1)
        elif query is key:                      # decoder self-attention: q = k = tgt + pos, v = tgt
            q, k = ops.linear(query, w[: 2 * e], b[: 2 * e]).split(e, dim=-1)
            v = ops.linear(value, w[2 * e:], b[2 * e:])
        else:                                   # cross-attention
            q = ops.linear(query, w[:e], b[:e])
            k = ops.linear(key, w[e: 2 * e], b[e: 2 * e])
            v = ops.linear(value, w[2 * e:], b[2 * e:])
        out = ops.attention(q, k, v, self.num_heads, self.dropout, self.training)
        return self.out_proj(out), None


class TransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, norm=None, weight_init_name="xavier_uniform"):
        super().__init__()
        self.layers = get_clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self._reset_parameters(weight_init_name)

    def _reset_parameters(self, weight_init_name):
        func = WEIGHT_INIT_DICT[weight_init_name]
        for p in self.parameters():
            if p.dim() > 1:
                func(p)

    def forward(self, src, mask: Optional[Tensor] = None, src_key_padding_mask: Optional[Tensor] = None,
                pos: Optional[Tensor] = None, xyz: Optional[Tensor] = None, transpose_swap: Optional[bool] = False):
        """src (L, B, C) -> (xyz, output (L, B, C), None)   [reference transformer.py:35-74]"""
        if transpose_swap:
            bs, c, h, w = src.shape
            src = src.flatten(2).permute(2, 0, 1)
            if pos is not None:
                pos = pos.flatten(2).permute(2, 0, 1)
        output = src
        if mask is not None:
            raise NotImplementedError("attention masks are not supported by the fused attention kernel")
        for layer in self.layers:
            output = layer(output, src_mask=None, src_key_padding_mask=src_key_padding_mask, pos=pos)
        if self.norm is not None:
            output = self.norm(output)
        if transpose_swap:
            output = output.permute(1, 2, 0).view(bs, c, h, w).contiguous()
        return xyz, output, None


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead=4, dim_feedforward=128, dropout=0.1, dropout_attn=None,
                 activation="relu", normalize_before=True, norm_name="ln", use_ffn=True, ffn_use_bias=True):
        super().__init__()
        if dropout_attn is None:
            dropout_attn = dropout
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout_attn)
        self.use_ffn = use_ffn
        if self.use_ffn:
            self.linear1 = ops.Linear(d_model, dim_feedforward, bias=ffn_use_bias)
            self.dropout = nn.Dropout(dropout, inplace=False)
            self.linear2 = ops.Linear(dim_feedforward, d_model, bias=ffn_use_bias)
            self.norm2 = NORM_DICT[norm_name](d_model)
            self.dropout2 = nn.Dropout(dropout, inplace=False)
        self.norm1 = NORM_DICT[norm_name](d_model)
        self.dropout1 = nn.Dropout(dropout, inplace=False)
        self.activation = ACTIVATION_DICT[activation]()
        self.normalize_before = normalize_before
        self.nhead = nhead

    @staticmethod
    def with_pos_embed(tensor, pos: Optional[Tensor]):
        return tensor if pos is None else tensor + pos

    def forward_pre(self, src, src_mask=None, src_key_padding_mask=None, pos=None, return_attn_weights=False):
        # reference transformer.py:461-479
        src2 = self.norm1(src)
        if pos is None:
            attn = self.self_attn(src2, src2, src2, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)[0]
        else:
            qk = src2 + pos
            attn = self.self_attn(qk, qk, src2, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)[0]
        src = ops.dropout_add(attn, src, self.dropout1.p, self.training)
        if self.use_ffn:
            src2 = self.norm2(src)
            src2 = self.linear2(ops.dropout(_ffn_hidden(self, src2), self.dropout.p, self.training))
            src = ops.dropout_add(src2, src, self.dropout2.p, self.training)
        if return_attn_weights:
            return src, None
        return src

    def forward_post(self, src, src_mask=None, src_key_padding_mask=None, pos=None):
        qk = self.with_pos_embed(src, pos)
        src2 = self.self_attn(qk, qk, src, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)[0]
        src = self.norm1(ops.dropout_add(src2, src, self.dropout1.p, self.training))
        if self.use_ffn:
            src2 = self.linear2(ops.dropout(_ffn_hidden(self, src), self.dropout.p, self.training))
            src = self.norm2(ops.dropout_add(src2, src, self.dropout2.p, self.training))
        return src

    def forward(self, src, src_mask=None, src_key_padding_mask=None, pos=None, return_attn_weights=False):
        if self.normalize_before:
            return self.forward_pre(src, src_mask, src_key_padding_mask, pos, return_attn_weights)
        return self.forward_post(src, src_mask, src_key_padding_mask, pos)

    def extra_repr(self):
        return f"attn_dr={self.self_attn.dropout}"


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm_fn_name="ln", return_intermediate=False,
                 weight_init_name="xavier_uniform"):
        super().__init__()
        self.layers = get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = None
        if norm_fn_name is not None:
            self.norm = NORM_DICT[norm_fn_name](self.layers[0].linear2.out_features)
        self.return_intermediate = return_intermediate
        self._reset_parameters(weight_init_name)

    def _reset_parameters(self, weight_init_name):
        func = WEIGHT_INIT_DICT[weight_init_name]
        for p in self.parameters():
            if p.dim() > 1:
                func(p)

    def forward(self, tgt, memory, image_features_clip=None, text_features_clip=None, tgt_mask=None,
                memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None,
                pos: Optional[Tensor] = None, query_pos: Optional[Tensor] = None,
                transpose_swap: Optional[bool] = False, return_attn_weights: Optional[bool] = False):
        """tgt (Q, B, C), memory (L, B, C) -> (stack of per-layer normed outputs (nl, Q, B, C), attns)
        [reference transformer.py:97-143]"""
        if transpose_swap:
            bs, c, h, w = memory.shape
            memory = memory.flatten(2).permute(2, 0, 1)
            if pos is not None:
                pos = pos.flatten(2).permute(2, 0, 1)
        if return_attn_weights:
            raise NotImplementedError("attention weights are not materialised by the fused kernel")
        # key of every cross-attention is memory + pos: formed once, not once per layer
        mem_key = memory if pos is None else memory + pos
        output = tgt
        intermediate = []
        for layer in self.layers:
            output, _ = layer(output, memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                              tgt_key_padding_mask=tgt_key_padding_mask,
                              memory_key_padding_mask=memory_key_padding_mask, pos=pos, query_pos=query_pos,
                              memory_key=mem_key)
            if self.return_intermediate:
                intermediate.append(self.norm(output))
        if self.norm is not None:
            output = self.norm(output)
            if self.return_intermediate:
                intermediate.pop()
                intermediate.append(output)
        if self.return_intermediate:
            return torch.stack(intermediate), []
        return output, []


def _ffn_hidden(layer, x):
    """activation(linear1(x)); with a ReLU activation the rectifier runs in the GEMM epilogue (one kernel less
    forward and backward)."""
    if isinstance(layer.activation, nn.ReLU):
        return ops.linear(x, layer.linear1.weight, layer.linear1.bias, relu=True)
    return layer.activation(layer.linear1(x))


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model, nhead=4, dim_feedforward=256, dropout=0.1, dropout_attn=None,
                 activation="relu", normalize_before=True, norm_fn_name="ln"):
        super().__init__()
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm1 = NORM_DICT[norm_fn_name](d_model)
        self.norm2 = NORM_DICT[norm_fn_name](d_model)
        self.norm3 = NORM_DICT[norm_fn_name](d_model)
        self.dropout1 = nn.Dropout(dropout, inplace=False)
        self.dropout2 = nn.Dropout(dropout, inplace=False)
        self.dropout3 = nn.Dropout(dropout, inplace=False)
        self.linear1 = ops.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout, inplace=False)
        self.linear2 = ops.Linear(dim_feedforward, d_model)
        self.activation = ACTIVATION_DICT[activation]()
        self.normalize_before = normalize_before

    @staticmethod
    def with_pos_embed(tensor, pos: Optional[Tensor]):
        return tensor if pos is None else tensor + pos

    def forward_pre(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                    memory_key_padding_mask=None, pos=None, query_pos=None, return_attn_weights=False,
                    memory_key=None):
        # reference transformer.py:556-580
        if memory_key is None:
            memory_key = self.with_pos_embed(memory, pos)
        tgt2 = self.norm1(tgt)
        qk = self.with_pos_embed(tgt2, query_pos)
        tgt2 = self.self_attn(qk, qk, tgt2, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0]
        tgt = ops.dropout_add(tgt2, tgt, self.dropout1.p, self.training)
        tgt2 = self.norm2(tgt)
        tgt2 = self.multihead_attn(self.with_pos_embed(tgt2, query_pos), memory_key, memory,
                                   attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask)[0]
        tgt = ops.dropout_add(tgt2, tgt, self.dropout2.p, self.training)
        tgt2 = self.norm3(tgt)
        tgt2 = self.linear2(ops.dropout(_ffn_hidden(self, tgt2), self.dropout.p, self.training))
        tgt = ops.dropout_add(tgt2, tgt, self.dropout3.p, self.training)
        return tgt, None

    def forward_post(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                     memory_key_padding_mask=None, pos=None, query_pos=None, return_attn_weights=False,
                     memory_key=None):
        if memory_key is None:
            memory_key = self.with_pos_embed(memory, pos)
        qk = self.with_pos_embed(tgt, query_pos)
        tgt2 = self.self_attn(qk, qk, tgt, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0]
        tgt = self.norm1(ops.dropout_add(tgt2, tgt, self.dropout1.p, self.training))
        tgt2 = self.multihead_attn(self.with_pos_embed(tgt, query_pos), memory_key, memory,
                                   attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask)[0]
        tgt = self.norm2(ops.dropout_add(tgt2, tgt, self.dropout2.p, self.training))
        tgt2 = self.linear2(ops.dropout(_ffn_hidden(self, tgt), self.dropout.p, self.training))
        tgt = self.norm3(ops.dropout_add(tgt2, tgt, self.dropout3.p, self.training))
        return tgt, None

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos=None, query_pos=None, return_attn_weights=False,
                memory_key=None):
        fn = self.forward_pre if self.normalize_before else self.forward_post
        return fn(tgt, memory, tgt_mask, memory_mask, tgt_key_padding_mask, memory_key_padding_mask, pos,
                  query_pos, return_attn_weights, memory_key)

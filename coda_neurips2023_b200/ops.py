"""torch-facing wrappers (autograd where needed) around the C-ABI kernels of
include/coda_detr.h and include/coda_attention.h.  Every function here launches
hand-written sm_100a code on the current stream; none has a CPU or PyTorch
fallback -- a CPU tensor raises.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import check, lib, ptr, stream_of

_i = ctypes.c_int
_ll = ctypes.c_longlong
_f = ctypes.c_float


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CPU not supported (coda_b200 kernels are CUDA-only)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------- LayerNorm
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        _need_cuda(x, "layer_norm")
        xc = _f32c(x)
        c = xc.shape[-1]
        rows = xc.numel() // c
        y = torch.empty_like(xc)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            st = lib().coda_layer_norm_fwd(_ll(rows), _i(c), _f(eps), ptr(xc), ptr(weight), ptr(bias), ptr(y),
                                           ptr(mean), ptr(rstd), stream_of(x))
        check(st, "layer_norm_fwd")
        ctx.save_for_backward(xc, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, weight, mean, rstd = ctx.saved_tensors
        dyc = _f32c(dy)
        c = xc.shape[-1]
        rows = xc.numel() // c
        dx = torch.empty_like(xc)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(weight)
        nscratch = lib().coda_layer_norm_bwd_scratch(_ll(rows), _i(c))
        partial = torch.empty(max(int(nscratch), 1), dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            st = lib().coda_layer_norm_bwd(_ll(rows), _i(c), ptr(dyc), ptr(xc), ptr(weight), ptr(mean), ptr(rstd),
                                           ptr(dx), ptr(dgamma), ptr(dbeta), ptr(partial), stream_of(xc))
        check(st, "layer_norm_bwd")
        return dx, dgamma, dbeta, None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last dim (multiple of 128, <= 1024); fp32."""
    return _LayerNorm.apply(x, weight, bias, float(eps))


@torch.no_grad()
def layer_norm_half(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Forward-only LayerNorm on fp16 activations with fp32 statistics (frozen CLIP towers)."""
    _need_cuda(x, "layer_norm_half")
    assert x.dtype == torch.float16
    xc = x.contiguous()
    c = xc.shape[-1]
    y = torch.empty_like(xc)
    with torch.cuda.device(x.device):
        st = lib().coda_layer_norm_fwd_half(_ll(xc.numel() // c), _i(c), _f(eps), ptr(xc), ptr(_f32c(weight)),
                                            ptr(_f32c(bias)), ptr(y), stream_of(x))
    check(st, "layer_norm_fwd_half")
    return y


class LayerNorm(torch.nn.LayerNorm):
    """nn.LayerNorm with the same parameters / state-dict keys, running the
    warp-per-row kernel (reference NORM_DICT["ln"], models/helpers.py:27-32)."""

    def forward(self, x):
        return layer_norm(x, self.weight, self.bias, self.eps)


# --------------------------------------------------------------------------- softmax
def softmax_rows(x: torch.Tensor, log: bool = False) -> torch.Tensor:
    """softmax / log-softmax over the last dimension (no autograd)."""
    _need_cuda(x, "softmax_rows")
    xc = _f32c(x.detach())
    c = xc.shape[-1]
    y = torch.empty_like(xc)
    with torch.cuda.device(x.device):
        st = lib().coda_softmax_rows(_ll(xc.numel() // c), _i(c), _i(1 if log else 0), ptr(xc), ptr(y), stream_of(x))
    check(st, "softmax_rows")
    return y


# --------------------------------------------------------------------------- Fourier pos-enc
@torch.no_grad()
def fourier_pos_embed(xyz: torch.Tensor, gauss_B: torch.Tensor, d_out: int, input_range=None) -> torch.Tensor:
    """xyz (B, N, 3) -> (B, 2*d_out, N) = [sin | cos](2 pi x_hat @ gauss_B[:, :d_out]);
    x_hat is xyz mapped to [0, 1] by `input_range` = [min (B, 3), max (B, 3)] if given."""
    _need_cuda(xyz, "fourier_pos_embed")
    xc = _f32c(xyz)
    gb = _f32c(gauss_B)
    b, n, _ = xc.shape
    out = torch.empty((b, 2 * d_out, n), dtype=torch.float32, device=xyz.device)
    rmin = rmax = None
    if input_range is not None:
        rmin, rmax = _f32c(input_range[0]), _f32c(input_range[1])
    with torch.cuda.device(xyz.device):
        st = lib().coda_fourier_pos_embed(_i(b), _i(n), _i(d_out), _i(gb.shape[1]), _i(0 if rmin is None else 1),
                                          ptr(xc), ptr(rmin), ptr(rmax), ptr(gb), ptr(out), stream_of(xyz))
    check(st, "fourier_pos_embed")
    return out


# --------------------------------------------------------------------------- GIoU / matcher
@torch.no_grad()
def giou3d(corners1: torch.Tensor, corners2: torch.Tensor, nums_k2: torch.Tensor, rotated,
           rot_k2_limit: int | None = None) -> torch.Tensor:
    """(B, K1, 8, 3), (B, K2, 8, 3), (B,) -> generalised IoU (B, K1, K2).
    `rotated` is a bool, or a 1-element device tensor (read by the kernel: no host sync)."""
    _need_cuda(corners1, "giou3d")
    c1, c2 = _f32c(corners1), _f32c(corners2)
    b, k1, k2 = c1.shape[0], c1.shape[1], c2.shape[1]
    nk = nums_k2.to(device=c1.device, dtype=torch.int32).contiguous()
    out = torch.empty((b, k1, k2), dtype=torch.float32, device=c1.device)
    lim = k2 if rot_k2_limit is None else int(rot_k2_limit)
    rdev = None
    if isinstance(rotated, torch.Tensor):
        rdev = rotated.to(device=c1.device, dtype=torch.int32).reshape(1).contiguous()
        rotated = False
    with torch.cuda.device(c1.device):
        st = lib().coda_giou3d(_i(b), _i(k1), _i(k2), _i(1 if rotated else 0), ptr(rdev), _i(lim), ptr(c1), ptr(c2),
                               ptr(nk), ptr(out), stream_of(c1))
    check(st, "giou3d")
    return out


@torch.no_grad()
def hungarian(cost: torch.Tensor, nactual: torch.Tensor):
    """cost (B, nprop, ngt) fp32, nactual (B,) -> (per_prop_gt_inds int64 (B, nprop),
    proposal_matched_mask fp32 (B, nprop)); same assignment as scipy's
    linear_sum_assignment on cost[b, :, :nactual[b]]."""
    _need_cuda(cost, "hungarian")
    cc = _f32c(cost)
    b, nprop, ngt = cc.shape
    na = nactual.to(device=cc.device, dtype=torch.int32).contiguous()
    inds = torch.empty((b, nprop), dtype=torch.int64, device=cc.device)
    mask = torch.empty((b, nprop), dtype=torch.float32, device=cc.device)
    with torch.cuda.device(cc.device):
        st = lib().coda_hungarian(_i(b), _i(nprop), _i(ngt), ptr(cc), ptr(na), ptr(inds), ptr(mask), stream_of(cc))
    check(st, "hungarian")
    return inds, mask


# --------------------------------------------------------------------------- attention
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, nhead: int, dropout_p: float = 0.0,
              training: bool = False, causal: bool = False) -> torch.Tensor:
    """Multi-head scaled-dot-product attention on projected, sequence-first tensors.

    q (Lq, B, E), k / v (Lk, B, E) -> (Lq, B, E).  The scale 1/sqrt(E/nhead) is applied to
    q; the forward kernel never materialises the probabilities in HBM.
    """
    _need_cuda(q, "attention")
    from . import attention_sm100  # tcgen05 kernels (include/coda_attention.h)

    return attention_sm100.attention(q, k, v, nhead, dropout_p, training, causal)


# --------------------------------------------------------------------------- CLIP crops
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@torch.no_grad()
def crop_resize_normalize(images: torch.Tensor, scene: torch.Tensor, boxes: torch.Tensor, valid: torch.Tensor,
                          res: int, dtype=torch.float16, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
    """images (B, H, W, 3) uint8, scene (N,) int32, boxes (N, 4) int32 [xmin, ymin, xmax, ymax],
    valid (N,) bool -> (N, 3, res, res) CLIP-normalised crops (white-padded to square, antialiased
    bicubic resize with torchvision's uint8 semantics)."""
    _need_cuda(images, "crop_resize_normalize")
    if images.dtype != torch.uint8:
        raise RuntimeError("images must be uint8 (HWC)")
    img = images.contiguous()
    nimg, h, w, _ = img.shape
    n = boxes.shape[0]
    sc = scene.to(torch.int32).contiguous()
    bx = boxes.to(torch.int32).contiguous()
    vd = valid.to(torch.uint8).contiguous()
    if dtype not in (torch.float16, torch.float32):
        raise RuntimeError("output dtype must be float16 or float32")
    out = torch.empty((n, 3, res, res), dtype=dtype, device=img.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(img.device):
        st = lib().coda_crop_resize_normalize(_i(nimg), _i(h), _i(w), _i(n), _i(res), ptr(img), ptr(sc), ptr(bx),
                                              ptr(vd), m, s, _i(1 if dtype == torch.float16 else 0), ptr(out),
                                              stream_of(img))
    check(st, "crop_resize_normalize")
    return out


# --------------------------------------------------------------------------- tcgen05 GEMM
# bf16 planes per fp32 operand.  3 -> six cross products, all 24 mantissa bits: fp32-class accuracy, which the
# 1e-4 parity bar needs through 13 layers (2 planes / 3 products measured 1.4e-4 on the class logits).
DEFAULT_NSPLIT = 3


def _pad64(k: int) -> int:
    return (k + 63) // 64 * 64


def pack_split(x: torch.Tensor, rows: int, k: int, row_stride: int, k_stride: int, nsplit: int = DEFAULT_NSPLIT,
               scale: float = 1.0, batch: int = 1, batch_stride: int = 0) -> torch.Tensor:
    """fp32 matrix/matrices addressed as x[b*batch_stride + r*row_stride + c*k_stride] -> bf16 planes
    (nsplit, batch, rows, kpad), K zero-padded to a multiple of 64 (operand format of gemm_nt)."""
    _need_cuda(x, "pack_split")
    assert x.dtype == torch.float32
    kpad = _pad64(k)
    out = torch.empty((nsplit, batch, rows, kpad), dtype=torch.bfloat16, device=x.device)
    base = x.data_ptr()
    with torch.cuda.device(x.device):
        for b in range(batch):
            # planes of one batch entry are (batch * rows * kpad) apart: pack plane-by-plane views
            st = lib().coda_pack_split_bf16_strided(
                _ll(rows), _i(k), _i(kpad), _ll(row_stride), _ll(k_stride),
                ctypes.c_void_p(base + 4 * b * batch_stride), _f(scale), _i(nsplit),
                ctypes.c_void_p(out.data_ptr() + 2 * b * rows * kpad), _ll(batch * rows * kpad), stream_of(x))
            check(st, "pack_split_bf16")
    return out


def gemm_nt(a_planes: torch.Tensor, b_planes: torch.Tensor, m: int, n: int, bias=None, relu: bool = False,
            out: torch.Tensor | None = None, act: int | None = None, out_dtype=torch.float32,
            residual: torch.Tensor | None = None) -> torch.Tensor:
    """C[b] = A[b] @ B[b]^T (+ bias) from packed planes (nsplit, batch, rows, kpad); B may have batch 1
    (shared weights).  fp16 planes (nsplit == 1) select the fp16 tensor-core path.  Returns fp32 (batch, m, n)."""
    _need_cuda(a_planes, "gemm_nt")
    nsplit, batch, _, kpad = a_planes.shape
    assert b_planes.shape[0] == nsplit and b_planes.shape[3] == kpad and a_planes.dtype == b_planes.dtype
    is_fp16 = a_planes.dtype == torch.float16
    bb = b_planes.shape[1]
    assert bb in (1, batch)
    if out is None:
        out = torch.empty((batch, m, n), dtype=out_dtype, device=a_planes.device)
    if act is None:
        act = 1 if relu else 0
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.shape == (m, n) and residual.stride(1) == 1
    with torch.cuda.device(a_planes.device):
        st = lib().coda_gemm_nt_res(
            _i(nsplit), _i(1 if is_fp16 else 0), _i(batch), _i(m), _i(n), _i(kpad), ptr(a_planes),
            _ll(a_planes.stride(0)), _ll(a_planes.stride(1)), ptr(b_planes), _ll(b_planes.stride(0)),
            _ll(b_planes.stride(1) if bb > 1 else 0), ptr(bias), _i(act), _i(1 if out.dtype == torch.float16 else 0),
            ptr(residual), _ll(residual.stride(0) if residual is not None else 0),
            ptr(out), _ll(out.stride(1)), _ll(out.stride(0)), stream_of(a_planes))
    check(st, "gemm_nt")
    return out


def gemm_tn(a_planes: torch.Tensor, b_planes: torch.Tensor, m: int, n: int) -> torch.Tensor:
    """C (m, n) = sum_r A[r, :m]^T B[r, :n] from ROW-packed planes (nsplit, 1, rows, pad64(cols)):
    the weight-gradient form dW = dY^T X on MN-major tensor-core operands (no transposed copies)."""
    _need_cuda(a_planes, "gemm_tn")
    nsplit, _, mc, lda = a_planes.shape
    assert b_planes.shape[0] == nsplit and b_planes.shape[2] == mc
    ldb = b_planes.shape[3]
    out = torch.empty((m, n), dtype=torch.float32, device=a_planes.device)
    with torch.cuda.device(a_planes.device):
        st = lib().coda_gemm_tn(_i(nsplit), _i(mc), _i(m), _i(n), ptr(a_planes), _ll(a_planes.stride(0)), _i(lda),
                                ptr(b_planes), _ll(b_planes.stride(0)), _i(ldb), ptr(out), _ll(n), stream_of(a_planes))
    check(st, "gemm_tn")
    return out


BACKWARD_NSPLIT = 2
USE_TN_WGRAD = True  # weight gradients from the row-packed operands (MN-major MMA); False: transposed packs


# --------------------------------------------------------------------------- Linear on the tcgen05 GEMM
_WEIGHT_EPOCH = 0
_WEIGHT_CACHE: dict = {}


_ACT_CACHE: dict = {}   # packed planes of activations that several layers consume within one step


def invalidate_weight_cache() -> None:
    """Call after parameters were updated through storage the tensors' version counters do not see
    (the flat-buffer optimiser step of engine.TrainStep).  Also ends the per-step activation-pack cache."""
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1
    _WEIGHT_CACHE.clear()
    _ACT_CACHE.clear()


def _packed_rows(x: torch.Tensor, nsplit: int) -> torch.Tensor:
    """Row-packed planes of a 2-D activation, shared by every layer that consumes the SAME tensor in
    this step (the six prediction heads read one box-feature tensor, the eight decoder layers project
    the same encoder memory).  The entry keeps `x` alive, so its address cannot be recycled."""
    key = (x.data_ptr(), tuple(x.shape), x.stride(0), x._version, nsplit)
    hit = _ACT_CACHE.get(key)
    if hit is None:
        planes = pack_split(x, x.shape[0], x.shape[1], x.stride(0), 1, nsplit)
        if len(_ACT_CACHE) > 64:
            _ACT_CACHE.clear()
        _ACT_CACHE[key] = (planes, x)
        return planes
    return hit[0]


def _packed_weight(w: torch.Tensor, transposed: bool, nsplit: int) -> torch.Tensor:
    """Planes of W (N, K) as a B operand: rows = N, k = K; or of W^T (rows = K, k = N) when transposed."""
    key = (w.data_ptr(), tuple(w.shape), w._version, _WEIGHT_EPOCH, transposed, nsplit)
    hit = _WEIGHT_CACHE.get(key)
    if hit is None:
        n, k = w.shape
        wd = w.detach()
        planes = pack_split(wd, k, n, 1, k, nsplit) if transposed else pack_split(wd, n, k, k, 1, nsplit)
        if len(_WEIGHT_CACHE) > 512:
            _WEIGHT_CACHE.clear()
        # the entry keeps the weight's storage alive: a freed weight's address could otherwise be handed to a
        # new parameter of the same shape and hit this entry with stale planes
        hit = _WEIGHT_CACHE[key] = (planes, wd)
    return hit[0]


def colsum(x: torch.Tensor) -> torch.Tensor:
    """sum over the rows of a contiguous (rows, c) fp32 matrix (bias gradients); falls back to torch for channel
    counts the row kernels do not cover."""
    rows, c = x.shape
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and rows > 0
            and 4 <= c <= 1024 and c % 4 == 0 and 256 % (c // 4) == 0):
        return x.sum(dim=0)
    out = torch.empty(c, dtype=torch.float32, device=x.device)
    scratch = torch.empty(148 * 4 * 2 * c, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = lib().coda_rows_colsum(_ll(rows), _i(c), ptr(x), ptr(out), ptr(scratch), stream_of(x))
    check(st, "rows_colsum")
    return out


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, nsplit):
        m, k = x.shape
        n = weight.shape[0]
        xa = _packed_rows(x, nsplit)
        y = gemm_nt(xa, _packed_weight(weight, False, nsplit), m, n, bias=bias, relu=relu)[0]
        # the backward needs x only as a GEMM operand: keep its packed planes instead of the fp32 tensor
        keep_planes = USE_TN_WGRAD and weight.requires_grad
        ctx.save_for_backward(None if keep_planes else x, weight, y if relu else None, xa if keep_planes else None)
        ctx.xshape = (m, k)
        ctx.has_bias, ctx.relu, ctx.nsplit = bias is not None, relu, nsplit
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y, xa = ctx.saved_tensors
        # gradients are held to a looser bar than the forward (5e-3 vs 1e-4): two planes suffice
        nsplit = min(ctx.nsplit, BACKWARD_NSPLIT)
        if xa is not None:
            xa = xa[:nsplit]
        dy = dy.contiguous()
        if ctx.relu:
            dy = dy * (y > 0).to(dy.dtype)
        m, k = ctx.xshape
        n = weight.shape[0]
        dx = dw = db = None
        dya = None
        if ctx.needs_input_grad[0] or (ctx.needs_input_grad[1] and xa is not None):
            dya = pack_split(dy, m, n, n, 1, nsplit)   # row-packed dY: operand of both gradients
        if ctx.needs_input_grad[0]:
            # dX (m, k) = dY (m, n) @ W (n, k): B operand = W^T planes (rows k, contraction n)
            dx = gemm_nt(dya, _packed_weight(weight, True, nsplit), m, k)[0]
        if ctx.needs_input_grad[1]:
            if xa is not None:
                # dW (n, k) = sum_m dY[m, n] X[m, k]: contraction over the ROWS of both packed operands
                dw = gemm_tn(dya, xa, n, k)
            else:
                dyt = pack_split(dy, n, m, 1, n, nsplit)
                xt = pack_split(x, k, m, 1, x.stride(0), nsplit)
                dw = gemm_nt(dyt, xt, n, k)[0]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dx, dw, db, None, None


_BIAS32: dict = {}


def _bias_fp32(bias):
    """fp32 copy of a (frozen, fp16) bias, cached: the epilogue adds the bias in fp32.  The entry keeps the
    source tensor alive so that its address cannot be recycled for another parameter while the entry exists."""
    key = (bias.data_ptr(), tuple(bias.shape), bias._version)
    hit = _BIAS32.get(key)
    if hit is None:
        if len(_BIAS32) > 256:
            _BIAS32.clear()
        hit = _BIAS32[key] = (bias.detach().float(), bias)
    return hit[0]


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None, relu: bool = False, nsplit: int | None = None,
           quick_gelu: bool = False, residual: torch.Tensor | None = None):
    """y = x @ weight^T + bias over the last dim of x, on the tcgen05 GEMM (fp32 in / out, bf16
    split-operand accumulation); fp16 x / weight take the fp16 tensor-core path (inference only)."""
    _need_cuda(x, "linear")
    lead = x.shape[:-1]
    k = x.shape[-1]
    n = weight.shape[0]
    x2 = x.reshape(-1, k)
    if x2.dtype == torch.float16:
        assert weight.dtype == torch.float16 and k % 64 == 0, "fp16 path needs fp16 weights and K % 64 == 0"
        x2 = x2.contiguous()
        res2 = None
        if residual is not None:      # y = act(x W^T + b) + residual, added in the GEMM epilogue
            res2 = residual.reshape(-1, n)
            if res2.dtype != torch.float16 or res2.stride(1) != 1 or res2.stride(0) % 8 != 0 or n % 8 != 0:
                res2 = None
        y = gemm_nt(x2.view(1, 1, x2.shape[0], k), weight.detach().contiguous().view(1, 1, n, k), x2.shape[0], n,
                    bias=None if bias is None else _bias_fp32(bias), act=2 if quick_gelu else (1 if relu else 0),
                    out_dtype=torch.float16, residual=res2)[0]
        y = y.reshape(*lead, n)
        return y + residual if (residual is not None and res2 is None) else y
    if residual is not None:
        raise NotImplementedError("fused residual exists on the fp16 inference path only")
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w2 = weight.reshape(n, -1)
    y = _Linear.apply(x2.float(), w2, bias, relu, DEFAULT_NSPLIT if nsplit is None else nsplit)
    return y.reshape(*lead, n)


class Linear(torch.nn.Linear):
    """nn.Linear (same parameters / state-dict keys) running on the tcgen05 GEMM."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)

"""GPU tests of the warp-primitive kernels (include/coda_detr.h, coda_image.h) against
plain PyTorch fp32 references / scipy / torchvision, through the C-ABI wrappers."""
import numpy as np
import pytest
import torch

import giou_ref
import ref_crop
from coda_neurips2023_b200 import ops, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _lib(built_lib):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("rows,c", [(2048 * 8, 256), (256 * 8, 512), (50 * 256, 768), (37, 128), (1, 1024)])
def test_layer_norm_fwd_bwd(rows, c):
    torch.manual_seed(0)
    x = (torch.randn(rows, c, device="cuda") * 3 + 1).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(c, device="cuda")).requires_grad_(True)
    b = (0.1 * torch.randn(c, device="cuda")).requires_grad_(True)
    y = ops.layer_norm(x, w, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x.double(), (c,), w.double(), b.double(), 1e-5)
    torch.testing.assert_close(y, ref.float(), rtol=1e-5, atol=1e-5)
    g = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), g)
    rx, rw, rb = torch.autograd.grad(ref, (x, w, b), g.double())
    torch.testing.assert_close(gx, rx, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(gw, rw, rtol=1e-4, atol=1e-3 * max(1.0, rows ** 0.5 / 30))
    torch.testing.assert_close(gb, rb, rtol=1e-4, atol=1e-3 * max(1.0, rows ** 0.5 / 30))


@pytest.mark.parametrize("rows,c", [(50 * 77, 768), (13, 512), (1, 128)])
def test_layer_norm_half_matches_fp32_upcast(rows, c):
    # CLIP's LayerNorm (CLIP/clip/model.py:254-260) = fp32 layer_norm of the fp16 input, rounded to fp16
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(rows, c, generator=g) * 2 + 0.3).half().cuda()
    w = (torch.rand(c, generator=g) + 0.5).cuda()
    b = torch.randn(c, generator=g).cuda()
    y = ops.layer_norm_half(x, w, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x.float(), (c,), w, b, 1e-5).half()
    assert y.dtype == torch.float16 and y.shape == x.shape
    # same fp32 arithmetic up to summation order: at most one fp16 ulp apart
    diff = (y.float() - ref.float()).abs()
    assert (diff <= ref.float().abs() * 2 ** -10 + 1e-6).all()


def test_layer_norm_module_matches_nn_layernorm_3d():
    ln = ops.LayerNorm(256).cuda()
    ref = torch.nn.LayerNorm(256).cuda()
    x = torch.randn(300, 4, 256, device="cuda")
    torch.testing.assert_close(ln(x), ref(x), rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        ops.layer_norm(torch.randn(4, 256), ln.weight.cpu(), ln.bias.cpu())


@pytest.mark.parametrize("rows,c", [(4096, 2), (1000, 10), (512, 1201), (3, 46)])
def test_softmax_rows(rows, c):
    x = torch.randn(rows, c, device="cuda") * 5
    torch.testing.assert_close(ops.softmax_rows(x), torch.softmax(x, -1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ops.softmax_rows(x, log=True), torch.log_softmax(x, -1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("b,n,d_out,normalize", [(8, 2048, 256, True), (8, 256, 256, True), (2, 77, 64, False), (1, 1, 2, True)])
def test_fourier_pos_embed_matches_reference_formula(b, n, d_out, normalize):
    """reference models/position_embedding.py:89-118 with torch ops"""
    torch.manual_seed(1)
    xyz = torch.rand(b, n, 3, device="cuda") * 6 - 3
    gauss_B = torch.randn(3, max(d_out, 256), device="cuda")
    rng = [xyz.amin(1) - 0.1, xyz.amax(1) + 0.1] if normalize else None
    got = ops.fourier_pos_embed(xyz, gauss_B, d_out, rng)
    x = xyz.clone()
    if normalize:
        x = ((x - rng[0][:, None, :]) * 1.0) / (rng[1][:, None, :] - rng[0][:, None, :]) + 0.0
    x *= 2 * np.pi
    proj = torch.mm(x.view(-1, 3).double(), gauss_B[:, :d_out].double()).view(b, n, d_out)
    exp = torch.cat([proj.sin(), proj.cos()], dim=2).permute(0, 2, 1).float()
    # |proj| reaches ~50 rad: one fp32 ulp of the argument is 4e-6
    torch.testing.assert_close(got, exp, rtol=0, atol=3e-5)


def _boxes(b, k, seed, rotated=True):
    from coda_neurips2023_b200.utils.box_util import flip_axis_to_camera_tensor, get_3d_box_batch_tensor

    g = torch.Generator().manual_seed(seed)
    size = torch.rand(b, k, 3, generator=g) * 1.7 + 0.3
    ang = (torch.rand(b, k, generator=g) * 2 - 1) * np.pi if rotated else torch.zeros(b, k)
    cen = torch.rand(b, k, 3, generator=g) * torch.tensor([4.0, 4.0, 1.5]) + torch.tensor([-2.0, 1.0, -0.7])
    return get_3d_box_batch_tensor(size, ang, flip_axis_to_camera_tensor(cen))


@pytest.mark.parametrize("rotated,limit", [(True, None), (False, None), (True, 4)])
def test_giou3d_vs_restatement(rotated, limit):
    c1, c2 = _boxes(3, 40, 0, rotated), _boxes(3, 12, 1, rotated)
    c2[:, :3] = c1[:, :3] + 0.05  # a few heavily overlapping pairs
    nk = torch.tensor([12, 5, 0])
    got = ops.giou3d(c1.cuda(), c2.cuda(), nk.cuda(), rotated, limit).cpu()
    exp = giou_ref.giou3d_ref(c1, c2, nk, rotated, limit)
    torch.testing.assert_close(got, exp, rtol=1e-4, atol=2e-5)
    assert (got[1, :, 5:] == 0).all() and (got[2] == 0).all()
    # device-side flag == host flag
    flag = torch.tensor([1 if rotated else 0], device="cuda", dtype=torch.int32)
    assert torch.equal(ops.giou3d(c1.cuda(), c2.cuda(), nk.cuda(), flag, limit).cpu(), got)


def test_giou3d_identical_axis_aligned_boxes_is_one():
    # (for ROTATED identical boxes the reference's strict-inequality clip degenerates, and so do we)
    c = _boxes(1, 6, 3, False)
    g = ops.giou3d(c.cuda(), c.cuda(), torch.tensor([6]).cuda(), False).cpu()
    torch.testing.assert_close(torch.diagonal(g[0]), torch.ones(6), rtol=0, atol=2e-4)


@pytest.mark.parametrize("b,nprop,ngt,seed", [(8, 256, 64, 0), (4, 128, 64, 1), (3, 16, 64, 2), (2, 300, 7, 3)])
def test_hungarian_equals_scipy(b, nprop, ngt, seed):
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(seed)
    cost = rng.standard_normal((b, nprop, ngt)).astype(np.float32)
    nact = rng.integers(0, ngt + 1, size=b).astype(np.int32)
    nact[0] = min(ngt, 20)
    if b > 2:
        nact[2] = 0
    inds, mask = ops.hungarian(torch.from_numpy(cost).cuda(), torch.from_numpy(nact).cuda())
    inds, mask = inds.cpu().numpy(), mask.cpu().numpy()
    for i in range(b):
        e_inds = np.zeros(nprop, np.int64)
        e_mask = np.zeros(nprop, np.float32)
        if nact[i] > 0:
            r, c = linear_sum_assignment(cost[i, :, : nact[i]])
            e_inds[r] = c
            e_mask[r] = 1
        assert np.array_equal(mask[i], e_mask), f"scene {i}: matched set differs"
        assert np.array_equal(inds[i], e_inds), f"scene {i}: assignment differs"


def test_hungarian_ties_follow_scipy():
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(7)
    cost = rng.integers(0, 3, size=(6, 64, 64)).astype(np.float32)  # small integers: many exact ties
    cost[0] = 1.0                                                  # constant matrix
    nact = np.array([10, 64, 33, 1, 17, 5], np.int32)
    inds, mask = ops.hungarian(torch.from_numpy(cost).cuda(), torch.from_numpy(nact).cuda())
    for i in range(6):
        r, c = linear_sum_assignment(cost[i, :, : nact[i]])
        e = np.zeros(64, np.int64); e[r] = c
        m = np.zeros(64, np.float32); m[r] = 1
        assert np.array_equal(mask[i].cpu().numpy(), m) and np.array_equal(inds[i].cpu().numpy(), e)


def test_crop_resize_normalize_vs_torchvision_sequence():
    """reference per-box sequence (model_3detr.py:1034-1088) with torchvision on the GPU"""
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, size=(2, 54, 73, 3), dtype=np.uint8)  # low-frequency content, upsampled
    imgs = torch.from_numpy(base).cuda().permute(0, 3, 1, 2).float()
    imgs = torch.nn.functional.interpolate(imgs, size=(531, 730), mode="bilinear").round().clamp(0, 255)
    imgs = imgs.to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    boxes = torch.tensor([[100, 50, 400, 500], [10, 10, 60, 40], [0, 0, 730, 531], [300, 200, 524, 424],
                          [5, 5, 10, 300], [0, 0, 729, 232], [7, 9, 8, 10]], dtype=torch.int32)
    scene = torch.tensor([0, 1, 0, 1, 0, 1, 0], dtype=torch.int32)
    valid = torch.tensor([1, 1, 1, 1, 1, 1, 0], dtype=torch.bool)
    out = ops.crop_resize_normalize(imgs, scene.cuda(), boxes.cuda(), valid.cuda(), 224, dtype=torch.float32)
    mean = torch.tensor(ops.CLIP_MEAN, device="cuda").view(3, 1, 1)
    std = torch.tensor(ops.CLIP_STD, device="cuda").view(3, 1, 1)
    for i in range(6):
        u8 = ref_crop.torchvision_sequence(imgs[int(scene[i])], [int(v) for v in boxes[i]], 224)
        exp = (u8 / 255.0 - mean) / std
        diff_lsb = ((out[i] - exp).abs() * std * 255.0)
        assert diff_lsb.max() <= 1.01, f"crop {i}: more than 1 LSB off"
        assert (diff_lsb > 0.5).float().mean() < 1e-3, f"crop {i}: too many rounding flips"
    assert (out[6] == 0).all()
    half = ops.crop_resize_normalize(imgs, scene.cuda(), boxes.cuda(), valid.cuda(), 224, dtype=torch.float16)
    torch.testing.assert_close(half.float(), out, rtol=2e-3, atol=2e-3)
    # the tile height of the separable kernel does not change a bit; patch-major output = the same pixels unfolded
    for tr in (1, 4, 8):
        again = ops.crop_resize_normalize(imgs, scene.cuda(), boxes.cuda(), valid.cuda(), 224, dtype=torch.float32,
                                          tile_rows=tr)
        assert torch.equal(again, out), tr
    pm = ops.crop_resize_normalize(imgs, scene.cuda(), boxes.cuda(), valid.cuda(), 224, dtype=torch.float32, patch=32)
    assert pm.shape == (7, 7, 7, 3, 32, 32)
    assert torch.equal(pm, out.view(7, 3, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5))


def test_crop_resize_small_resolution_and_many_crops():
    """a resolution that is not a multiple of the tile height, crops of every aspect, 300 crops in one launch:
    against the reference's torchvision sequence within one LSB"""
    rng = np.random.default_rng(3)
    imgs = torch.from_numpy(rng.integers(0, 256, size=(3, 120, 160, 3), dtype=np.uint8)).cuda()
    n = 300
    x0 = rng.integers(0, 150, n); y0 = rng.integers(0, 110, n)
    x1 = np.minimum(x0 + rng.integers(1, 160, n), 160); y1 = np.minimum(y0 + rng.integers(1, 120, n), 120)
    boxes = torch.from_numpy(np.stack([x0, y0, x1, y1], 1).astype(np.int32))
    scene = torch.from_numpy(rng.integers(0, 3, n).astype(np.int32))
    valid = torch.from_numpy(rng.random(n) > 0.1)
    res = 36
    out = ops.crop_resize_normalize(imgs, scene.cuda(), boxes.cuda(), valid.cuda(), res, dtype=torch.float32)
    mean = torch.tensor(ops.CLIP_MEAN, device="cuda").view(3, 1, 1)
    std = torch.tensor(ops.CLIP_STD, device="cuda").view(3, 1, 1)
    for i in range(0, n, 7):
        if not bool(valid[i]):
            assert (out[i] == 0).all()
            continue
        u8 = ref_crop.torchvision_sequence(imgs[int(scene[i])], [int(v) for v in boxes[i]], res)
        exp = (u8 / 255.0 - mean) / std
        assert ((out[i] - exp).abs() * std * 255.0).max() <= 1.01, f"crop {i}"


def test_clip_image_tower_fp16_kernels_vs_fp32_math():
    """The fp16 CLIP ViT on our GEMM / attention kernels against the same weights in fp32 torch math."""
    from coda_neurips2023_b200 import clip as clip_mod

    torch.manual_seed(0)
    cfg = dict(clip_mod.model.VIT_B32, vision_layers=2, transformer_layers=1)
    model = clip_mod.CLIP(**cfg).cuda().eval()
    ref = clip_mod.CLIP(**cfg).cuda().eval()
    ref.load_state_dict(model.state_dict())
    clip_mod.convert_weights(model)
    x = torch.randn(6, 3, 224, 224, device="cuda")
    with torch.no_grad():
        got = model.encode_image(x)[0].float()
        ref.float()
        # fp32 reference path: plain torch ops (x is fp32 -> none of the fp16 kernel branches trigger)
        exp = ref.visual(x)[0]
    assert ((got - exp).abs().max() / exp.abs().max()).item() < 2e-2   # fp16 weights + activations


@pytest.mark.parametrize("q,g,seed", [(256, 64, 0), (128, 5, 1), (300, 0, 2), (1000, 64, 3)])
def test_novel_candidates_matches_reference_loop(q, g, seed):
    """ops.novel_candidates (2-D NMS + ground-truth rejection + objectness threshold on the device) against the
    step-by-step restatement of the reference loop (oracle/discovery_ref.py: torchvision.ops.nms + cal_iou)."""
    import discovery_ref

    from coda_neurips2023_b200 import ops

    gen = torch.Generator().manual_seed(seed)
    b = 3
    xy = torch.randint(0, 600, (b, q, 2), generator=gen)
    wh = torch.randint(1, 250, (b, q, 2), generator=gen)
    boxes = torch.cat((xy, xy + wh), dim=-1).to(torch.int32)
    boxes[:, : q // 8] = boxes[:, q // 8: 2 * (q // 8)]            # exact duplicates: IoU 1 -> suppressed
    valid = torch.rand(b, q, generator=gen) > 0.15
    obj = torch.rand(b, q, generator=gen)
    obj[:, 5] = obj[:, 6]                                           # a score tie
    ctr = torch.rand(b, q, 1, 3, generator=gen) * 4
    half = torch.rand(b, q, 1, 3, generator=gen) * 0.8 + 0.1
    sign = torch.tensor([[1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1], [-1, 1, 1], [-1, 1, -1], [-1, -1, 1],
                         [-1, -1, -1]], dtype=torch.float32)
    pred = ctr + half * sign
    gctr = torch.rand(b, max(g, 1), 1, 3, generator=gen) * 4
    gt = (gctr + (torch.rand(b, max(g, 1), 1, 3, generator=gen) * 0.8 + 0.1) * sign)[:, :g]
    present = (torch.rand(b, max(g, 1), generator=gen) > 0.5).float()[:, :g]
    for cap in (32, q):
        exp_idx, exp_cnt = discovery_ref.novel_candidates_ref(boxes, valid, obj, pred, gt, present, 0.25, 0.25, 0.4, cap)
        got_idx, got_cnt = ops.novel_candidates(boxes.cuda(), valid.cuda(), obj.cuda(), pred.cuda(), gt.cuda(),
                                                present.cuda(), 0.25, 0.25, 0.4, cap)
        assert torch.equal(got_cnt.cpu(), exp_cnt), (got_cnt, exp_cnt)
        assert torch.equal(got_idx.cpu(), exp_idx)


@pytest.mark.parametrize("b,q,hw,seed", [(4, 256, (531, 730), 0), (2, 1000, (968, 1296), 1), (1, 7, (531, 730), 2)])
def test_boxes_in_image_matches_reference_projection_chain(b, q, hw, seed):
    """ops.boxes_in_image (one fp64 kernel) against the reference's chain of fp64 tensor ops restated in
    oracle/cpu_step.py (models/model_3detr.py:912-968, :1034-1051): integer boxes and the usability flag are exact."""
    import cpu_step

    batch = synthetic.make_batch(b, 3000, seed=seed, image_hw=hw)
    batch["x_offset"] = np.arange(b, dtype=np.int64) * 3
    batch["y_offset"] = np.arange(b, dtype=np.int64) * 5 + 1
    inputs = {k: torch.from_numpy(v) for k, v in batch.items()}
    gen = torch.Generator().manual_seed(seed)
    ctr = torch.rand(b, q, 1, 3, generator=gen) * 6 - 3
    ctr[..., 1] = ctr[..., 1].abs() + 0.2                          # mostly in front of the camera (depth axis = y)
    ctr[:, : q // 5, :, 1] *= -1                                    # ... and a fifth of them behind it
    half = torch.rand(b, q, 1, 3, generator=gen) * 0.9 + 0.05
    sign = torch.tensor([[1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1], [-1, 1, 1], [-1, 1, -1], [-1, -1, 1],
                         [-1, -1, -1]], dtype=torch.float32)
    corners = (ctr + half * sign).contiguous()
    size = (2 * half[:, :, 0]).contiguous()
    size[:, -1] = 0                                                 # zero-size box -> not usable
    exp_boxes, exp_valid = cpu_step._boxes_in_image(corners, size, inputs)
    got_boxes, got_valid = ops.boxes_in_image(corners.cuda(), size.cuda(), {k: v.cuda() for k, v in inputs.items()})
    assert got_valid.dtype == torch.bool and got_boxes.dtype == torch.int32
    assert torch.equal(got_valid.cpu(), exp_valid)
    # truncation of fp64 values that differ in the last ulp (fused multiply-adds) can move a coordinate by one pixel
    # when the value sits on an integer; none of these inputs does
    assert torch.equal(got_boxes.cpu(), exp_boxes)
    assert 0 < int(exp_valid.sum()) < b * q

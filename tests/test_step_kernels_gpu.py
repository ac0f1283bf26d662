"""GPU tests of the step-glue kernels (include/coda_step.h) against plain PyTorch fp32 references, and of
engine.TrainStep's bookkeeping (probe / restore, inactive parameters, parameter groups)."""
import copy
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _lib(built_lib):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_dropout_add_mask_is_shared_by_forward_and_backward():
    from coda_neurips2023_b200 import attention_launch, ops

    torch.manual_seed(0)
    x = torch.randn(2048, 8, 256, device="cuda", requires_grad=True)
    r = torch.randn_like(x, requires_grad=True)
    attention_launch.advance_seed(x.device)
    out = ops.dropout_add(x, r, 0.1, True)
    delta = (out - r).detach()                         # x * (0 or 1 / 0.9) up to the rounding of the addition
    keep = (delta / x.detach()) > 0.5
    mult = keep.float() / 0.9
    assert torch.allclose(delta, x.detach() * mult, rtol=1e-5, atol=2e-6)
    frac = keep.float().mean().item()
    assert abs(frac - 0.9) < 2e-3, frac
    g = torch.randn_like(out)
    out.backward(g)
    assert torch.equal(r.grad, g)
    assert torch.allclose(x.grad, g * mult, rtol=1e-6, atol=1e-7)
    # a second call site draws a different mask, the next step (advanced seed) too
    out2 = ops.dropout_add(x, r, 0.1, True)
    assert ((out2 - r) / x > 0.5).ne(keep).any()
    # p = 0 / eval: exact add, odd length exercises the scalar tail
    a, b = torch.randn(1027, device="cuda"), torch.randn(1027, device="cuda")
    assert torch.equal(ops.dropout_add(a, b, 0.3, False), a + b)
    d = ops.dropout(a, 0.5, True)
    assert ((d == 0) | torch.isclose(d, a * 2)).all() and 0.4 < (d != 0).float().mean().item() < 0.6


@pytest.mark.parametrize("rows,c,relu", [(16384, 512, True), (1000, 64, True), (777, 256, False)])
def test_bn_act_rows_matches_torch(rows, c, relu):
    from coda_neurips2023_b200 import ops

    torch.manual_seed(1)
    bn = torch.nn.BatchNorm1d(c).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    ref_bn = copy.deepcopy(bn)
    h = (torch.randn(rows, c, device="cuda") * 1.7 + 0.3).requires_grad_(True)
    h2 = h.detach().clone().requires_grad_(True)
    out = ops.bn_act_rows(h, bn, relu, 0.0, True)
    exp = ref_bn(h2)
    exp = torch.relu(exp) if relu else exp
    assert torch.allclose(out, exp, rtol=1e-5, atol=2e-6)
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1
    g = torch.randn_like(out)
    out.backward(g)
    exp.backward(g)
    scale = h2.grad.abs().max()
    assert (h.grad - h2.grad).abs().max() <= 2e-5 * scale
    assert torch.allclose(bn.weight.grad, ref_bn.weight.grad, rtol=1e-4, atol=1e-4 * ref_bn.weight.grad.abs().max().item())
    assert torch.allclose(bn.bias.grad, ref_bn.bias.grad, rtol=1e-4, atol=1e-4 * ref_bn.bias.grad.abs().max().item())
    # eval mode: running statistics
    bn.eval(); ref_bn.eval()
    with torch.no_grad():
        o, e = ops.bn_act_rows(h, bn, relu, 0.3, False), ref_bn(h)
    assert torch.allclose(o, torch.relu(e) if relu else e, rtol=1e-5, atol=1e-5)


def test_bn_act_rows_dropout_backward_consistent():
    from coda_neurips2023_b200 import attention_launch, ops

    torch.manual_seed(2)
    bn = torch.nn.BatchNorm1d(512).cuda().train()
    h = torch.randn(4096, 512, device="cuda", requires_grad=True)
    attention_launch.advance_seed(h.device)
    out = ops.bn_act_rows(h, bn, True, 0.3, True)
    plain = torch.relu(torch.nn.functional.batch_norm(h.detach(), None, None, bn.weight, bn.bias, True))
    pos = plain > 1e-3
    mult = (out.detach() / plain)[pos]
    assert ((mult - 1 / 0.7).abs() < 1e-4).logical_or(mult.abs() < 1e-6).all()
    keepfrac = (mult > 0.5).float().mean().item()
    assert abs(keepfrac - 0.7) < 5e-3
    # backward == autograd through the same mask
    mask = torch.zeros_like(plain)
    mask[pos] = mult
    mask[~pos] = (out.detach()[~pos] != 0).float() / 0.7
    h2 = h.detach().clone().requires_grad_(True)
    ref = torch.relu(torch.nn.functional.batch_norm(h2, None, None, bn.weight, bn.bias, True)) * mask
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g)
    assert (h.grad - h2.grad).abs().max() <= 5e-5 * h2.grad.abs().max()


@pytest.mark.parametrize("filter_biases", [False, True])
def test_flat_adamw_matches_torch_adamw_with_clip(filter_biases):
    from coda_neurips2023_b200.engine import FlatAdamW, FlatParameters, _no_decay

    torch.manual_seed(3)
    # odd sizes: chunks start at offsets that are not multiples of 4, one tensor longer than a chunk
    model = torch.nn.Sequential(torch.nn.Linear(37, 501), torch.nn.ReLU(), torch.nn.Linear(501, 129),
                                torch.nn.LayerNorm(129), torch.nn.Linear(129, 7)).cuda()
    ref = copy.deepcopy(model)
    flat = FlatParameters(model)
    lr = torch.tensor(3e-3, device="cuda")
    wds = [0.0 if (filter_biases and _no_decay(n, p)) else 0.1 for n, p in zip(flat.names, flat.params)]
    opt = FlatAdamW(flat, lr, wds, max_norm=0.1)
    decay = [p for n, p in ref.named_parameters() if not (filter_biases and _no_decay(n, p))]
    nodecay = [p for n, p in ref.named_parameters() if filter_biases and _no_decay(n, p)]
    groups = [{"params": decay, "weight_decay": 0.1}] + ([{"params": nodecay, "weight_decay": 0.0}] if nodecay else [])
    topt = torch.optim.AdamW(groups, lr=3e-3)
    for it in range(5):
        x = torch.randn(64, 37, device="cuda")
        flat.zero_grad()
        model(x).pow(2).mean().backward()
        opt.step()
        topt.zero_grad()
        ref(x).pow(2).mean().backward()
        norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.1)
        topt.step()
        assert abs(float(opt.grad_norm) - float(norm)) <= 1e-5 * float(norm)
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-7), it


def test_flat_adamw_leaves_inactive_parameters_alone():
    from coda_neurips2023_b200.engine import FlatAdamW, FlatParameters

    torch.manual_seed(4)
    model = torch.nn.ModuleDict({"used": torch.nn.Linear(8, 8), "unused": torch.nn.Linear(8, 8)}).cuda()
    before = model["unused"].weight.detach().clone()
    flat = FlatParameters(model)
    active = [n.startswith("used") for n in flat.names]
    opt = FlatAdamW(flat, torch.tensor(1e-2, device="cuda"), 0.1, active)
    flat.zero_grad()
    model["used"](torch.randn(4, 8, device="cuda")).sum().backward()
    opt.step()
    assert torch.equal(model["unused"].weight, before)          # no weight decay on a parameter without gradient
    assert not torch.equal(model["used"].weight, flat.flat_param.data.new_zeros(8, 8))


def test_train_step_prepare_and_capture_leave_no_trace():
    """probe + graph warm-up must not advance the optimizer, BatchNorm buffers or the dropout counter;
    parameters without gradient (text head when its loss weight is 0) are not decayed; the step is deterministic
    given the same state, eager == captured graph."""
    from coda_neurips2023_b200 import attention_launch, synthetic
    from coda_neurips2023_b200.criterion import build_criterion
    from coda_neurips2023_b200.engine import TrainStep
    from coda_neurips2023_b200.models import build_model

    args = synthetic.make_args(nqueries=128, preenc_npoints=256, dec_dim=128, dec_nlayers=2, dec_ffn_dim=64,
                               enc_dropout=0.0, dec_dropout=0.0, mlp_dropout=0.0,   # call-site salts differ eager/graph
                               loss_predicted_region_embed_l1_weight=0.0, filter_biases_wd=True)
    cfg = synthetic.SyntheticDatasetConfig(args)

    def make():
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model, _ = build_model(args, cfg)
        return model.cuda().train(), build_criterion(args, cfg).cuda()

    batch = synthetic.to_device(synthetic.make_batch(2, 3000, seed=1), "cuda")
    model, crit = make()
    w0 = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    bn0 = {n: b.detach().clone() for n, b in model.named_buffers() if "clip_model" not in n}
    seed0 = int(attention_launch.seed_counter(torch.device("cuda", 0)))
    step = TrainStep(args, model, crit, torch.device("cuda", 0))
    np.random.seed(11)
    step.capture(batch, warmup=2)
    assert int(step.optimizer.state[0]) == 0, "warm-up advanced the optimizer"
    assert int(attention_launch.seed_counter(torch.device("cuda", 0))) == seed0
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert torch.equal(p, w0[n]), n
    for n, b in model.named_buffers():
        if "clip_model" not in n:
            assert torch.equal(b, bn0[n]), n
    assert any("text_correlation_head" in n for n in step.inactive_names)
    np.random.seed(12)
    loss_g, _ = step(batch, 0.0)
    loss_g = float(loss_g)
    assert int(step.optimizer.state[0]) == 1
    for n, p in model.named_parameters():
        if "text_correlation_head" in n:
            assert torch.equal(p, w0[n]), f"{n} was decayed without a gradient"
    wg = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    # the same step eagerly from the same initial state
    model2, crit2 = make()
    attention_launch.seed_counter(torch.device("cuda", 0)).fill_(seed0)
    step2 = TrainStep(args, model2, crit2, torch.device("cuda", 0))
    step2.prepare(batch)
    np.random.seed(12)
    loss_e, _ = step2(batch, 0.0)
    assert abs(float(loss_e) - loss_g) <= 1e-5 * abs(loss_g)
    changed = 0
    for n, p in model2.named_parameters():
        if p.requires_grad:
            assert torch.allclose(p, wg[n], rtol=1e-4, atol=1e-6), n
            changed += int(not torch.equal(p, w0[n]))
    assert changed > 100

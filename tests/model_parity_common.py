"""Shared by the CPU plumbing test and the GPU parity test: build OUR model/criterion
for a golden case, with the name-keyed weights and the golden text features."""
from pathlib import Path

import numpy as np
import torch

from coda_neurips2023_b200 import clip as clip_mod
from coda_neurips2023_b200 import synthetic
from coda_neurips2023_b200.criterion import build_criterion
from coda_neurips2023_b200.models import build_model
from param_fill import fill_by_name

GOLDEN = Path(__file__).resolve().parent / "golden"

CASES = {
    "stage1_small": (2, 3000, dict(nqueries=128, preenc_npoints=256, dec_dim=128, dec_nlayers=2, dec_ffn_dim=64,
                                   enc_dropout=0.0, dec_dropout=0.0, mlp_dropout=0.0)),
    "stage2_weak": (2, 2500, dict(nqueries=128, preenc_npoints=256, dec_dim=128, dec_nlayers=2, dec_ffn_dim=64,
                                  enc_dropout=0.0, dec_dropout=0.0, mlp_dropout=0.0, if_clip_weak_labels=True,
                                  loss_feat_seen_softmax_weakly_loss_with_novel_cate_confi_weight=1.0,
                                  confidence_type="clip-max-prob")),
}
TINY_CLIP = dict(embed_dim=512, image_resolution=224, vision_layers=2, vision_width=128, vision_patch_size=32,
                 context_length=77, vocab_size=49408, transformer_width=64, transformer_heads=1,
                 transformer_layers=1)


def build(name: str, device: str):
    batch, npoints, over = CASES[name]
    golden = np.load(GOLDEN / f"model_{name}.npz")
    args = synthetic.make_args(**over)
    cfg = synthetic.SyntheticDatasetConfig(args)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, _ = build_model(args, cfg)
    # the golden generator fills the WHOLE reference model by name, CLIP included
    # (parameter paths clip_model.*), so install the small CLIP first
    tiny = clip_mod.CLIP(**TINY_CLIP).float().eval()
    for p in tiny.parameters():
        p.requires_grad = False
    model.clip_model = tiny
    model.test_clip_model = tiny
    model.res_encoder = tiny.visual
    model.logit_scale = tiny.logit_scale
    model.clip_resolution = 224
    fill_by_name(model, seed=3)
    model.device = device
    model = model.to(device)
    model.text_features_fg_norm = torch.from_numpy(golden["text_features_fg_norm"]).to(device)
    criterion = build_criterion(args, cfg).to(device)
    model.train()
    model.clip_model.eval()
    inputs = synthetic.to_device(synthetic.make_batch(batch, npoints, seed=5), device)
    return args, model, criterion, inputs, golden


def run(name: str, device: str):
    args, model, criterion, inputs, golden = build(name, device)
    np.random.seed(123)
    out = model(inputs, curr_epoch=0)
    loss, loss_dict = criterion(out, inputs)
    loss.backward()
    return model, out, loss, loss_dict, golden


def compare(model, out, loss, loss_dict, golden, rtol, atol, check_grads=True, grad_rtol=None):
    """Returns a dict name -> max relative error; raises on the first mismatch."""
    errs = {}

    def chk(key, got, sl=None, rtol=rtol):
        exp = golden[key]
        g = got.detach().float().cpu().numpy()
        if sl is not None:
            g = g[sl]
        assert g.shape == exp.shape, (key, g.shape, exp.shape)
        if exp.dtype.kind in "iu":
            assert np.array_equal(g.astype(exp.dtype), exp), key
            errs[key] = 0.0
            return
        scale = max(float(np.abs(exp).max()), 1e-6)
        err = float(np.abs(g - exp).max()) / scale
        errs[key] = err
        assert err <= rtol + atol / scale, f"{key}: max err {err:.3e} (scale {scale:.3e})"

    last = out["outputs"]
    for k in ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits", "angle_residual",
              "angle_continuous", "objectness_prob", "box_corners", "box_corners_xyz"):
        chk(f"last.{k}", last[k])
    chk("last.text_correlation_embedding", last["text_correlation_embedding"], np.s_[:, ::4, ::8])
    chk("last.gt_text_correlation_embedding_mask", last["gt_text_correlation_embedding_mask"])
    chk("last.gt_text_correlation_embedding", last["gt_text_correlation_embedding"], np.s_[:, :, ::8])
    chk("last.weak_confidence_weight", last["weak_confidence_weight"])
    for i, aux in enumerate(out["aux_outputs"]):
        chk(f"aux{i}.sem_cls_logits", aux["sem_cls_logits"])
        chk(f"aux{i}.center_normalized", aux["center_normalized"])
    exp_loss = float(golden["loss"])
    errs["loss"] = abs(float(loss) - exp_loss) / abs(exp_loss)
    assert errs["loss"] <= rtol, f"loss {float(loss)} vs {exp_loss}"
    for k in golden.files:
        if k.startswith("loss_dict."):
            name = k[len("loss_dict."):]
            assert name in loss_dict, f"missing loss_dict key {name}"
            e = float(golden[k])
            err = abs(float(loss_dict[name]) - e) / max(abs(e), 1e-6)
            errs[k] = err
            assert err <= max(rtol, 1e-5) + atol, f"{k}: {float(loss_dict[name])} vs {e}"
    ours = sorted(k for k in model.state_dict().keys() if "clip_model" not in k)
    assert ours == list(golden["state_dict_keys"]), "state-dict keys differ from the reference model"
    if check_grads:
        params = dict(model.named_parameters())
        for k in golden.files:
            if k.startswith("grad."):
                # gradients cross ~13 layers and train-mode BatchNorm: looser than the forward bar
                chk(k, params[k[len("grad."):]].grad, rtol=grad_rtol if grad_rtol is not None else 10 * rtol)
    return errs

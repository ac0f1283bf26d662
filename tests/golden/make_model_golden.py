"""Generates tests/golden/model_*.npz from the REFERENCE's own Python model and
criterion, run on CPU in this container (tests/golden/_reference_harness.py):
reference modules are imported from /root/reference, the CUDA-only pointnet2
extension is replaced by the oracle, the absent CLIP checkpoint by a small
random-init CLIP of the reference's own class.  Weights are filled by name
(tests/param_fill.py), inputs come from coda_neurips2023_b200.synthetic, so only
outputs are stored.

    python tests/golden/make_model_golden.py            (writes into tests/golden/)
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(HERE))

import _reference_harness as H  # noqa: E402
from coda_neurips2023_b200 import synthetic  # noqa: E402
from param_fill import fill_by_name  # noqa: E402

import model_parity_common as mpc  # noqa: E402  (the case table is shared with the parity tests)

CASES = mpc.CASES
TINY_CLIP = mpc.TINY_CLIP

# arguments the reference reads that our make_args does not carry (weight-0 / off everywhere)
REF_EXTRA = dict(
    begin_keep_epoch=10 ** 14, online_nms_update_save_novel_label_clip_driven_with_cate_confidence=False,
    save_objectness=0.75, online_nms_update_save_epoch=10, clip_driven_keep_thres=1e6, eval_layer_id=-1,
    if_accumulate_former_pseudo_labels=False, if_with_clip=False, if_with_clip_embed=False, if_use_gt_box=False,
    if_expand_box=False, if_with_fake_classes=False, pooling_methods="average", if_keep_box=False,
    if_select_box_by_objectness=False, online_nms_update_novel_label=False,
    online_nms_update_accumulate_novel_label=False, online_nms_update_accumulate_epoch=10,
    only_image_class=False, only_prompt_loss=False, if_skip_no_seen_scene_objectness=False,
    if_only_seen_in_loss=False, confidence_type_in_datalayer="zero-out", reset_scannet_num=0,
)


def reference_args(overrides):
    from coda_neurips2023_b200.criterion import _WEIGHT_ARGS

    a = synthetic.make_args(**overrides)
    for k, v in REF_EXTRA.items():
        if not hasattr(a, k):
            setattr(a, k, v)
    for attr in _WEIGHT_ARGS.values():
        if not hasattr(a, attr):
            setattr(a, attr, 0.0)
    return a


def run_reference(name):
    batch, npoints, over, extra = mpc.case(name)
    args = reference_args(over)
    m3 = H.load("models.model_3detr")
    crit_mod = H.load("criterion")
    box_util = H.load("utils.box_util")
    clip_pkg = H.load("CLIP.clip.clip")
    clip_model_mod = H.load("CLIP.clip.model")

    class Cfg(synthetic.SyntheticDatasetConfig):  # corner builders of the REFERENCE
        def box_parametrization_to_corners(self, c, s, a):
            return box_util.get_3d_box_batch_tensor(s, a, box_util.flip_axis_to_camera_tensor(c))

        def box_parametrization_to_corners_xyz(self, c, s, a):
            return box_util.get_3d_box_batch_tensor_xyz(s, a, c)

    cfg = Cfg(args)

    def fake_clip_load(path, device="cpu", download_root=None, if_transform_tensor=True, **kw):
        torch.manual_seed(0)
        model = clip_model_mod.CLIP(**TINY_CLIP).float().eval()
        fill_by_name(model, seed=11)
        return model, clip_pkg._transform_for_tensor(model.visual.input_resolution)

    clip_pkg.load = fake_clip_load
    sys.modules["CLIP.clip"].clip.load = fake_clip_load
    torch.manual_seed(0)
    model, _ = m3.build_3detr_predictedbox_distillation_head(args, cfg)
    fill_by_name(model, seed=3)
    # re-derive the text features from the filled CLIP (they were computed in __init__)
    with torch.no_grad():
        if "text_rows" in extra:   # seeded random unit rows stand in for the class prompts
            model.text_features_fg_norm = torch.from_numpy(mpc.text_rows(extra["text_rows"]))
            model.text_features_fg = model.text_features_fg_norm.clone()
        else:
            model.text_features_fg = model.clip_model.encode_text(model.text).to(torch.float32)
            model.text_features_fg_norm = model.text_features_fg / model.text_features_fg.norm(dim=1, keepdim=True)
    criterion = crit_mod.build_criterion(args, cfg)
    model.train()
    model.clip_model.eval()
    inputs = {k: torch.from_numpy(v) for k, v in
              synthetic.make_batch(batch, npoints, seed=5, image_hw=extra.get("image_hw", (531, 730))).items()}
    if extra.get("pseudo"):
        import tempfile

        tmp = tempfile.mkdtemp(prefix="coda_pseudo_ref_")
        inputs["pseudo_box_path"] = [f"{tmp}/scene{i}.npy" for i in range(batch)]
    np.random.seed(123)  # box selection draws (model_3detr.py:991)
    out = model(inputs, curr_epoch=extra.get("curr_epoch", 0))
    model._golden_pseudo_paths = inputs.get("pseudo_box_path")
    loss, loss_dict = criterion(out, inputs)
    loss.backward()
    return args, model, out, loss, loss_dict


KEEP = ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits", "angle_residual",
        "angle_continuous", "objectness_prob", "box_corners", "box_corners_xyz")
GRADS = ("pre_encoder.mlp_module.layer0.conv.weight", "pre_encoder.mlp_module.layer2.conv.weight",
         "encoder.layers.0.self_attn.in_proj_weight", "encoder.layers.2.linear1.weight",
         "encoder_to_decoder_projection.layers.0.weight", "query_projection.layers.0.weight",
         "decoder.layers.0.self_attn.in_proj_weight", "decoder.layers.1.multihead_attn.out_proj.weight",
         "decoder.layers.{last}.multihead_attn.in_proj_weight", "decoder.layers.{last}.linear2.weight",
         "mlp_heads.center_head.layers.0.weight", "mlp_heads.text_correlation_head.layers.8.weight",
         "mlp_heads.sem_cls_head.layers.8.bias", "decoder.norm.weight")


def thin(a: np.ndarray) -> np.ndarray:
    """Large gradients are stored as a [::4, ::4] lattice (tests/model_parity_common.py applies the same rule)."""
    return a[::4, ::4] if (a.ndim >= 2 and a.size > 65536) else a


def main():
    only = sys.argv[1:]
    for name in CASES:
        if only and name not in only:
            continue
        args, model, out, loss, loss_dict = run_reference(name)
        last = out["outputs"]
        blob = {f"last.{k}": last[k].detach().numpy() for k in KEEP}
        blob["last.text_correlation_embedding"] = last["text_correlation_embedding"].detach().numpy()[:, ::4, ::8]
        blob["last.gt_text_correlation_embedding"] = last["gt_text_correlation_embedding"].numpy()[:, :, ::8]
        blob["last.gt_text_correlation_embedding_mask"] = last["gt_text_correlation_embedding_mask"].numpy()
        blob["last.weak_box_cate_label"] = last["weak_box_cate_label"].numpy()
        blob["last.weak_confidence_weight"] = last["weak_confidence_weight"].numpy()
        blob["text_features_fg_norm"] = model.text_features_fg_norm.numpy()
        full = name in mpc.FULL_SIZE
        for i, aux in enumerate(out["aux_outputs"]):
            blob[f"aux{i}.sem_cls_logits"] = aux["sem_cls_logits"].detach().numpy()
            blob[f"aux{i}.center_normalized"] = aux["center_normalized"].detach().numpy()
            if full:   # every decoder layer, every head
                for k in ("size_normalized", "angle_logits", "angle_residual"):
                    blob[f"aux{i}.{k}"] = aux[k].detach().numpy()
                blob[f"aux{i}.text_correlation_embedding"] = aux["text_correlation_embedding"].detach().numpy()[:, ::4, ::8]
        if getattr(model, "_golden_pseudo_paths", None):
            import os

            arrs = [np.load(p) if os.path.exists(p) else np.zeros((0, 10), np.float32) for p in model._golden_pseudo_paths]
            blob["pseudo.count"] = np.array([len(a) for a in arrs], np.int64)
            blob["pseudo.rows"] = np.concatenate(arrs, axis=0).astype(np.float32).reshape(-1, 10)
            print("pseudo labels per scene:", blob["pseudo.count"], flush=True)
        blob["loss"] = np.float32(loss.item())
        for k, v in loss_dict.items():
            blob[f"loss_dict.{k}"] = np.float32(float(v))
        blob["state_dict_keys"] = np.array(sorted(k for k in model.state_dict().keys() if "clip_model" not in k))
        g = dict(model.named_parameters())
        names = GRADS if full else (GRADS[0], GRADS[2], GRADS[7], GRADS[10], GRADS[13])
        for pname in names:
            pname = pname.format(last=args.dec_nlayers - 1)
            blob[f"grad.{pname}"] = thin(g[pname].grad.numpy())
        np.savez_compressed(HERE / f"model_{name}.npz", **blob)
        print("wrote", name, "loss", float(loss), flush=True)


if __name__ == "__main__":
    main()

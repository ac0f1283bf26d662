"""CPU (no GPU) plumbing parity: OUR model + criterion Python code, with the CUDA ops
replaced by the CPU stand-ins of oracle/cpu_step.py, against golden outputs of
the REFERENCE model + criterion run on CPU (tests/golden/make_model_golden.py).
This pins everything that is not a kernel: module wiring, parameter names, box
decoding, projection to the image, crop selection, matching cost, every active
loss and its gradient.  The kernels themselves are tested on the GPU."""
import numpy as np
import pytest
import torch

import cpu_step as cpu_shims
import model_parity_common as mpc


@pytest.mark.parametrize("name", list(mpc.CASES))
def test_model_and_criterion_match_reference_on_cpu(name):
    torch.manual_seed(0)
    with cpu_shims.installed():
        model, out, loss, loss_dict, golden = mpc.run(name, "cpu")
        # full-size cases: two fp32 evaluations of the encoder-side gradients differ by up to 5e-3 (make_cpu_noise.py)
        errs = mpc.compare(model, out, loss, loss_dict, golden, rtol=2e-4, atol=1e-5,
                           grad_rtol=1e-2 if name in mpc.FULL_SIZE else None)
    worst = max(errs, key=errs.get)
    print(f"{name}: worst {worst} = {errs[worst]:.2e}")


@pytest.mark.parametrize("name", list(mpc.CASES))
def test_stacked_criterion_call_equals_the_per_layer_loop(name):
    """SetCriterion.forward takes the seven auxiliary decoder outputs in ONE stacked call when the model hands it
    `stacked_layers`; without that key (a reference-style outputs dict) it loops layer by layer like
    criterion.py:1205-1216.  Both must give the same total and the same loss_dict entries."""
    torch.manual_seed(0)
    with cpu_shims.installed():
        args, model, criterion, inputs, _ = mpc.build(name, "cpu")
        np.random.seed(123)
        out = model(inputs, curr_epoch=0)
        assert "stacked_layers" in out and len(out["aux_outputs"]) >= 1
        loss_a, dict_a = criterion(out, dict(inputs))
        plain = {"outputs": dict(out["outputs"]), "aux_outputs": [dict(a) for a in out["aux_outputs"]]}
        loss_b, dict_b = criterion(plain, dict(inputs))
    assert set(dict_a) == set(dict_b)
    assert abs(float(loss_a) - float(loss_b)) <= 2e-6 * abs(float(loss_b))
    for k in dict_a:
        a, b = float(dict_a[k]), float(dict_b[k])
        assert abs(a - b) <= 2e-6 * max(abs(b), 1e-3), k

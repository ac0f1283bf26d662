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
        errs = mpc.compare(model, out, loss, loss_dict, golden, rtol=2e-4, atol=1e-5)
    worst = max(errs, key=errs.get)
    print(f"{name}: worst {worst} = {errs[worst]:.2e}")

"""How far does ANOTHER fp32 implementation of the same step sit from the reference's fp32 goldens?

Runs OUR model / criterion Python on CPU through the CPU restatement (oracle/cpu_step.py: torch CPU ops + the C
oracle) for the full-size golden cases and records, per golden key, the max relative deviation from the golden in
tests/golden/model_<case>_cpu_noise.json.  Forward outputs agree to ~1e-5; some gradients (weights in front of a
train-mode BatchNorm, 13 layers deep) differ by up to 5e-3 between two fp32 evaluations -- the GPU parity test
therefore holds a gradient to max(5e-3, 3 x this fp32-vs-fp32 noise), not to a bar fp32 itself does not meet.

    python tests/golden/make_cpu_noise.py
"""
import json
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))

import cpu_step  # noqa: E402
import model_parity_common as mpc  # noqa: E402


def main():
    for name in mpc.FULL_SIZE:
        torch.manual_seed(0)
        with cpu_step.installed():
            model, out, loss, ld, golden = mpc.run(name, "cpu")
            errs = mpc.compare(model, out, loss, ld, golden, rtol=1.0, atol=1e-5, grad_rtol=1.0)
        (HERE / f"model_{name}_cpu_noise.json").write_text(json.dumps({k: float(f"{v:.3e}") for k, v in sorted(errs.items())}, indent=0))
        worst = max(errs, key=errs.get)
        print(name, "worst", worst, errs[worst])


if __name__ == "__main__":
    main()

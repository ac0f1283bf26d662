"""Runs a script of an unmodified CoDA checkout with this package in place of its
`models`, `criterion` and `third_party_pointnet2`:

    cd /path/to/CoDA_NeurIPS2023
    python -m coda_neurips2023_b200.dropin.launch main.py --dataset_name ... --model_name 3detr_predictedbox_distillation ...

(`python main.py` would put the checkout itself first on sys.path; this launcher puts the shim
directory before it, then executes the script as __main__.)
"""
import runpy
import sys
from pathlib import Path


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = Path(sys.argv[1]).resolve()
    shim = str(Path(__file__).resolve().parent)
    sys.argv = sys.argv[1:]
    sys.path.insert(0, str(script.parent))
    sys.path.insert(0, shim)
    runpy.run_path(str(script), run_name="__main__")


if __name__ == "__main__":
    main()

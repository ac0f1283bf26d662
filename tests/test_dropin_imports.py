"""The import surface the reference's callers use resolves to this package (no GPU needed)."""
import importlib
import sys
from pathlib import Path

SHIM = Path(__file__).resolve().parent.parent / "coda_neurips2023_b200" / "dropin"


def test_reference_import_paths_resolve(built_lib):
    sys.path.insert(0, str(SHIM))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in
             ("models", "criterion", "third_party_pointnet2", "pointnet2")}
    try:
        models = importlib.import_module("models")
        assert set(models.MODEL_FUNCS) == {"3detrmulticlasshead", "3detr_predictedbox_distillation"}
        crit = importlib.import_module("criterion")
        assert callable(crit.build_criterion)
        pu = importlib.import_module("third_party_pointnet2.pointnet2.pointnet2_utils")
        for name in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate",
                     "grouping_operation", "ball_query", "QueryAndGroup", "GroupAll"):
            assert hasattr(pu, name), name
        pm = importlib.import_module("third_party_pointnet2.pointnet2.pointnet2_modules")
        assert hasattr(pm, "PointnetSAModuleVotes")
        ext = importlib.import_module("pointnet2._ext")
        for name in ("furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
                     "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"):
            assert callable(getattr(ext, name)), name   # bindings.cpp:9-22
    finally:
        sys.path.remove(str(SHIM))
        for k in list(sys.modules):
            if k.split(".")[0] in ("models", "criterion", "third_party_pointnet2", "pointnet2"):
                del sys.modules[k]
        sys.modules.update(saved)


def test_library_exports_every_declared_symbol(built_lib):
    from coda_neurips2023_b200 import _lib

    names = _lib.declared_symbols()
    assert len(names) >= 25
    lib = _lib.lib()
    assert all(hasattr(lib, n) for n in names)
    assert lib.coda_abi_version() >= 1

"""`import pointnet2._ext as _ext` (reference pointnet2_utils.py:23) -> the C-ABI binding."""
from coda_neurips2023_b200.pointnet2._ext import *  # noqa: F401,F403
from coda_neurips2023_b200.pointnet2 import _ext as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})

"""reference third_party_pointnet2/pointnet2/pointnet2_utils.py -> coda_neurips2023_b200.pointnet2.pointnet2_utils"""
from coda_neurips2023_b200.pointnet2.pointnet2_utils import *  # noqa: F401,F403
from coda_neurips2023_b200.pointnet2 import pointnet2_utils as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})

"""Small invocation of every pointnet2 kernel, meant to run under compute-sanitizer."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from coda_neurips2023_b200 import synthetic  # noqa: E402
from coda_neurips2023_b200.pointnet2 import _ext as ours  # noqa: E402
from coda_neurips2023_b200._lib import lib  # noqa: E402

for cl in (1, 8):
    lib().coda_fps_set_cluster(cl)
    xyz = torch.from_numpy(synthetic.point_clouds(2, 5000, seed=0)).cuda()
    inds = ours.furthest_point_sampling(xyz, 64)
lib().coda_fps_set_cluster(0)
new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
idx, g = ours.query_and_group_xyz(xyz, new_xyz, 0.3, 16, True)
ours.ball_query(new_xyz, xyz, 0.3, 16)
flipped = xyz.transpose(1, 2).contiguous()
ours.group_points_grad(ours.group_points(flipped, idx), idx, 5000)
ours.gather_points_grad(ours.gather_points(flipped, inds), inds, 5000)
d2, nn = ours.three_nn(xyz[:, :100].contiguous(), new_xyz)
f = torch.randn(2, 4, 64, device="cuda")
w = torch.rand(2, 100, 3, device="cuda")
ours.three_interpolate_grad(ours.three_interpolate(f, nn, w), nn, w, 64)
torch.cuda.synchronize()
print("sanitize run ok")

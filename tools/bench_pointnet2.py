"""Micro-benchmark of the set-abstraction ops on one GPU: our kernels vs the
reference's extension (oracle/_ref) when present.  CUDA-event timing, warm-up,
median of `reps`.  Prints one JSON line per op."""
import importlib
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from coda_neurips2023_b200 import synthetic  # noqa: E402
from coda_neurips2023_b200.pointnet2 import _ext as ours  # noqa: E402
from coda_neurips2023_b200._lib import lib  # noqa: E402

try:
    sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
    ref = importlib.import_module("pointnet2._ext")
except Exception:  # noqa: BLE001
    ref = None


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    for (b, n, m) in [(8, 20000, 2048), (8, 2048, 256), (8, 40000, 2048)]:
        xyz = torch.from_numpy(synthetic.point_clouds(b, n, seed=0)).cuda()
        row = {"op": "fps", "b": b, "n": n, "m": m}
        for cl in (0, 1, 2, 4, 8):
            lib().coda_fps_set_cluster(cl)
            try:
                row[f"ours_cl{cl}_ms"] = round(timeit(lambda: ours.furthest_point_sampling(xyz, m)), 4)
            except Exception as e:  # noqa: BLE001
                row[f"ours_cl{cl}_ms"] = str(e)[:40]
        lib().coda_fps_set_cluster(0)
        row["us_per_round"] = round(1e3 * row["ours_cl0_ms"] / max(m - 1, 1), 4)
        if ref is not None:
            row["ref_ms"] = round(timeit(lambda: ref.furthest_point_sampling(xyz, m), reps=5, warm=1), 4)
        print(json.dumps(row), flush=True)
        inds = ours.furthest_point_sampling(xyz, m)
        new_xyz = torch.gather(xyz, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
        flipped = xyz.transpose(1, 2).contiguous()
        row = {"op": "ball_query", "b": b, "n": n, "m": m,
               "ours_ms": round(timeit(lambda: ours.ball_query(new_xyz, xyz, 0.2, 64)), 4),
               "ours_fused_group_ms": round(timeit(lambda: ours.query_and_group_xyz(xyz, new_xyz, 0.2, 64, True)), 4)}
        if ref is not None:
            def ref_seq():
                idx = ref.ball_query(new_xyz, xyz, 0.2, 64)
                g = ref.group_points(flipped, idx)
                g -= new_xyz.transpose(1, 2).unsqueeze(-1)
                g /= 0.2
                return g
            row["ref_ball_query_ms"] = round(timeit(lambda: ref.ball_query(new_xyz, xyz, 0.2, 64), reps=5, warm=1), 4)
            row["ref_query_group_norm_ms"] = round(timeit(ref_seq, reps=5, warm=1), 4)
        print(json.dumps(row), flush=True)
        d = {"op": "three_nn", "b": b, "n": n, "m": m,
             "ours_ms": round(timeit(lambda: ours.three_nn(xyz, new_xyz)), 4)}
        if ref is not None:
            d["ref_ms"] = round(timeit(lambda: ref.three_nn(xyz, new_xyz), reps=5, warm=1), 4)
        print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()

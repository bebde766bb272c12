"""Generic per-pose scoring (pcl_score) vs structured-grid scoring (pcl_score_grid) on the C2 workload and on a
yaw-only grid:   python scripts/grid_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from piccolo_b200.utils import generate_rot_points, grid_poses
from scripts.perf_probe import timeit

dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
fmt = sys.argv[1] if len(sys.argv) > 1 else "auto"
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img, fmt)
trans = torch.from_numpy(np.ascontiguousarray(synth.pose_grid(sc.room, (5, 5, 3), 1)[:, :3])).to(dev)
cases = {"lattice 75x24": generate_rot_points({"yaw_only": False, "num_yaw": 4, "num_pitch": 4, "num_roll": 4}).to(dev),
         "yaw-only 75x8": generate_rot_points({"yaw_only": True, "num_yaw": 8}).to(dev),
         "yaw-only 75x16": generate_rot_points({"yaw_only": True, "num_yaw": 16}).to(dev)}
for name, rot in cases.items():
    poses = grid_poses(trans, rot)
    a, _ = engine.score(cloud, image, poses)
    b, _ = engine.score_grid(cloud, image, trans, rot)
    rel = ((a - b).abs() / a.abs()).max().item()
    same = torch.equal(engine.topk(a, 50), engine.topk(b, 50))
    for swap in ("0", "1"):
        _lib.set_option("SWAP", int(swap)); _lib.set_option("GRID_SWAP", int(swap))
        ta = timeit(lambda: engine.score(cloud, image, poses), iters=10)
        tb = timeit(lambda: engine.score_grid(cloud, image, trans, rot), iters=10)
        print(f"[{fmt} swap={swap}] {name}: per-pose {ta:.3f} ms ({len(poses)*1e6/ta/1e6:.1f} G/s)  structured {tb:.3f} ms ({len(poses)*1e6/tb/1e6:.1f} G/s)  "
              f"speed-up {ta/tb:.2f}x  max rel diff {rel:.2e}  same top-50 {same}", flush=True)

"""Structured-grid scoring at C3 sizes (10 M points, 2048x4096, 64 translations x 16 yaws) per texel format:
   python scripts/grid_probe_c3.py [fmt ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from piccolo_b200.utils import grid_poses
from scripts.perf_probe import timeit

dev = torch.device("cuda:0")
N, H = int(os.environ.get("N", 10_000_000)), int(os.environ.get("H", 2048))
sc = synth.make_scene(N, H, 2 * H, room=(40.0, 30.0, 3.0), seed=5)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud = engine.Cloud(xyz, rgb)
g = torch.from_numpy(synth.pose_grid(sc.room, (8, 8, 1), 16)).to(dev)
trans, rot = g[::16, :3].contiguous(), g[:16, 3:].contiguous()
poses = grid_poses(trans, rot)
ref = None
for fmt in (sys.argv[1:] or ["tex", "u8p", "u8q", "f16d"]):
    image = engine.Image(img, fmt)
    a, _ = engine.score(cloud, image, poses)
    b, _ = engine.score_grid(cloud, image, trans, rot)
    ref = a if ref is None else ref
    _lib.set_option("SWAP", 1)
    ts = timeit(lambda: engine.score(cloud, image, poses), iters=2, warm=1, repeats=3)
    _lib.set_option("SWAP", 0)
    ta = timeit(lambda: engine.score(cloud, image, poses), iters=2, warm=1, repeats=3)
    tb = timeit(lambda: engine.score_grid(cloud, image, trans, rot), iters=2, warm=1, repeats=3)
    ev = len(poses) * N
    print(f"[{fmt}] N={N} {H}x{2*H} {len(trans)}x{len(rot)}: per-pose {ta:.2f} ms ({ev/ta/1e6:.1f} G/s; block order swapped {ts:.2f} ms)  structured {tb:.2f} ms ({ev/tb/1e6:.1f} G/s)  "
          f"speed-up {ta/tb:.2f}x  vs first fmt per-pose: max rel {((b-ref).abs()/ref).max().item():.2e}", flush=True)
    del image

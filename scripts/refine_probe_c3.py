"""Refinement loop (B=6, moving poses) at large sizes per texel format: python scripts/refine_probe_c3.py [fmt ...]  (env N, H)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import engine, synth
from scripts.perf_probe import timeit

dev = torch.device("cuda:0")
N, H = int(os.environ.get("N", 10_000_000)), int(os.environ.get("H", 2048))
sc = synth.make_scene(N, H, 2 * H, room=(40.0, 30.0, 3.0), seed=5)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud = engine.Cloud(xyz, rgb)
rng = np.random.default_rng(0)
starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.3, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
for fmt in (sys.argv[1:] or ["tex", "u8q", "f16d", "u8p"]):
    image = engine.Image(img, fmt)
    ref = engine.Refiner(6, 0.1, 0.8, 5, True)
    def run():
        ref.reset(starts); ref.run(cloud, image, 30)
    ms = timeit(run, iters=2, warm=1, repeats=3)
    print(f"[{fmt}] N={N} {H}x{2*H}: refine B=6 {ms/30*1e3:.1f} us/iter ({6*N/(ms/30)/1e6:.1f} G/s)", flush=True)
    del image

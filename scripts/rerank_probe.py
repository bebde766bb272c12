"""Histogram re-rank timing at C2 sizes (50 candidates x 1 M points, 1024x2048) + agreement with the CPU oracle on the small fixture."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sc = synth.make_scene(N, H, 2 * H, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud = engine.Cloud(xyz, rgb)
poses = torch.from_numpy(synth.pose_grid(sc.room, (5, 5, 2), 1)[:50]).to(dev)
ms = timeit(lambda: engine.hist_rerank(cloud, img, poses, 4, 4), iters=5)
print(f"re-rank of 50 candidates, N={N}, {H}x{2*H}: {ms:.3f} ms")

"""Refinement loop timing (C2 sizes by default): persistent vs per-iteration launches, candidates per pose block.
usage: python scripts/refine_probe.py [N] [H] [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
Bs = [int(sys.argv[3])] if len(sys.argv) > 3 else [6]
sc = synth.make_scene(N, H, 2 * H, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
rng = np.random.default_rng(0)
starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(16)]).astype(np.float32)).to(dev)
for B in Bs:
    ref = engine.Refiner(B, 0.1, 0.8, 5, True)
    def run():
        ref.reset(starts[:B]); ref.run(cloud, image, 100)
    for persist in (1, 0):
        for npb in ((0, 1, 2, 3, 4) if persist else (0,)):
            _lib.set_option("PERSIST", persist); _lib.set_option("RF_NPB", npb)
            ms = timeit(run, iters=3, warm=1)
            print(f"N={N} H={H} B={B} persistent={persist} npb={npb}: {ms*10:.2f} us per iteration  ({B*N/(ms*10e-6)/1e9:.1f} G pose*point/s, "
                  f"{24*B*N/(ms*10e-6)/1e9/6548.2:.3f} of HBM roofline)", flush=True)
    _lib.set_option("PERSIST", -1); _lib.set_option("RF_NPB", -1)

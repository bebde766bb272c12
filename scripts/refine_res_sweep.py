"""Fused refinement: resident points on/off (option RF_RES) x texel table (SMALL_TABLE: U8Q companion vs F16D) x cloud size.  usage: python scripts/refine_res_sweep.py [N ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
Ns = [int(a) for a in sys.argv[1:]] or [1_000_000]
for N in Ns:
    H = 1024
    sc = synth.make_scene(N, H, 2 * H, seed=3)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    rng = np.random.default_rng(0)
    starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
    ref = engine.Refiner(6, 0.1, 0.8, 5, True)
    def run():
        ref.reset(starts); ref.run(cloud, image, 100)
    for res, small in ((1, 1), (0, 1), (1, 0), (0, 0)):
        _lib.set_option("RF_RES", res); _lib.set_option("SMALL_TABLE", small)
        ms = timeit(run, iters=3, warm=1)
        print(f"N={N} resident points={res} compact table={small}: {ms*10:.2f} us/iter  ({24*6*N/(ms*10e-6)/1e9/6548.8:.3f} of HBM roofline)", flush=True)
    _lib.set_option("RF_RES", -1); _lib.set_option("SMALL_TABLE", -1)

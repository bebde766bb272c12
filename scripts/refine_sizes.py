"""Fused refinement (B = 6, 100 iterations) over cloud sizes with the default settings.  usage: python scripts/refine_sizes.py [H] N [N ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
H = int(sys.argv[1]); Ns = [int(a) for a in sys.argv[2:]]
for N in Ns:
    sc = synth.make_scene(N, H, 2 * H, seed=3)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    rng = np.random.default_rng(0)
    starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
    ref = engine.Refiner(6, 0.1, 0.8, 5, True)
    def run():
        ref.reset(starts); ref.run(cloud, image, 100)
    for res in (1, 0):
        _lib.set_option("RF_RES", res)
        ms = timeit(run, iters=3, warm=1)
        print(f"{os.environ.get('PCL_LIB', 'default')} N={N} H={H} resident={res}: {ms*10:.2f} us/iter  ({24*6*N/(ms*10e-6)/1e9/6548.8:.3f} of HBM roofline)", flush=True)
    _lib.set_option("RF_RES", -1)
    del cloud, image, xyz, rgb, img

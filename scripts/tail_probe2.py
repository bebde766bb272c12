"""Per-iteration cost of the fused refinement vs cloud size: the fixed part (barrier + finalize) shows on tiny clouds."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
for n in (4096, 262144, 1_000_000):
    sc = synth.make_scene(n, 1024, 2048, seed=3)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    rng = np.random.default_rng(0)
    starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
    for B in (1, 6):
        ref = engine.Refiner(B, 0.1, 0.8, 5, True)
        def run():
            ref.reset(starts[:B]); ref.run(cloud, image, 100)
        for persist in (1, 0):
            _lib.set_option("PERSIST", persist)
            ms = timeit(run, iters=3, warm=1)
            print(f"N={n} B={B}: {ms*10:.2f} us per iteration (persistent={persist})", flush=True)
        _lib.set_option("PERSIST", -1)

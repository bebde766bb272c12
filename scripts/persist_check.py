"""Persistent refinement vs per-iteration launches: same trajectory?  (early iterations must agree to fp32 rounding)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(300_000, 512, 1024, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
rng = np.random.default_rng(0)
starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.2, 3)]) for _ in range(16)]).astype(np.float32)).to(dev)
for B in (1, 2, 3, 4, 5, 6, 7, 11, 16):
    for batch in (True, False):
        res = {}
        for persist in ("0", "1"):
            _lib.set_option("PERSIST", int(persist))
            out = []
            for iters in (3, 40):
                ref = engine.Refiner(B, 0.1, 0.8, 5, batch).reset(starts[:B]).run(cloud, image, iters)
                o = ref.read()
                out.append((o["pose"].cpu().numpy(), o["loss"].cpu().numpy(), o["param"].cpu().numpy()))
            # split run: 2 + 38 iterations must equal 40 in one go
            ref = engine.Refiner(B, 0.1, 0.8, 5, batch).reset(starts[:B]).run(cloud, image, 2).run(cloud, image, 38)
            o = ref.read()
            out.append((o["pose"].cpu().numpy(), o["loss"].cpu().numpy(), o["param"].cpu().numpy()))
            res[persist] = out
        d3 = np.abs(res["0"][0][0] - res["1"][0][0]).max()
        d40 = np.abs(res["0"][1][0] - res["1"][1][0]).max()
        split = np.abs(res["1"][1][0] - res["1"][2][0]).max()
        l40 = np.abs(res["0"][1][1] - res["1"][1][1]).max()
        print(f"B={B} batch={batch}: |pose diff| after 3 it {d3:.2e}, after 40 it {d40:.2e} (loss diff {l40:.2e}); persistent 2+38 vs 40: {split:.2e}", flush=True)

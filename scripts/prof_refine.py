"""ncu target: steady-state refinement launches (B=6, C2 sizes)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
rng = np.random.default_rng(0)
cand = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
ref = engine.Refiner(6, 0.1, 0.8, 5, True).reset(cand)
ref.run(cloud, image, 30)
torch.cuda.synchronize()
print("done")

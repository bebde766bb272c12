"""Short GPU run for ncu: 2x scoring (P poses), 2x fused fwd+bwd (B=64), 2x fused fwd+bwd (B=6)
on the C2-sized scene (N=1M, 1024x2048).  Usage: python scripts/prof_run.py [N] [H] [P] [fmt]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
P = int(sys.argv[3]) if len(sys.argv) > 3 else 256
fmt = sys.argv[4] if len(sys.argv) > 4 else "auto"
dev = torch.device("cuda:0")
sc = synth.make_scene(N, H, 2 * H, seed=2)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
poses = torch.from_numpy(synth.pose_grid(sc.room, (8, 8, 2), max(1, P // 128))[:P]).to(dev)
rng = np.random.default_rng(0)
cand = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(64)]).astype(np.float32)).to(dev)
cloud = engine.Cloud(xyz, rgb)
image = engine.Image(img, fmt)
for _ in range(2):
    engine.score(cloud, image, poses)
for _ in range(2):
    engine.loss_fwd_bwd(cloud, image, cand)
for _ in range(2):
    engine.loss_fwd_bwd(cloud, image, cand[:6])
torch.cuda.synchronize()
print("done")

"""Where the time of color_match goes at C4 sizes (5 M points, 1024x2048): python scripts/color_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import synth
from piccolo_b200.color_utils import color_match
dev = torch.device("cuda:0")
room = (8.0, 6.0, 3.0)
xyz, rgb8 = synth.sample_room_points(5_000_000, room, seed=2)
gt = synth.random_gt_pose(room, seed=101, yaw_only=True)
pano8 = synth.perturb_panorama(synth.render_panorama(gt, 1024, 2048, room), seed=3, gamma=1.1, wb=(1.0, 0.97, 1.02), retexture_frac=0.1)
img, rgb = torch.from_numpy(synth.img_from_u8(pano8)).to(dev), torch.from_numpy(synth.rgb_from_u8(rgb8)).to(dev)
for _ in range(3):
    color_match(img, rgb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    out = color_match(img, rgb)
torch.cuda.synchronize()
print(f"color_match on device: {(time.perf_counter() - t0) * 100:.2f} ms per call")
t0 = time.perf_counter()
from oracle.color_oracle import color_match_np
host = color_match_np(img.cpu(), rgb.cpu())
print(f"color_match CPU restatement: {(time.perf_counter() - t0) * 1e3:.1f} ms per call; identical: {bool(torch.equal(host, out.cpu()))}")

"""Short GPU run for ncu: structured-grid scoring of the C2 start grid (75 x 24) — python scripts/prof_grid.py [fmt]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth  # noqa: E402
from piccolo_b200.utils import generate_rot_points  # noqa: E402

fmt = sys.argv[1] if len(sys.argv) > 1 else "auto"
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img, fmt)
trans = torch.from_numpy(np.ascontiguousarray(synth.pose_grid(sc.room, (5, 5, 3), 1)[:, :3])).to(dev)
rot = generate_rot_points({"yaw_only": False, "num_yaw": 4, "num_pitch": 4, "num_roll": 4}).to(dev)
for _ in range(3):
    engine.score_grid(cloud, image, trans, rot)
torch.cuda.synchronize()
print("done")

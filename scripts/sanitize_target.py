"""Small run that touches every kernel of the library (all texel formats, scoring, fwd+bwd, refinement with PDL,
re-rank, top-k, ragged cloud sizes) — target for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(7001, 64, 128, seed=2)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
rng = np.random.default_rng(0)
poses = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.2, 3)]) for _ in range(70)]).astype(np.float32)).to(dev)
for order in (0, 1):
    cloud = engine.Cloud(xyz, rgb, 0.05, order)
    for fmt in ("auto", "f16d", "u8q", "u8p", "tex", "f32"):
        image = engine.Image(img, fmt)
        loss, cnt = engine.score(cloud, image, poses)
        from piccolo_b200.utils import generate_rot_points
        for rot in (generate_rot_points({"yaw_only": True, "num_yaw": 8}), generate_rot_points({"yaw_only": False, "num_yaw": 4, "num_pitch": 4, "num_roll": 4}),
                    poses[:40, 3:].cpu()):                              # one group | 6 groups of 4 | R > 32: expanded + generic kernel
            gl, gc = engine.score_grid(cloud, image, poses[:7, :3].contiguous(), rot.to(dev))
        l2, c2, g = engine.loss_fwd_bwd(cloud, image, poses[:40])
        l3, c3, g3 = engine.loss_fwd_bwd(cloud, image, poses[:6])
        out = engine.Refiner(6, 0.1, 0.8, 5, True).reset(poses[:6]).run(cloud, image, 4).read()
    idx = engine.topk(loss, 10)
    s = engine.hist_rerank(cloud, img, poses[:12], 4, 4)
from piccolo_b200.color_utils import color_match, color_mod
cm = color_match(img, rgb)
mi, mr = color_mod(img, rgb, 256)
torch.cuda.synchronize()
print("sanitize target ok", float(loss.min()), idx[:3].tolist(), float(s.max()))

import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth, pipeline
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
for seed in (2, 3, 4):
    sc = synth.make_scene(N, H, 2 * H, seed=seed)
    grid = bench.stanford_grid(sc, dev)
    cloud = engine.Cloud(torch.from_numpy(sc.xyz).to(dev), torch.from_numpy(sc.rgb).to(dev))
    image = engine.Image(torch.from_numpy(sc.img).to(dev))
    loss, _ = engine.score(cloud, image, grid)
    idx = engine.topk(loss, 50)
    starts = grid[idx]
    ref = engine.Refiner(50, 0.1, 0.8, 5, True).reset(starts).run(cloud, image, 100).read()
    pose = ref["pose"].cpu().numpy(); fl = ref["loss"].cpu().numpy()
    terr = np.linalg.norm(pose[:, :3] - sc.gt_pose[:3], axis=1)
    order = np.argsort(fl)
    print("seed", seed, "gt", np.round(sc.gt_pose, 3))
    print(" rank-by-score of best-final candidates:", order[:8], "final loss", np.round(fl[order[:8]], 4), "t_err", np.round(terr[order[:8]], 3))
    print(" start losses top8", np.round(loss[idx[:8]].cpu().numpy(), 4), " t_err of first 6 after refine", np.round(terr[:6], 3))

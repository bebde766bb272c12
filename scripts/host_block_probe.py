"""Which call of one query blocks the host?  Host time of each step, no synchronisation in between."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from piccolo_b200 import engine, pipeline, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
grid = bench.stanford_grid(sc, dev)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
cfg = pipeline.STANFORD_PARALLEL
for it in range(3):
    torch.cuda.synchronize()
    T = [time.perf_counter()]
    def m(): T.append(time.perf_counter())
    loss = engine.score_grid(cloud, image, grid.trans, grid.rot)[0]; m()
    idx = engine.topk(loss, 50); m()
    mid = grid.index_select(0, idx); m()
    scores = engine.hist_rerank(cloud, img, mid, 4, 4); m()
    keep = engine.topk(-scores, 6); idx2 = idx.index_select(0, keep); starts = grid.index_select(0, idx2); m()
    ref = engine.Refiner(6, cfg.lr, cfg.factor, cfg.patience, True); m()
    ref.reset(starts); m()
    ref.run(cloud, image, 100); m()
    out = ref.read(); best = out["loss"].argmin(); m()
    torch.cuda.synchronize(); m()
    names = ["score_grid", "topk50", "index_select", "hist_rerank", "topk6+select", "Refiner()", "reset", "run", "read", "sync"]
    print(" | ".join(f"{n} {1e3*(b-a):.2f}" for n, a, b in zip(names, T[:-1], T[1:])))

"""Where the end-to-end (host buffers in, pose out) time of one C2 query goes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from piccolo_b200 import engine, pipeline, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
grid = bench.stanford_grid(sc, dev)
xyz_h, rgb_h, img_h, grid_h = [torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)] + [grid.cpu().pin_memory()]
cfg = pipeline.STANFORD_PARALLEL
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = T(); xyz, rgb, img, g = [a.to(dev, non_blocking=True) for a in (xyz_h, rgb_h, img_h, grid_h)]
    t1 = T(); cloud = engine.Cloud(xyz, rgb, 0.05)
    t2 = T(); image = engine.Image(img)
    t3 = T(); loss, _ = engine.score(cloud, image, g)
    t4 = T(); idx = engine.topk(loss, 50); mid = g.index_select(0, idx); sc_ = engine.hist_rerank(cloud, img, mid); keep = engine.topk(-sc_, 6); starts = mid.index_select(0, keep)
    t5 = T(); ref = engine.Refiner(6, 0.1, 0.8, 5, True).reset(starts); ref.run(cloud, image, 100); out = ref.read()
    t6 = T(); res = torch.cat([out["pose"][out["loss"].argmin()], out["loss"].min().reshape(1)]).cpu()
    t7 = T()
    print(f"it{it}: h2d {1e3*(t1-t0):.2f} cloud {1e3*(t2-t1):.2f} image {1e3*(t3-t2):.2f} score {1e3*(t4-t3):.2f} rerank {1e3*(t5-t4):.2f} refine {1e3*(t6-t5):.2f} d2h {1e3*(t7-t6):.2f} total {1e3*(t7-t0):.2f} ms")
t0 = T()
for _ in range(5): pipeline.localize_query_host(xyz_h, rgb_h, img_h, grid_h, cfg, dev)
print(f"localize_query_host: {1e3*(T()-t0)/5:.2f} ms/query")

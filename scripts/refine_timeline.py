"""Per-iteration timeline of the fused refinement inside a bench-like query (L2 flushed, scoring and re-rank first) vs
back-to-back warm runs: where does the in-query refinement lose time against the steady-state probe?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from piccolo_b200 import _lib, engine, pipeline, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
grid = bench.stanford_grid(sc, dev)
cfg = pipeline.STANFORD_PARALLEL
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_lib.set_option("RF_DEBUG", 1)
def query(flush_l2):
    if flush_l2: flush.fill_(1)
    loss = engine.score_grid(cloud, image, grid.trans, grid.rot)[0]
    idx = engine.topk(loss, 50); mid = grid.index_select(0, idx)
    scores = engine.hist_rerank(cloud, img, mid, 4, 4)
    keep = engine.topk(-scores, 6); starts = grid.index_select(0, idx.index_select(0, keep))
    ref = engine.Refiner(6, cfg.lr, cfg.factor, cfg.patience, True).reset(starts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ref.run(cloud, image, 100); e1.record(); torch.cuda.synchronize()
    return ref, starts, e0.elapsed_time(e1)
for _ in range(2): query(True)
for name, fl in (("in a query, L2 flushed first", True), ("in a query, no flush", False)):
    ref, starts, ms = query(fl)
    t = ref.debug_timeline().astype(np.int64); d = np.diff(t) / 1e3
    print(f"{name}: event window {ms*1e3:.0f} us; first stamp to last {(t[-1]-t[0])/1e3:.0f} us; iteration times us: first 5 {np.round(d[:5],1).tolist()}  "
          f"median {np.median(d):.2f}  mean {d.mean():.2f}  last 5 {np.round(d[-5:],1).tolist()}  => launch+prologue+first iteration {ms*1e3 - (t[-1]-t[0])/1e3:.0f} us")
ref = engine.Refiner(6, cfg.lr, cfg.factor, cfg.patience, True)
for rep in range(3):
    ref.reset(starts); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ref.run(cloud, image, 100); e1.record(); torch.cuda.synchronize()
    t = ref.debug_timeline().astype(np.int64); d = np.diff(t) / 1e3
    print(f"back-to-back run {rep}: event window {e0.elapsed_time(e1)*1e3:.0f} us; iteration median {np.median(d):.2f} mean {d.mean():.2f}; launch+prologue+first iteration {e0.elapsed_time(e1)*1e3 - (t[-1]-t[0])/1e3:.0f} us")

"""Host-side timeline of pipeline.localize_stream at C2 sizes: how long do launch / stage / read block the host?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from piccolo_b200 import engine, pipeline, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
grid = bench.stanford_grid(sc, dev)
xyz_h, rgb_h, img_h = [torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)]
grid_h = pipeline.StartGrid(grid.trans.cpu(), grid.rot.cpu()).pin_memory()
cfg = pipeline.STANFORD_PARALLEL
q = (xyz_h, rgb_h, img_h, grid_h)
for _ in range(2): pipeline.localize_query_host(*q, cfg, dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): pipeline.localize_query_host(*q, cfg, dev)
torch.cuda.synchronize(); print(f"one at a time: {(time.perf_counter()-t0)/5*1e3:.2f} ms/query")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = []
    for r in pipeline.localize_stream((q for _ in range(8)), cfg, dev):
        ts.append(time.perf_counter())
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"stream: {(t1-t0)/8*1e3:.2f} ms/query; per-result intervals ms:", [f"{(b-a)*1e3:.2f}" for a, b in zip([t0] + ts[:-1], ts)])

# the same loop with the three host-side pieces timed
main, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
def stage(q):
    with torch.cuda.stream(side):
        xyz = q[0].to(dev, non_blocking=True); rgb = q[1].to(dev, non_blocking=True)
        cloud = engine.Cloud(xyz, rgb, cfg.out_of_room_quantile)
        img = q[2].to(dev, non_blocking=True); grid = q[3].to(dev, non_blocking=True)
        image = engine.Image(img)
        ev = torch.cuda.Event(); ev.record(side)
    return cloud, image, img, grid, ev
def launch(st):
    cloud, image, img, grid, ev = st
    main.wait_event(ev)
    for t in (img, grid.trans, grid.rot): t.record_stream(main)
    out = pipeline.localize_query(cloud, image, grid, cfg, img=img)
    return torch.cat([out["pose"], out["loss"].reshape(1)])
cur = stage(q)
for i in range(8):
    t0 = time.perf_counter(); r = launch(cur)
    t1 = time.perf_counter(); nxt = stage(q)
    t2 = time.perf_counter(); side.synchronize()
    t3 = time.perf_counter(); res = r.cpu()
    t4 = time.perf_counter(); cur = nxt
    print(f"q{i}: launch {1e3*(t1-t0):.2f}  stage(host) {1e3*(t2-t1):.2f}  side stream done after +{1e3*(t3-t2):.2f}  result after +{1e3*(t4-t3):.2f} ms")

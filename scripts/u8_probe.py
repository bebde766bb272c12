"""Host-buffer entries with uint8 colours / panorama vs float32 (C2 sizes): one at a time and as a stream."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from piccolo_b200 import pipeline, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
grid = bench.stanford_grid(sc, dev)
grid_h = pipeline.StartGrid(grid.trans.cpu(), grid.rot.cpu()).pin_memory()
cfg = pipeline.STANFORD_PARALLEL
f32 = tuple(torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)) + (grid_h,)
u8 = (f32[0], torch.from_numpy(np.rint(sc.rgb * 255).astype(np.uint8)).pin_memory(), torch.from_numpy(np.rint(sc.img * 255).astype(np.uint8)).pin_memory(), grid_h)
for name, q in (("float32 host buffers (49 MB)", f32), ("uint8 colours + panorama (21 MB)", u8)):
    for _ in range(3): pipeline.localize_query_host(*q, cfg, dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): r = pipeline.localize_query_host(*q, cfg, dev)
    torch.cuda.synchronize(); one = (time.perf_counter() - t0) / 8
    for _ in pipeline.localize_stream((q for _ in range(6)), cfg, dev): pass
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in pipeline.localize_stream((q for _ in range(12)), cfg, dev): pass
    torch.cuda.synchronize(); st = (time.perf_counter() - t0) / 12
    print(f"{name}: one at a time {one*1e3:.2f} ms/query, stream {st*1e3:.2f} ms/query, pose {np.round(r[0].numpy(), 4).tolist()}")

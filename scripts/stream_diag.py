"""Reproduce the slow first stream after one-at-a-time queries: per-result intervals + allocator statistics."""
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
q = tuple(torch.from_numpy(a).pin_memory() for a in (sc.xyz, sc.rgb, sc.img)) + (grid_h,)
def stats(tag):
    s = torch.cuda.memory_stats(dev)
    free, total = torch.cuda.mem_get_info(dev)
    print(f"  [{tag}] torch reserved {s['reserved_bytes.all.current']>>20} MB, cudaMalloc calls {s['num_device_alloc']}, cudaFree calls {s['num_device_free']}, device used {(total-free)>>20} MB", flush=True)
for _ in range(11): pipeline.localize_query_host(*q, cfg, dev)
torch.cuda.synchronize(); stats("after 11 one-at-a-time queries")
for rep in range(3):
    t0 = time.perf_counter(); ts = [t0]
    for r in pipeline.localize_stream((q for _ in range(9)), cfg, dev): ts.append(time.perf_counter())
    torch.cuda.synchronize()
    print(f"stream {rep}: intervals ms", [f"{1e3*(b-a):.1f}" for a, b in zip(ts[:-1], ts[1:])], flush=True)
    stats(f"after stream {rep}")

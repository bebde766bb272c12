"""Look for intermittent stalls in pipeline.localize_stream: many streams of queries, per-result host intervals."""
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
for _ in range(3): pipeline.localize_query_host(*q, cfg, dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
allv = []
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = [t0]
    for r in pipeline.localize_stream((q for _ in range(10)), cfg, dev):
        ts.append(time.perf_counter())
    iv = [1e3 * (b - a) for a, b in zip(ts[:-1], ts[1:])]
    allv.append(iv)
    if max(iv[1:]) > 13 or iv[0] > 20 or rep < 2:
        print(f"rep {rep}: mean {np.mean(iv):.2f} ms; intervals", [f"{v:.1f}" for v in iv], flush=True)
a = np.array(allv)
print(f"{reps} streams of 10: steady intervals median {np.median(a[:,1:]):.2f} p99 {np.percentile(a[:,1:],99):.2f} max {a[:,1:].max():.2f}; first interval median {np.median(a[:,0]):.2f} max {a[:,0].max():.2f}")

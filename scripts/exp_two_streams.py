"""Experiment: refine 6 candidates as ONE chain of launches vs TWO chains of 3 candidates on two streams
(each chain's serial tail — fence, last-CTA reduction, Adam, launch gap — overlaps the other chain's compute)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import engine, synth, pipeline
import bench
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
grid = bench.stanford_grid(sc, dev)
out = pipeline.localize_query(cloud, image, grid, pipeline.STANFORD_PARALLEL, img=img)
starts = grid.index_select(0, out["start_index"])
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def one_chain():
    r = engine.Refiner(6, 0.1, 0.8, 5, True).reset(starts); r.run(cloud, image, 100); return r.read()["pose"]
def two_chains():
    main = torch.cuda.current_stream(dev)
    s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s1):
        a = engine.Refiner(3, 0.1, 0.8, 5, True).reset(starts[:3])
    with torch.cuda.stream(s2):
        b = engine.Refiner(3, 0.1, 0.8, 5, True).reset(starts[3:])
    for _ in range(10):                      # interleave the enqueueing so neither stream starves
        with torch.cuda.stream(s1):
            a.run(cloud, image, 10)
        with torch.cuda.stream(s2):
            b.run(cloud, image, 10)
    with torch.cuda.stream(s1):
        pa = a.read()["pose"]
    with torch.cuda.stream(s2):
        pb = b.read()["pose"]
    main.wait_stream(s1); main.wait_stream(s2)
    return torch.cat([pa, pb])
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r
t1, p1 = timeit(one_chain); t2, p2 = timeit(two_chains)
print(f"one chain of 6: {t1:.3f} ms ({t1*10:.1f} us/iter);  two chains of 3 on two streams: {t2:.3f} ms ({t2*10:.1f} us/iter-equivalent)")
print("max pose difference:", float((p1 - p2).abs().max()))

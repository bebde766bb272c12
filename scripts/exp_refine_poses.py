import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import engine, synth, pipeline
from scripts.perf_probe import timeit
import bench
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img, os.environ.get("PROBE_FMT", "auto"))
grid = bench.stanford_grid(sc, dev)
out = pipeline.localize_query(cloud, image, grid, pipeline.STANFORD_PARALLEL, img=img)
starts = grid.index_select(0, out["start_index"])
rng = np.random.default_rng(0)
near = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
for name, s in (("near-GT starts", near), ("bench starts (grid top-6 after re-rank)", starts)):
    ref = engine.Refiner(6, 0.1, 0.8, 5, True)
    def run():
        ref.reset(s); ref.run(cloud, image, 100)
    ms = timeit(run, iters=3, warm=1)
    one = timeit(lambda: engine.loss_fwd_bwd(cloud, image, s), iters=20)
    print(f"{name}: refine {ms*10:.1f} us/iter; single fwd+bwd launch {one*1e3:.1f} us; poses {s[:2].cpu().numpy().round(2).tolist()}")
    fin = ref.read()["pose"]
    one = timeit(lambda: engine.loss_fwd_bwd(cloud, image, fin), iters=20)
    print(f"   at the refined poses: single launch {one*1e3:.1f} us")

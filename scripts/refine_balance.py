"""Fused refinement at C2 sizes: timing + per-CTA cycle statistics (option RF_DEBUG): how evenly do the SMs run?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sc = synth.make_scene(N, H, 2 * H, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
rng = np.random.default_rng(0)
starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
ref = engine.Refiner(6, 0.1, 0.8, 5, True)
def run():
    ref.reset(starts); ref.run(cloud, image, 100)
_lib.set_option("RF_DEBUG", 0)
ms = timeit(run, iters=3, warm=1)
_lib.set_option("RF_DEBUG", 1)
run(); torch.cuda.synchronize()
st = ref.debug_stats().astype(np.float64)          # (ctas + 1, 4), last row = service CTA
sv, st = st[-1], st[:-1]
busy, wait = st[:, 0] / 100, st[:, 1] / 100
print(f"service CTA per iteration: waiting for records {sv[0]/100:.0f} cycles, reducing + stepping {sv[1]/100:.0f} cycles (of which reducing the records {sv[2]/100:.0f}) | compute warp 0: phases with prefetched "
      f"poses {st[:,2].mean():.0f} of 200 (min {st[:,2].min():.0f}))")
print(f"{ms*10:.2f} us/iter | per-CTA busy cycles/iter: mean {busy.mean():.0f} min {busy.min():.0f} p5 {np.percentile(busy,5):.0f} p50 {np.median(busy):.0f} "
      f"p95 {np.percentile(busy,95):.0f} max {busy.max():.0f} | wait cycles/iter: mean {wait.mean():.0f} min {wait.min():.0f} max {wait.max():.0f}", flush=True)
order = np.argsort(-busy)
print("slowest CTAs:", [(int(c), int(busy[c])) for c in order[:8]], " fastest:", [(int(c), int(busy[c])) for c in order[::-1][:8]], flush=True)
np.save("gpurun_out/r2_refine_cta_cycles.npy", st)
_lib.set_option("RF_DEBUG", -1)

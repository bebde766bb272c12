"""Config C5: scaling sweep, N points x P poses, forward-only scoring and fused forward+backward, on one GPU
(run under torchrun for more: every rank sweeps its own copy, i.e. weak scaling).  Writes a markdown table.
    python scripts/sweep_c5.py [out.md]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth
from scripts.perf_probe import timeit

out_path = sys.argv[1] if len(sys.argv) > 1 else None
dev = torch.device("cuda:0")
peak = 6548.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
lines = ["# C5 sweep on one B200: pose·point evaluations/s and fraction of the HBM roofline (24 B per evaluation, "
         f"{peak:.0f} GB/s measured)\n", "Best of 3 x (CUDA-event average over the launches); 1024x2048 panorama (F16D table), Morton-ordered cloud.\n",
         "| N points | P poses | scoring G pp/s | frac | fwd+bwd B=P G pp/s | frac |", "|---|---|---|---|---|---|"]
sc_img = None
for N in (1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000):
    t0 = time.time()
    sc = synth.make_scene(N, 1024, 2048, seed=3)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    del xyz, rgb
    rng = np.random.default_rng(0)
    for P in (64, 256, 1024, 4096, 8192):
        poses = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.4, 3)]) for _ in range(P)]).astype(np.float32)).to(dev)
        iters = max(1, min(10, int(2e11 / (N * P))))
        ms = timeit(lambda: engine.score(cloud, image, poses), iters=iters, warm=1, repeats=3)
        g_f = N * P / ms / 1e6
        if N * P <= 2.1e11:
            ms_b = timeit(lambda: engine.loss_fwd_bwd(cloud, image, poses), iters=iters, warm=1, repeats=3)
            g_b = N * P / ms_b / 1e6
            lines.append(f"| {N/1e6:.0f} M | {P} | {g_f:.1f} | {24*g_f/peak:.2f} | {g_b:.1f} | {24*g_b/peak:.2f} |")
        else:
            lines.append(f"| {N/1e6:.0f} M | {P} | {g_f:.1f} | {24*g_f/peak:.2f} | – | – |")
        print(lines[-1], flush=True)
    del cloud, image
    torch.cuda.empty_cache()
text = "\n".join(lines) + "\n"
if out_path:
    open(out_path, "w").write(text)

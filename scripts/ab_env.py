"""Time the refine loop / fwd+bwd under different env settings in one GPU session:  python scripts/ab_env.py "K=V ..." "K=V ..." """
import os, subprocess, sys
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
rng = np.random.default_rng(0)
cand = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(64)]).astype(np.float32)).to(dev)
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
import bench
grid = bench.stanford_grid(sc, dev)
ref = engine.Refiner(6, 0.1, 0.8, 5, True).reset(cand[:6])
out = []
for rep in range(2):
    b6 = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand[:6]), iters=20)
    b = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand), iters=10)
    c = timeit(lambda: ref.run(cloud, image, 100), iters=3, warm=1)
    sc_ms = timeit(lambda: engine.score(cloud, image, grid), iters=3)
    out.append(f"score1800 {sc_ms:.3f} ms  bwd6 {b6*1e3:.1f} us  bwd64 {64e6/b/1e6:.1f} G  refine {c*10:.1f} us/iter")
print(os.environ.get("AB_TAG", ""), " | ".join(out))
'''
for setting in sys.argv[1:]:
    env = dict(os.environ, AB_TAG=setting)
    for kv in setting.split():
        k, v = kv.split("=")
        env[k] = v
    subprocess.run([sys.executable, "-c", code], env=env)

"""A/B timing of two builds of the library in one GPU session (each in its own subprocess):
   python scripts/ab_probe.py libA.so libB.so"""
import os, subprocess, sys
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import engine, synth
from scripts.perf_probe import timeit
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
poses = torch.from_numpy(synth.pose_grid(sc.room, (8, 8, 2), 4)[:512]).to(dev)
rng = np.random.default_rng(0)
cand = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(64)]).astype(np.float32)).to(dev)
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
ref = engine.Refiner(6, 0.1, 0.8, 5, True).reset(cand[:6])
out = []
for rep in range(3):
    a = timeit(lambda: engine.score(cloud, image, poses)); b = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand), iters=10)
    c = timeit(lambda: ref.run(cloud, image, 100), iters=3, warm=1)
    out.append(f"score {512e6/a/1e6:.1f} G  bwd64 {64e6/b/1e6:.1f} G  refine {c*10:.1f} us/iter")
print(os.environ.get("PCL_LIB", "default"), os.environ.get("PCL_NT_BWD", ""), " | ".join(out))
'''
for lib in sys.argv[1:]:
    for nt in ("256",):
        env = dict(os.environ, PCL_LIB=os.path.abspath(lib), PCL_NT_BWD=nt)
        subprocess.run([sys.executable, "-c", code], env=env)

"""A/B the fused refinement of several builds of the library in one GPU session (each in its own subprocess):
   python scripts/ab_refine.py libA.so libB.so ...   ('default' = the in-tree build)"""
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
ref = engine.Refiner(6, 0.1, 0.8, 5, True)
def run():
    ref.reset(cand[:6]); ref.run(cloud, image, 100)
out = []
for rep in range(3):
    c = timeit(run, iters=3, warm=1)
    b = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand), iters=10)
    out.append(f"refine {c*10:.2f} us/iter  bwd64 {64e6/b/1e6:.1f} G")
print(os.environ.get("PCL_LIB", "default"), " | ".join(out), flush=True)
'''
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["PCL_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", code], env=env)

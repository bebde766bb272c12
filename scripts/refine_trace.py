"""Fused refinement at C2 sizes with a TRACE build (scripts/build_variant.sh trace "-DPCL_RF_TRACE"; PCL_LIB=build/trace.so):
where does a warp's time inside a phase go — full groups, remainder + flush, waiting at the phase barrier?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from piccolo_b200 import _lib, engine, synth
dev = torch.device("cuda:0")
sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
rng = np.random.default_rng(0)
starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
ref = engine.Refiner(6, 0.1, 0.8, 5, True)
_lib.set_option("RF_DEBUG", 1)
for _ in range(2):
    ref.reset(starts); ref.run(cloud, image, 100); torch.cuda.synchronize()
st = ref.debug_stats().astype(np.float64)[:-1] / 200          # per phase
print("per phase, cycles (mean over the warps of a CTA; then mean / min / max over the 147 CTAs):")
for k, name in enumerate(("full groups", "remainder + flush", "barrier wait (mean warp)", "barrier wait (worst warp)")):
    print(f"  {name:28s} mean {st[:,k].mean():8.0f}  min {st[:,k].min():8.0f}  max {st[:,k].max():8.0f}")
print("  phase total (mean warp)      ", f"{(st[:,0]+st[:,1]+st[:,2]).mean():8.0f}")

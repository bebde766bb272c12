"""Timing probe (GPU): sweeps texel format / points-per-thread / pose-block for scoring and fwd+bwd.
Usage: python scripts/perf_probe.py [N] [H] [P]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import engine, synth  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    P = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    dev = torch.device("cuda:0")
    t0 = time.time()
    sc = synth.make_scene(N, H, 2 * H, seed=2)
    print(f"scene N={N} {H}x{2*H} built in {time.time()-t0:.1f}s", flush=True)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    grid = synth.pose_grid(sc.room, (8, 8, 2), max(1, P // 128))[:P]
    poses = torch.from_numpy(grid).to(dev)
    rng = np.random.default_rng(0)
    cand = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.1, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
    for order in (1, 0):
        cloud = engine.Cloud(xyz, rgb, 0.05, order)
        for fmt in ("u8q", "u8p", "f32"):
            image = engine.Image(img, fmt)
            for K in (2, 4, 8):
                os.environ["PCL_K"] = str(K)
                for PB in (8, 32):
                    os.environ["PCL_PB_FWD"] = str(PB)
                    ms = timeit(lambda: engine.score(cloud, image, poses))
                    print(f"order={order} fmt={fmt} K={K} PB={PB} SCORE P={len(poses)}: {ms:.3f} ms  {len(poses)*N/ms/1e6:.1f} G pp/s", flush=True)
                for PBB in (1, 6):
                    os.environ["PCL_PB_BWD"] = str(PBB)
                    ms = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand), iters=20, warm=3)
                    print(f"order={order} fmt={fmt} K={K} PBB={PBB} FWDBWD B=6: {ms:.4f} ms  {6*N/ms/1e6:.1f} G pp/s", flush=True)
            if order == 0:
                break
    # large-batch fwd+bwd (amortises launch + tail): 64 candidates
    os.environ["PCL_K"] = "4"; os.environ["PCL_PB_BWD"] = "8"
    cloud = engine.Cloud(xyz, rgb, 0.05, 1)
    image = engine.Image(img, "u8q")
    cand64 = cand.repeat(11, 1)[:64].contiguous()
    ms = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand64), iters=10)
    print(f"FWDBWD B=64 u8q K=4: {ms:.3f} ms {64*N/ms/1e6:.1f} G pp/s")
    ref = engine.Refiner(6, 0.1, 0.8, 5, True).reset(cand)
    ms = timeit(lambda: ref.run(cloud, image, 100), iters=3, warm=1)
    print(f"REFINE 100 iters B=6: {ms:.2f} ms  -> {ms/100*1000:.1f} us/iter, {600*N/ms/1e6:.1f} G pp/s")


if __name__ == "__main__":
    main()

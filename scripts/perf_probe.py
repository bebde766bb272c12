"""Timing probe (GPU): sweeps texel format / points-per-thread / pose-block for scoring and fwd+bwd.
Usage: python scripts/perf_probe.py [N] [H] [P]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piccolo_b200 import _lib, engine, synth  # noqa: E402


def timeit(fn, iters=5, warm=2, repeats=5):
    """best-of-`repeats` average over `iters` launches (min filters out clock / neighbour noise)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


FMTS = tuple(os.environ.get('PROBE_FMTS', 'u8q,tex,u8p,f32').split(','))


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
        for fmt in FMTS:
            image = engine.Image(img, fmt)
            ms = timeit(lambda: engine.score(cloud, image, poses))
            print(f"order={order} fmt={fmt} SCORE P={len(poses)}: {ms:.3f} ms  {len(poses)*N/ms/1e6:.1f} G pp/s", flush=True)
            ms = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand), iters=20, warm=3)
            print(f"order={order} fmt={fmt} FWDBWD B=6: {ms:.4f} ms  {6*N/ms/1e6:.1f} G pp/s", flush=True)
            c64 = cand.repeat(11, 1)[:64].contiguous()
            ms = timeit(lambda: engine.loss_fwd_bwd(cloud, image, c64), iters=10, warm=2)
            print(f"order={order} fmt={fmt} FWDBWD B=64: {ms:.4f} ms  {64*N/ms/1e6:.1f} G pp/s", flush=True)
            if order == 0:
                break
    # large-batch fwd+bwd (amortises launch + tail): 64 candidates
    cloud = engine.Cloud(xyz, rgb, 0.05, 1)
    image = engine.Image(img)
    cand64 = cand.repeat(11, 1)[:64].contiguous()
    ms = timeit(lambda: engine.loss_fwd_bwd(cloud, image, cand64), iters=10)
    print(f"FWDBWD B=64 auto: {ms:.3f} ms {64*N/ms/1e6:.1f} G pp/s")
    ref = engine.Refiner(6, 0.1, 0.8, 5, True).reset(cand)
    for pdl in ("1", "0"):
        _lib.set_option("PDL", int(pdl))
        ms = timeit(lambda: ref.run(cloud, image, 100), iters=3, warm=1)
        print(f"REFINE 100 iters B=6 pdl={pdl}: {ms:.2f} ms  -> {ms/100*1000:.1f} us/iter, {600*N/ms/1e6:.1f} G pp/s")


if __name__ == "__main__":
    main()

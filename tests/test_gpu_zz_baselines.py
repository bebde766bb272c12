"""Same-box baseline: the reference's own op chain — the oracle's ATen-chain port: einsum, atan2,
grid_sample, autograd, torch.optim.Adam + ReduceLROnPlateau — executed by PyTorch on the SAME B200, beside the CUDA
path, on a bounded sample of the C2 workload.  A reported comparison with a loose gate, not a parity test.
(Named to run last: it is a measurement.)"""
import time

import numpy as np
import pytest
import torch

from oracle import piccolo_oracle as orc
from piccolo_b200 import synth

pytestmark = pytest.mark.gpu


def test_reference_op_chain_on_the_same_gpu(capsys):
    import bench
    from piccolo_b200 import engine, pipeline
    dev = torch.device("cuda:0")
    cfg = pipeline.STANFORD_PARALLEL
    sc = synth.make_scene(1_000_000, 1024, 2048, seed=bench.SCENE_SEED)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    grid = bench.stanford_grid(sc, dev)
    poses = grid.poses()
    n, n_score, n_iter = xyz.shape[0], 100, 10

    def chain_score(k):
        with torch.no_grad():
            for i in range(k):
                orc.sampling_loss_torch(xyz, rgb, img, poses[i:i + 1])          # one pose per call, as the loop of utils.py:484-499

    def chain_refine(k):
        orc.refine_torch(xyz, rgb, img, poses[: cfg.num_input].clone(), lr=cfg.lr, num_iter=k, patience=cfg.patience, factor=cfg.factor,
                         q=cfg.out_of_room_quantile, batch_semantics=True)
    try:
        chain_score(3); chain_refine(2)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        chain_score(n_score)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        chain_refine(n_iter)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    except torch.cuda.OutOfMemoryError:                              # the op chain materialises ~0.5 KB per pose*point
        pytest.skip("the reference op chain does not fit this GPU's free memory")
    chain = {"score": n * n_score / (t1 - t0), "refine": n * n_iter * cfg.num_input / (t2 - t1),
             "query_s": (t1 - t0) / n_score * len(poses) + (t2 - t1) / n_iter * cfg.num_iter}

    cloud, image = engine.Cloud(xyz, rgb, cfg.out_of_room_quantile), engine.Image(img)
    pipeline.localize_query(cloud, image, grid, cfg, img=img)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        out = pipeline.localize_query(cloud, image, grid, cfg, img=img)
    torch.cuda.synchronize()
    ours_s = (time.perf_counter() - t0) / 5
    with capsys.disabled():
        print(f"\n[same-box baseline] reference op chain on cuda:0: scoring {chain['score'] / 1e9:.2f} G, refinement {chain['refine'] / 1e9:.3f} G pose*point/s, "
              f"{chain['query_s']:.2f} s per query (extrapolated); this library: {ours_s * 1e3:.1f} ms per query = {chain['query_s'] / ours_s:.0f}x")
    assert np.linalg.norm(out["pose"].cpu().numpy()[:3] - sc.gt_pose[:3]) < 0.05
    assert chain["query_s"] / ours_s > 20

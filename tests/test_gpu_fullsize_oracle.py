"""Oracle parity at the FULL size of BASELINE configs C2 and C3 (missing #4): the structured-grid scoring
kernel (translations x rotations, shared per-point work inside rotation groups) against the oracle's restatement of the
reference chain (oracle/piccolo_oracle.py: sampling_loss_torch = utils.py:16-103 + omniloc.py:198-200) in fp32 AND fp64.
The oracle is executed by torch on the GPU here only to finish in seconds at 1 M / 10 M points — it stays the checker.

Gates: every loss within 1e-4 of the fp64 oracle; the top-50 of the table identical to the top-50 of the fp64 ranking
wherever the fp64 gap between neighbours exceeds the measured fp32 noise (ours and the oracle's own fp32 run)."""
import numpy as np
import pytest
import torch

from oracle import piccolo_oracle as orc
from piccolo_b200 import synth

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-4


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def oracle_losses(sc, poses, dtype, chunk):
    xyz, rgb, img = cu(sc.xyz, dtype), cu(sc.rgb, dtype), cu(sc.img, dtype)
    p = cu(poses, dtype)
    out = []
    with torch.no_grad():
        for i in range(0, len(p), chunk):
            out.append(orc.sampling_loss_torch(xyz, rgb, img, p[i:i + chunk])[0].to(torch.float64))
    return torch.cat(out).cpu().numpy()


def adjudicate_topk(ours, l64, l32, k):
    """ours / l64 / l32: losses of the SAME pose list.  The top-k by `ours` must be the top-k by the fp64 oracle wherever the
    fp64 ordering is decided by more than the fp32 noise measured on this very table."""
    noise = max(np.abs(ours - l64).max(), np.abs(l32 - l64).max())
    mine = np.argsort(ours, kind="stable")[:k]
    ref = np.argsort(l64, kind="stable")[:k]
    cut = np.sort(l64)[k - 1]
    # (1) nothing outside my selection beats the fp64 cut by more than the noise, nothing inside loses to it by more
    outside = np.setdiff1d(np.arange(len(l64)), mine)
    assert (l64[mine] <= cut + 2 * noise).all(), (l64[mine].max(), cut, noise)
    assert (l64[outside] >= cut - 2 * noise).all(), (l64[outside].min(), cut, noise)
    # (2) where the fp64 gap at the cut is larger than the noise the SETS are identical
    gap = np.sort(l64)[k] - cut
    if gap > 2 * noise:
        assert set(mine.tolist()) == set(ref.tolist())
    # (3) my order follows the fp64 order up to the noise
    assert (np.diff(l64[mine]) >= -2 * noise).all()
    return noise, gap, len(set(mine.tolist()) & set(ref.tolist()))


def test_c2_full_grid_scoring_against_fp64_oracle():
    """C2: 1 M points, 1024x2048, the 75 x 24 Euler-lattice start grid (6 rotation groups of 4): ALL 1 800 poses."""
    from piccolo_b200 import engine
    import bench
    sc = synth.make_scene(1_000_000, 1024, 2048, seed=3)
    grid = bench.stanford_grid(sc, torch.device("cuda"))
    cloud, image = engine.Cloud(cu(sc.xyz), cu(sc.rgb)), engine.Image(cu(sc.img))
    ours = engine.score_grid(cloud, image, grid.trans, grid.rot)[0].cpu().numpy().astype(np.float64)
    per_pose = engine.score(cloud, image, grid.poses())[0].cpu().numpy().astype(np.float64)
    poses = grid.poses().cpu().numpy()
    l64 = oracle_losses(sc, poses, torch.float64, 8)
    l32 = oracle_losses(sc, poses, torch.float32, 8)
    assert np.isfinite(l64).all()
    np.testing.assert_allclose(ours, l64, rtol=LOSS_RTOL)
    np.testing.assert_allclose(per_pose, l64, rtol=LOSS_RTOL)
    noise, gap, common = adjudicate_topk(ours, l64, l32, 50)
    adjudicate_topk(per_pose, l64, l32, 50)
    print(f"\n[C2 full grid] max rel err vs fp64: structured {np.abs(ours / l64 - 1).max():.2e}, per-pose {np.abs(per_pose / l64 - 1).max():.2e}, "
          f"oracle fp32 {np.abs(l32 / l64 - 1).max():.2e}; top-50 common with fp64 {common}/50 (gap at the cut {gap:.2e}, noise {noise:.2e})")


def test_c3_grid_scoring_against_fp64_oracle():
    """C3: 10 M points, 2048x4096, 4096-pose grid (16 x 16 translations x 16 yaws, one rotation group).  The whole table
    through the kernel; the fp64 oracle on the 128 best poses of the table plus 64 seeded others."""
    from piccolo_b200 import engine, pipeline
    room = (40.0, 30.0, 3.0)
    sc = synth.make_scene(10_000_000, 2048, 4096, room=room, seed=5)
    g = cu(synth.pose_grid(room, (16, 16, 1), 16))
    grid = pipeline.StartGrid(g[::16, :3], g[:16, 3:])
    cloud, image = engine.Cloud(cu(sc.xyz), cu(sc.rgb)), engine.Image(cu(sc.img))
    table = engine.score_grid(cloud, image, grid.trans, grid.rot)[0].cpu().numpy().astype(np.float64)
    assert len(table) == 4096 and np.isfinite(table).all()
    order = np.argsort(table, kind="stable")
    rng = np.random.default_rng(7)
    sel = np.concatenate([order[:128], rng.choice(order[128:], 64, replace=False)])
    poses = grid.poses().cpu().numpy()[sel]
    l64 = oracle_losses(sc, poses, torch.float64, 2)
    l32 = oracle_losses(sc, poses, torch.float32, 2)
    np.testing.assert_allclose(table[sel], l64, rtol=LOSS_RTOL)
    # the 128 best of the table contain the top-50; the 64 others must stay behind the cut
    noise, gap, common = adjudicate_topk(table[sel], l64, l32, 50)
    print(f"\n[C3 grid] max rel err vs fp64 on {len(sel)} poses: {np.abs(table[sel] / l64 - 1).max():.2e} (oracle fp32: {np.abs(l32 / l64 - 1).max():.2e}); "
          f"top-50 common with fp64 {common}/50 (gap at the cut {gap:.2e}, noise {noise:.2e})")


def test_c4_perturbed_query_against_fp64_oracle():
    """C4: 5 M points, a colour-perturbed 1024x2048 query after `color_match` + uint8 re-quantisation (localize.py:402-404), the
    omniscenes-style yaw-only grid (13 x 13 translations at the z prior x 8 yaws = 1 352 poses): the 64 best poses of the table and 64
    seeded others against the fp64 oracle, then loss AND gradient of the refined pose."""
    from piccolo_b200 import engine, pipeline
    from piccolo_b200.color_utils import color_match, requantize
    sc = synth.make_scene(5_000_000, 1024, 2048, seed=3, yaw_only=True)
    gt = synth.random_gt_pose(sc.room, seed=41, yaw_only=True)
    gt[3] = np.round(gt[3] / (np.pi / 4)) * (np.pi / 4) + 0.1
    img8 = synth.perturb_panorama(synth.render_panorama(gt, 1024, 2048, sc.room), seed=1, gamma=1.1, wb=(1.0, 0.97, 1.03), retexture_frac=0.1)
    xyz, rgb = cu(sc.xyz), cu(sc.rgb)
    img = requantize(color_match(cu(synth.img_from_u8(img8)), rgb))
    sc_q = synth.Scene(sc.xyz, sc.rgb8, None, gt, sc.room)

    class Q:                                                   # what oracle_losses needs: the matched panorama
        pass
    q = Q(); q.xyz, q.rgb, q.img = sc.xyz, sc.rgb, img.cpu().numpy()
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    assert image.format == engine.IMAGE_F16D                   # still exact uint8/255 data
    g = cu(synth.pose_grid(sc.room, (13, 13, 1), 8))
    g[:, 2] = float(gt[2])
    grid = pipeline.StartGrid(g[::8, :3], g[:8, 3:])
    table = engine.score_grid(cloud, image, grid.trans, grid.rot)[0].cpu().numpy().astype(np.float64)
    order = np.argsort(table, kind="stable")
    rng = np.random.default_rng(9)
    sel = np.concatenate([order[:64], rng.choice(order[64:], 64, replace=False)])
    poses = grid.poses().cpu().numpy()[sel]
    l64 = oracle_losses(q, poses, torch.float64, 4)
    l32 = oracle_losses(q, poses, torch.float32, 4)
    np.testing.assert_allclose(table[sel], l64, rtol=LOSS_RTOL)
    noise, gap, common = adjudicate_topk(table[sel], l64, l32, 6)
    out = pipeline.localize_query(cloud, image, grid, pipeline.STANFORD._replace(parallel=True), img=img)
    pose = out["pose"].reshape(1, 6)
    l, c, gr = engine.loss_fwd_bwd(cloud, image, pose)
    p64 = cu(pose.cpu().numpy().astype(np.float64), torch.float64).requires_grad_()
    L = orc.sampling_loss_torch(cu(sc.xyz, torch.float64), cu(sc.rgb, torch.float64), img.to(torch.float64), p64)[0][0]
    L.backward()
    g64 = p64.grad[0].cpu().numpy()
    p32 = pose.clone().requires_grad_()
    orc.sampling_loss_torch(xyz, rgb, img, p32)[0][0].backward()
    g32 = p32.grad[0].cpu().numpy().astype(np.float64)
    assert abs(l.item() - float(L)) <= LOSS_RTOL * float(L)
    tol = max(1e-4 * np.abs(g64).max(), 3 * np.abs(g32 - g64).max())
    assert np.abs(gr[0].cpu().numpy() - g64).max() <= tol, (gr[0].cpu().numpy(), g64, tol)
    print(f"\n[C4 query] max rel err vs fp64 on {len(sel)} poses: {np.abs(table[sel] / l64 - 1).max():.2e}; top-6 common {common}/6; refined pose: loss rel err "
          f"{abs(l.item() / float(L) - 1):.2e}, |grad - grad64|max {np.abs(gr[0].cpu().numpy() - g64).max():.2e} (fp32 chain: {np.abs(g32 - g64).max():.2e}, "
          f"|grad64|max {np.abs(g64).max():.2e}); t error {np.linalg.norm(out['pose'][:3].cpu().numpy() - gt[:3]) * 100:.1f} cm")

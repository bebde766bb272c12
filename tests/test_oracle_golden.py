"""Pin the oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py; reference omniloc.py:160-202, :299-356, :11-102, :205-296,
utils.py:462-507, :208-229)."""
import numpy as np
import pytest
import torch

from oracle import piccolo_oracle as orc
from piccolo_b200 import synth


def rot_err_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return np.rad2deg(np.arccos(np.clip(c, -1, 1)))


def rot_of(pose):
    return orc.rot_and_derivs_np(pose[3:6], np.float64)[0]


@pytest.fixture(scope="module")
def small(golden):
    g = golden("loss_small")
    g["rgb"] = synth.rgb_from_u8(g["rgb8"])
    g["img"] = synth.img_from_u8(g["img8"])
    return g


def test_np64_loss_and_grad_match_reference_fp64(small):
    for i, p in enumerate(small["poses"]):
        loss, cnt, grad = orc.loss_and_grad_np(small["xyz"], small["rgb"], small["img"], p.astype(np.float64), np.float64)
        assert abs(loss - small["loss64"][i]) <= 1e-10 * abs(small["loss64"][i]), i
        np.testing.assert_allclose(grad, small["grad64"][i], rtol=0, atol=1e-8 * np.abs(small["grad64"][i]).max() + 1e-13)


def test_np32_loss_matches_reference_fp32(small):
    for i, p in enumerate(small["poses"]):
        loss, cnt, grad = orc.loss_and_grad_np(small["xyz"], small["rgb"], small["img"], p, np.float32)
        assert abs(loss - small["loss32"][i]) <= 2e-6 * abs(small["loss32"][i]), i
        g64 = small["grad64"][i]
        tol = max(1e-4 * np.abs(g64).max(), 4 * np.abs(small["grad32"][i] - g64).max())
        assert np.abs(grad - g64).max() <= tol, i


def test_torch_chain_matches_reference(small):
    xyz, rgb, img = [torch.from_numpy(small[k]) for k in ("xyz", "rgb", "img")]
    pose = torch.from_numpy(small["poses"]).clone().requires_grad_()
    loss, cnt = orc.sampling_loss_torch(xyz, rgb, img, pose)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().numpy(), small["loss32"], rtol=2e-6)
    g = pose.grad.numpy()
    for i in range(len(g)):
        g64 = small["grad64"][i]
        tol = max(1e-4 * np.abs(g64).max(), 4 * np.abs(small["grad32"][i] - g64).max())
        assert np.abs(g[i] - g64).max() <= tol, i
    # BatchSamplingLoss contract: (sum, list)  (omniloc.py:355-356)
    np.testing.assert_allclose(loss[:4].detach().numpy(), small["batch_list"], rtol=2e-6)
    assert abs(float(loss[:4].sum()) - float(small["batch_total"])) <= 4e-6 * float(small["batch_total"])


def test_empty_mask_is_nan(small):
    black = np.zeros_like(small["img"])
    loss, cnt, grad = orc.loss_and_grad_np(small["xyz"], small["rgb"], black, small["poses"][0], np.float32)
    assert cnt == 0 and np.isnan(loss) and np.isnan(small["black_loss"])
    l, c = orc.sampling_loss_torch(torch.from_numpy(small["xyz"]), torch.from_numpy(small["rgb"]), torch.from_numpy(black),
                                   torch.from_numpy(small["poses"][:1]))
    assert torch.isnan(l[0]) and int(c[0]) == 0


def test_quantile_box(small):
    lo, hi = orc.quantile_box_np(small["xyz"], 0.05)
    np.testing.assert_array_equal(lo, small["box_lo"])
    np.testing.assert_array_equal(hi, small["box_hi"])
    lo_t, hi_t = orc.quantile_box_torch(torch.from_numpy(small["xyz"]), 0.05)
    np.testing.assert_array_equal(lo_t.numpy(), small["box_lo"])
    np.testing.assert_array_equal(hi_t.numpy(), small["box_hi"])


def test_grid_scoring_and_topk(small, golden):
    g = golden("score_small")
    table, cnt = orc.score_poses_np(small["xyz"], small["rgb"], small["img"], g["grid"], np.float32)
    np.testing.assert_allclose(table, g["loss_table"], rtol=2e-6)
    idx = orc.topk_ascending(table, 10)
    R = len(g["rot"])
    np.testing.assert_array_equal(g["trans"][idx // R], g["top10_trans"])
    np.testing.assert_array_equal(g["rot"][idx % R], g["top10_rot"])
    tt, rr, tab = orc.score_grid_torch(torch.from_numpy(small["img"]), torch.from_numpy(small["xyz"]), torch.from_numpy(small["rgb"]),
                                       torch.from_numpy(g["trans"]), torch.from_numpy(g["rot"]), 10, pose_chunk=4)
    np.testing.assert_array_equal(tt.numpy(), g["top10_trans"])
    np.testing.assert_array_equal(rr.numpy(), g["top10_rot"])


def test_topk_ties_and_nan():
    loss = np.array([0.5, np.nan, 0.2, 0.2, 0.9, 0.1], dtype=np.float32)
    np.testing.assert_array_equal(orc.topk_ascending(loss, 4), [5, 2, 3, 0])
    np.testing.assert_array_equal(orc.topk_ascending(loss, 99), [5, 2, 3, 0, 4, 1])


@pytest.mark.parametrize("name", ["refine_small", "refine_medium"])
def test_refinement_trajectories(golden, name):
    from parity_util import EARLY_R, EARLY_T, final_pose_gates
    g = golden(name)
    rgb, img = synth.rgb_from_u8(g["rgb8"]), synth.img_from_u8(g["img8"])
    kw = dict(lr=0.1, patience=5, factor=float(g["factor"]), q=0.05)
    lo, hi = orc.quantile_box_np(g["xyz"], 0.05)
    chaotic = g["seq_loss"] > 3 * g["seq_loss"].min()   # stuck at the box corner: fp32 noise is amplified
    # early checkpoint: tight
    seq = orc.refine_np(g["xyz"], rgb, img, g["starts"], num_iter=int(g["early_iter"]), batch_semantics=False, dtype=np.float32, **kw)
    bat = orc.refine_np(g["xyz"], rgb, img, g["starts"], num_iter=int(g["early_iter"]), batch_semantics=True, dtype=np.float32, **kw)
    for b in range(len(g["starts"])):
        if chaotic[b]:
            continue
        assert np.linalg.norm(seq["pose"][b, :3] - g["early_seq_t"][b]) < EARLY_T, b
        assert rot_err_deg(rot_of(seq["pose"][b]), g["early_seq_R"][b]) < EARLY_R, b
        assert abs(seq["loss"][b] - g["early_seq_loss"][b]) <= 1e-2 * g["early_seq_loss"][b], b
    k = int(np.argmin(bat["loss"]))
    assert np.linalg.norm(bat["pose"][k, :3] - g["early_bat_t"]) < EARLY_T
    assert rot_err_deg(rot_of(bat["pose"][k]), g["early_bat_R"]) < EARLY_R
    # end state
    seq = orc.refine_np(g["xyz"], rgb, img, g["starts"], num_iter=int(g["num_iter"]), batch_semantics=False, dtype=np.float32, **kw)
    for b in range(len(g["starts"])):
        assert np.all(seq["pose"][b, :3] >= lo) and np.all(seq["pose"][b, :3] <= hi)
        if chaotic[b]:
            assert abs(seq["loss"][b] - g["seq_loss"][b]) <= 0.05 * g["seq_loss"][b], b
            continue
        gate_t, gate_r = final_pose_gates(g, b)
        assert np.linalg.norm(seq["pose"][b, :3] - g["seq_t"][b]) < gate_t, b
        assert rot_err_deg(rot_of(seq["pose"][b]), g["seq_R"][b]) < gate_r, b
        # the last-forward loss jitters with Adam's final steps; poses are the parity gate
        assert abs(seq["loss"][b] - g["seq_loss"][b]) <= 0.05 * g["seq_loss"][b], b
    bat = orc.refine_np(g["xyz"], rgb, img, g["starts"], num_iter=int(g["num_iter"]), batch_semantics=True, dtype=np.float32, **kw)
    k = int(np.argmin(bat["loss"]))
    gate_t, gate_r = final_pose_gates(g, None)
    assert np.linalg.norm(bat["pose"][k, :3] - g["bat_t"]) < gate_t
    assert rot_err_deg(rot_of(bat["pose"][k]), g["bat_R"]) < gate_r
    assert abs(bat["loss"][k] - g["bat_loss"]) <= 0.05 * g["bat_loss"]


def test_refine_torch_matches_np(golden):
    g = golden("refine_small")
    rgb, img = synth.rgb_from_u8(g["rgb8"]), synth.img_from_u8(g["img8"])
    kw = dict(lr=0.1, num_iter=int(g["num_iter"]), patience=5, factor=float(g["factor"]), q=0.05)
    xyz_t, rgb_t, img_t = torch.from_numpy(g["xyz"]), torch.from_numpy(rgb), torch.from_numpy(img)
    for bs in (False, True):
        a = orc.refine_torch(xyz_t, rgb_t, img_t, torch.from_numpy(g["starts"]), batch_semantics=bs, **kw)
        b = orc.refine_np(g["xyz"], rgb, img, g["starts"], batch_semantics=bs, dtype=np.float32, **kw)
        # candidate 2 (stuck at the box corner) is chaotic, see test_refinement_trajectories
        assert np.abs(a["pose"].numpy()[:2, :3] - b["pose"][:2, :3]).max() < 0.01
        np.testing.assert_allclose(a["loss"].numpy()[:2], b["loss"][:2], rtol=5e-2)
        np.testing.assert_allclose(a["loss"].numpy()[2], b["loss"][2], rtol=0.1)
    # the clamp quirk is visible in the fixture: candidate 2 starts outside the box
    assert np.isfinite(b["loss"]).all()


def test_histogram_rerank_matches_reference(golden):
    """make_pano + 8x8x8 block histograms + intersection (utils.py:510-588) incl. its ordering."""
    g = golden("rerank_small")
    rgb, img = synth.rgb_from_u8(g["rgb8"]), synth.img_from_u8(g["img8"])
    pose = g["poses"][20]
    R = orc.rot_and_derivs_np(pose[3:6], np.float32)[0]
    q = ((g["xyz"] - pose[None, :3]) @ R.T).astype(np.float32)
    pano = orc.make_pano_np(q, rgb, 128, 256)
    # Which of several points landing on one pixel wins is a RACE in the reference (index_put_ with duplicate
    # indices, accumulate=False, is parallel on CPU and non-deterministic on CUDA): the oracle
    # implements the intended painter's order (nearest point, centre write last).  Coverage is identical,
    # a few percent of pixels show a neighbouring point's colour.
    np.testing.assert_array_equal(pano.sum(2) > 0, g["pano20"].sum(2) > 0)
    assert (np.abs(pano - g["pano20"]).max(axis=2) > 0).mean() < 0.06
    scores = orc.hist_rerank_scores_np(img, g["xyz"], rgb, g["poses"], 4, 4)
    order = orc.hist_rerank_select(scores, len(g["poses"]))
    np.testing.assert_array_equal(g["poses"][order[:6], :3], g["top6_trans"])
    np.testing.assert_array_equal(g["poses"][order[:6], 3:], g["top6_rot"])
    # full ordering: identical except possibly swaps between near-tied neighbours
    ref_order = [int(np.where((g["poses"][:, :3] == t).all(1) & (g["poses"][:, 3:] == r).all(1))[0][0]) for t, r in zip(g["all_trans"], g["all_rot"])]
    assert sum(int(a != b) for a, b in zip(order, ref_order)) <= 2
    assert order[0] == 20                      # the near-GT pose wins the re-rank

"""Property tests (hypothesis) of the CUDA path against the numpy oracle on random small problems: random
cloud sizes (ragged tiles), panorama sizes down to 2x4 (zero padding at every border), arbitrary poses incl.
gimbal-lock pitch, black regions, both point orders and every texel format."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import piccolo_oracle as orc

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 3000), h=st.sampled_from([2, 3, 5, 8, 17, 64]), w=st.sampled_from([4, 7, 16, 33, 128]),
       fmt=st.sampled_from(["f16d", "u8q", "u8p", "tex", "f32"]), order=st.sampled_from([0, 1]), black=st.booleans())
def test_loss_and_gradient_match_oracle_on_random_problems(seed, n, h, w, fmt, order, black):
    from piccolo_b200 import engine
    rng = np.random.default_rng(seed)
    xyz = (rng.random((n, 3)) * np.array([8.0, 6.0, 3.0])).astype(np.float32)
    rgb8 = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    img8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if black:
        img8[: max(1, h // 3)] = 0
    rgb = (rgb8.astype(np.float64) / 255.0).astype(np.float32)
    img = img8.astype(np.float32) / np.float32(255.0)
    if fmt == "f32":                                   # arbitrary float data on the fp32 path
        img = np.clip(img + rng.normal(0, 0.01, img.shape).astype(np.float32), 0, 1).astype(np.float32) * (img8.sum(-1, keepdims=True) > 0)
        rgb = rng.random((n, 3)).astype(np.float32)
    poses = np.concatenate([rng.random((5, 3)) * np.array([8.0, 6.0, 3.0]), rng.uniform(-np.pi, 2 * np.pi, (5, 3))], axis=1).astype(np.float32)
    poses[1, 4] = np.float32(np.pi / 2)                # gimbal lock
    poses[2, 3:] = 0.0
    cloud = engine.Cloud(cu(xyz), cu(rgb), 0.05, order)
    image = engine.Image(cu(img), fmt)
    loss, cnt, grad = engine.loss_fwd_bwd(cloud, image, cu(poses))
    loss_s, cnt_s = engine.score(cloud, image, cu(poses))
    loss, cnt, grad = loss.cpu().numpy(), cnt.cpu().numpy(), grad.cpu().numpy()
    for i, p in enumerate(poses):
        l64, m64, g64 = orc.loss_and_grad_np(xyz, rgb, img, p.astype(np.float64), np.float64)
        l32, m32, g32 = orc.loss_and_grad_np(xyz, rgb, img, p, np.float32)
        if abs(m32 - m64) > 0 or abs(cnt[i] - m64) > max(2, 0.002 * n):
            continue        # a sample sits exactly on a black/non-black or pixel boundary: fp32 and fp64 already disagree
        if m64 == 0:
            assert np.isnan(loss[i]) and np.isnan(loss_s.cpu().numpy()[i])
            continue
        # small problems have no averaging: the gate is the reference's own fp32-vs-fp64 deviation (x4) or 1e-4
        assert abs(loss[i] - l64) <= max(1e-4 * abs(l64), 4 * abs(l32 - l64)) + 1e-7, (i, loss[i], l64, l32)
        if abs(cnt[i] - m64) == 0:
            gtol = max(1e-4 * np.abs(g64).max(), 4 * np.abs(g32 - g64).max()) + 1e-6
            assert np.abs(grad[i] - g64).max() <= gtol, (i, grad[i], g64, g32)
        assert abs(loss_s.cpu().numpy()[i] - loss[i]) <= 3e-6 * abs(loss[i]) + 1e-7


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), p=st.integers(1, 400), k=st.integers(1, 450))
def test_topk_matches_stable_argsort(seed, p, k):
    from piccolo_b200 import engine
    rng = np.random.default_rng(seed)
    loss = np.round(rng.random(p), 2).astype(np.float32)
    loss[rng.random(p) < 0.05] = np.nan
    np.testing.assert_array_equal(engine.topk(cu(loss), k).cpu().numpy(), orc.topk_ascending(loss, k))

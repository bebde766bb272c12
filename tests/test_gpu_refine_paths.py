"""The two synchronisation schemes of the fused refinement must produce the SAME bits: the persistent cooperative
kernel (split-phase grid barrier, all iterations in one launch) and the per-iteration launches (last-block-done
tickets) share ranges, phase arithmetic, record order and finalize (pcl_refine.cuh).  Any ordering or visibility bug
in the hand-written barrier shows up as a diverging trajectory.  Reference loops: omniloc.py:44-58, :249-269."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    from piccolo_b200 import engine, synth
    dev = torch.device("cuda:0")
    sc = synth.make_scene(150_001, 256, 512, seed=3)          # ragged size: every CTA range has a remainder
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    rng = np.random.default_rng(0)
    starts = np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.2, 3)]) for _ in range(16)]).astype(np.float32)
    return cloud, image, torch.from_numpy(starts).to(dev), sc


def _run(cloud, image, starts, B, batch, persist, plan):
    from piccolo_b200 import _lib, engine
    _lib.set_option("PERSIST", persist)
    try:
        ref = engine.Refiner(B, 0.1, 0.8, 5, batch).reset(starts[:B])
        for n in plan:
            ref.run(cloud, image, n)
        o = ref.read()
        return [o[k].cpu().numpy().copy() for k in ("pose", "param", "loss", "lr")]
    finally:
        _lib.set_option("PERSIST", -1)


@pytest.mark.parametrize("B", [1, 2, 3, 5, 6, 7, 9, 11, 16])
@pytest.mark.parametrize("batch", [False, True])
def test_persistent_equals_per_iteration_bitwise(scene, B, batch):
    cloud, image, starts, _ = scene
    a = _run(cloud, image, starts, B, batch, 1, (40,))
    b = _run(cloud, image, starts, B, batch, 0, (40,))
    c = _run(cloud, image, starts, B, batch, 1, (2, 37, 1))      # split runs carry the state exactly
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y, equal_nan=True), (B, batch, np.abs(x - y).max())
        assert np.array_equal(x, z, equal_nan=True), (B, batch, np.abs(x - z).max())
    assert np.isfinite(a[2]).all()


def test_fused_refinement_matches_generic_kernel_first_iterations(scene):
    """B <= 16 takes the fused path, B > 16 the generic fwd+bwd kernel: different decompositions of the same sums.
    The first iterations (before Adam amplifies rounding noise) must agree to fp32 accuracy."""
    from piccolo_b200 import engine
    cloud, image, starts, _ = scene
    s17 = torch.cat([starts, starts[:1]], dim=0)
    fused = engine.Refiner(16, 0.1, 0.8, 5, True).reset(starts).run(cloud, image, 3).read()
    gen = engine.Refiner(17, 0.1, 0.8, 5, True).reset(s17).run(cloud, image, 3).read()
    assert (fused["loss"] - gen["loss"][:16]).abs().max().item() <= 2e-5 * gen["loss"][:16].abs().max().item()
    assert (fused["pose"] - gen["pose"][:16]).abs().max().item() < 2e-4


def test_option_names_are_validated():
    from piccolo_b200 import _lib
    with pytest.raises(_lib.PiccoloError):
        _lib.set_option("NO_SUCH_KNOB", 1)


def test_partly_resident_cloud_equals_per_iteration_bitwise():
    """A cloud too large to keep whole in shared memory (1.3 M points: 3 of its 4.3 groups per CTA resident, the rest streamed
    evict-first) and one streamed entirely (option RF_RES=0) must give the same bits as the per-iteration launches, which read
    everything from global memory."""
    from piccolo_b200 import _lib, engine, synth
    dev = torch.device("cuda:0")
    sc = synth.make_scene(1_300_003, 256, 512, seed=4)
    xyz, rgb, img = [torch.from_numpy(a).to(dev) for a in (sc.xyz, sc.rgb, sc.img)]
    cloud, image = engine.Cloud(xyz, rgb), engine.Image(img)
    rng = np.random.default_rng(1)
    starts = torch.from_numpy(np.stack([sc.gt_pose + np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.2, 3)]) for _ in range(6)]).astype(np.float32)).to(dev)
    a = _run(cloud, image, starts, 6, True, 1, (12,))
    b = _run(cloud, image, starts, 6, True, 0, (12,))
    _lib.set_option("RF_RES", 0)
    try:
        c = _run(cloud, image, starts, 6, True, 1, (12,))
    finally:
        _lib.set_option("RF_RES", -1)
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y, equal_nan=True) and np.array_equal(x, z, equal_nan=True)
    assert np.isfinite(a[2]).all()

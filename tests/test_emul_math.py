"""Host emulation of the device arithmetic (piccolo_b200/csrc/pcl_eval.cuh compiled with g++) against
the golden vectors of the reference.  CPU-only; checks the maths the CUDA kernels execute."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from piccolo_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "pcl_emul.cpp")
SO = os.path.join(HERE, "emul", "_pcl_emul.so")
FMT = {"u8q": 1, "f32": 2, "u8p": 3, "f16d": 5}


@pytest.fixture(scope="module")
def emul():
    hdr = os.path.join(HERE, "..", "piccolo_b200", "csrc", "pcl_eval.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    lib = ctypes.CDLL(SO)
    fp = ctypes.POINTER(ctypes.c_float)

    def call(xyz, rgb, img, poses, fmt, bwd=True):
        xyz, rgb, img, poses = [np.ascontiguousarray(a, dtype=np.float32) for a in (xyz, rgb, img, poses)]
        P = len(poses)
        loss, cnt, grad = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros((P, 6), np.float32)
        lib.emul_loss_grad(xyz.ctypes.data_as(fp), rgb.ctypes.data_as(fp), ctypes.c_long(len(xyz)), img.ctypes.data_as(fp),
                           img.shape[0], img.shape[1], FMT[fmt], poses.ctypes.data_as(fp), P, int(bwd),
                           loss.ctypes.data_as(fp), cnt.ctypes.data_as(fp), grad.ctypes.data_as(fp))
        return loss, cnt, grad
    return call


@pytest.mark.parametrize("fmt", ["u8q", "f32", "u8p", "f16d"])
def test_emulated_kernel_math_matches_reference(emul, golden, fmt):
    g = golden("loss_small")
    rgb, img = synth.rgb_from_u8(g["rgb8"]), synth.img_from_u8(g["img8"])
    loss, cnt, grad = emul(g["xyz"], rgb, img, g["poses"], fmt)
    np.testing.assert_allclose(loss, g["loss32"], rtol=1e-5)          # gate is 1e-4
    for i in range(len(loss)):
        g64 = g["grad64"][i]
        tol = max(1e-4 * np.abs(g64).max(), np.abs(g["grad32"][i] - g64).max())
        assert np.abs(grad[i] - g64).max() <= tol, (i, grad[i], g64)


def test_emulated_black_image_is_nan(emul, golden):
    g = golden("loss_small")
    rgb, img = synth.rgb_from_u8(g["rgb8"]), synth.img_from_u8(g["img8"])
    loss, cnt, grad = emul(g["xyz"], rgb, np.zeros_like(img), g["poses"][:2], "u8q")
    assert np.isnan(loss).all() and (cnt == 0).all()


def test_emulated_structured_grid_matches_reference_losses(golden):
    """Shared elevation / shifted azimuth evaluation of rotations related by an in-plane rotation (pcl_grid_base +
    pcl_grid_member) against the reference's per-pose losses: yaw-only members of the golden scoring grid."""
    g = golden("score_small")
    small = golden("loss_small")
    rgb, img = synth.rgb_from_u8(small["rgb8"]), synth.img_from_u8(small["img8"])
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    lib = ctypes.CDLL(SO)
    fp = ctypes.POINTER(ctypes.c_float)
    xyz = np.ascontiguousarray(small["xyz"], dtype=np.float32)
    R = len(g["rot"])                                   # 8 yaws, pitch = roll = 0: one group
    delta = ((g["rot"][:, 0] - g["rot"][0, 0] + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    for ti in range(len(g["trans"])):
        base = np.concatenate([g["trans"][ti], g["rot"][0]]).astype(np.float32)
        loss, cnt = np.zeros(R, np.float32), np.zeros(R, np.float32)
        lib.emul_grid(xyz.ctypes.data_as(fp), rgb.ctypes.data_as(fp), ctypes.c_long(len(xyz)), img.ctypes.data_as(fp), img.shape[0], img.shape[1],
                      base.ctypes.data_as(fp), delta.ctypes.data_as(fp), R, loss.ctypes.data_as(fp), cnt.ctypes.data_as(fp))
        np.testing.assert_allclose(loss, g["loss_table"][ti * R:(ti + 1) * R], rtol=2e-5)


def _group_rotations(rot, tol=1e-6):
    """numpy mirror of pcl_grid_plan_kernel (pcl_grid.cu): greedy grouping on the third row of R, delta from R_j R_b^T"""
    from oracle import piccolo_oracle as orc
    Rs = [orc.rot_and_derivs_np(r.astype(np.float64), np.float64)[0] for r in rot]
    bases, groups = [], []
    for j, R in enumerate(Rs):
        for g, b in enumerate(bases):
            if np.abs(R[2] - Rs[b][2]).max() < tol:
                groups[g].append(j)
                break
        else:
            bases.append(j); groups.append([j])
    out = []
    for b, members in zip(bases, groups):
        deltas = []
        for j in members:
            M = Rs[j] @ Rs[b].T
            assert np.abs(M - np.array([[M[0, 0], -M[1, 0], 0], [M[1, 0], M[0, 0], 0], [0, 0, 1]])).max() < 5e-6     # a turn about z
            deltas.append(0.0 if j == b else np.arctan2(M[1, 0], M[0, 0]))
        out.append((b, members, np.asarray(deltas, np.float32)))
    return out


def test_emulated_structured_grid_on_the_euler_lattice(golden):
    """The 24 distinct rotations of the reference's 4x4x4 Euler lattice (utils.py:326-360) are 6 groups of 4 rotations
    related by an in-plane turn; evaluating each group through pcl_grid_base / pcl_grid_member reproduces the oracle's
    per-pose fp64 losses to the 1e-4 gate (measured ~1e-6) for every member."""
    from oracle import piccolo_oracle as orc
    from piccolo_b200 import utils as pu
    small = golden("loss_small")
    rgb, img = synth.rgb_from_u8(small["rgb8"]), synth.img_from_u8(small["img8"])
    xyz = np.ascontiguousarray(small["xyz"], dtype=np.float32)
    rot = pu.generate_rot_points({"yaw_only": False, "num_yaw": 4, "num_pitch": 4, "num_roll": 4}).numpy()
    groups = _group_rotations(rot)
    assert len(rot) == 24 and sorted(len(m) for _, m, _ in groups) == [4] * 6
    yaw8 = pu.generate_rot_points({"yaw_only": True, "num_yaw": 8}).numpy()
    assert [len(m) for _, m, _ in _group_rotations(yaw8)] == [8]
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    lib = ctypes.CDLL(SO)
    fp = ctypes.POINTER(ctypes.c_float)
    t = (xyz.min(0) + (xyz.max(0) - xyz.min(0)) * np.array([0.4, 0.55, 0.5])).astype(np.float32)
    worst = 0.0
    for b, members, delta in groups:
        base = np.concatenate([t, rot[b]]).astype(np.float32)
        loss, cnt = np.zeros(len(members), np.float32), np.zeros(len(members), np.float32)
        lib.emul_grid(xyz.ctypes.data_as(fp), rgb.ctypes.data_as(fp), ctypes.c_long(len(xyz)), img.ctypes.data_as(fp), img.shape[0], img.shape[1],
                      base.ctypes.data_as(fp), delta.ctypes.data_as(fp), len(members), loss.ctypes.data_as(fp), cnt.ctypes.data_as(fp))
        for k, j in enumerate(members):
            want = orc.loss_and_grad_np(xyz, rgb, img, np.concatenate([t, rot[j]]).astype(np.float64), np.float64, want_grad=False)[0]
            worst = max(worst, abs(loss[k] - want) / want)
    assert worst < 1e-4, worst


def test_emulated_structured_grid_reproduces_reference_lattice_table(golden):
    """tests/golden/score_lattice.npz: trim_input_loss of the UNMODIFIED reference over 5 translations x the 24 lattice
    rotations.  The structured evaluation (6 groups of 4, numpy mirror of the plan kernel + host emulation of
    pcl_grid_base / pcl_grid_member) reproduces the reference's loss table to the 1e-4 gate and its top-10 order."""
    from oracle import piccolo_oracle as orc
    g, small = golden("score_lattice"), golden("loss_small")
    rgb, img = synth.rgb_from_u8(small["rgb8"]), synth.img_from_u8(small["img8"])
    xyz = np.ascontiguousarray(small["xyz"], dtype=np.float32)
    trans, rot = g["trans"], g["rot"]
    groups = _group_rotations(rot)
    assert sorted(len(m) for _, m, _ in groups) == [4] * 6
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    lib = ctypes.CDLL(SO)
    fp = ctypes.POINTER(ctypes.c_float)
    table = np.zeros(len(trans) * len(rot), np.float32)
    for i, t in enumerate(trans):
        for b, members, delta in groups:
            base = np.concatenate([t, rot[b]]).astype(np.float32)
            loss, cnt = np.zeros(len(members), np.float32), np.zeros(len(members), np.float32)
            lib.emul_grid(xyz.ctypes.data_as(fp), rgb.ctypes.data_as(fp), ctypes.c_long(len(xyz)), img.ctypes.data_as(fp), img.shape[0], img.shape[1],
                          base.ctypes.data_as(fp), delta.ctypes.data_as(fp), len(members), loss.ctypes.data_as(fp), cnt.ctypes.data_as(fp))
            table[i * len(rot) + np.asarray(members)] = loss
    np.testing.assert_allclose(table, g["loss_table"], rtol=1e-4)
    idx = orc.topk_ascending(table, 10)
    np.testing.assert_array_equal(trans[idx // len(rot)], g["top10_trans"])
    np.testing.assert_array_equal(rot[idx % len(rot)], g["top10_rot"])

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

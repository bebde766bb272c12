"""The C-ABI library loads on a machine without a GPU and exports every symbol include/piccolo_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "piccolo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from piccolo_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)
    assert lib.pcl_abi_version() == 2
    assert lib.pcl_launch_count() >= 0


def test_calls_fail_loudly_without_valid_arguments():
    from piccolo_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.pcl_cloud_create(None, None, 0, 0.05, 1, None, ctypes.byref(h)) == -1
    assert b"bad cloud" in lib.pcl_last_error()
    assert lib.pcl_refine_create(0, 0.1, 0.8, 5, 0, ctypes.byref(h)) == -1
    assert lib.pcl_score(None, None, None, 1, None, None, None) == -1


def test_no_cpu_fallback_in_python_api():
    import numpy as np
    import pytest
    import torch
    from piccolo_b200 import _lib, engine
    with pytest.raises(_lib.PiccoloError):
        engine.Cloud(torch.zeros(4, 3), torch.zeros(4, 3))
    with pytest.raises(_lib.PiccoloError):
        engine.Image(torch.zeros(8, 16, 3))
    # the product never imports the oracle
    import piccolo_b200
    pkg = os.path.dirname(piccolo_b200.__file__)
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("# oracle", ""), f


def test_options_and_communicator_arguments_are_validated():
    from piccolo_b200 import _lib
    lib = _lib.load()
    assert lib.pcl_set_option(b"PERSIST", 1) == 0 and lib.pcl_set_option(b"persist", -1) == 0
    assert lib.pcl_set_option(b"NO_SUCH_KNOB", 1) == -1 and b"unknown option" in lib.pcl_last_error()
    h = ctypes.c_void_p()
    assert lib.pcl_comm_create(3, 2, 0, ctypes.byref(h)) == -1          # rank outside the communicator
    assert lib.pcl_comm_create(0, 9, 0, ctypes.byref(h)) == -1          # more ranks than one box has GPUs
    assert lib.pcl_comm_barrier(None, None) == -1
    assert lib.pcl_refine_run_sharded(None, None, None, 1, None, None) == -1

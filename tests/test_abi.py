"""The C-ABI library loads, exports every symbol include/rsb_c_api.h declares, and its boundary
PODs have the byte layout of the reference structs (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rocketsim_b200 import lib, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsb_c_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_lib_binds():
    assert set(lib.EXPORTS) == set(declared_symbols())


def test_cuda_library_exports_every_symbol():
    if os.path.isdir("/usr/local/cuda/bin"):
        from rocketsim_b200 import build
        build.build_cuda()  # incremental: rebuilds only when a source is newer than the library
    L = C.CDLL(lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), name


def test_pod_sizes_match_reference_layouts():
    L = lib.load()
    for name, size in T.SIZES.items():
        assert L.rsb_sizeof(name.encode()) == size, name
    # reference sizeof (SURVEY.md A.4 probe): CarState 352, BallState 144, CarControls 24, MutatorConfig 128
    assert T.CarState.itemsize == 352 and T.BallState.itemsize == 144
    assert T.CarControls.itemsize == 24 and T.MutatorConfig.itemsize == 128


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(lib.RsbError):
        lib.load(str(tmp_path / "nope.so"))


def test_no_device_fails_loudly(mesh_sets):
    """Without a CUDA device the product library refuses to create a batch (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    L = lib.load()
    lib.init_from_mem(mesh_sets, L)  # idempotent per process, like RocketSim::InitFromMem
    h = C.c_void_p()
    rc = L.rsb_batch_create(0, T.MODE_THE_VOID, 1, 2, None, None, None, C.c_float(120.0), 0, C.byref(h))
    assert rc != 0
    assert b"CUDA" in L.rsb_last_error() or b"device" in L.rsb_last_error()


def test_batch_create_argument_errors(hostlib):
    from rocketsim_b200.arena_batch import ArenaBatch
    with pytest.raises(lib.RsbError):  # Arena.cpp:419 tick-rate precondition
        ArenaBatch(T.MODE_THE_VOID, 1, [0, 1], tick_rate=10.0, lib=hostlib)
    with pytest.raises(lib.RsbError):
        ArenaBatch(T.MODE_THE_VOID, 0, [0, 1], lib=hostlib)
    with pytest.raises(lib.RsbError):
        ArenaBatch(T.MODE_THE_VOID, 1, [0] * 7, lib=hostlib)
    b = ArenaBatch(T.MODE_THE_VOID, 2, [0, 1], lib=hostlib)
    with pytest.raises(lib.RsbError):
        b.set_ball_states([5], np.zeros(1, T.BallState))

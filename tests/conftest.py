"""Test tiers:
  -m "not gpu": oracle vs golden vectors, host logic, the C-ABI library loads and exports every
                declared symbol, and the device code compiled for the host (tests/_hostbuild,
                TEST ONLY) against the oracle / golden records.
  -m gpu      : the parity tests proper -- the CUDA library through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


collect_ignore = ["tools"]  # developer scripts that drive the oracle (not tests)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mesh_sets():
    from rocketsim_b200 import synthetic_meshes as sm
    return sm.mesh_sets()


@pytest.fixture(scope="session")
def oracle(mesh_sets):
    """The unmodified reference (oracle/_ref).  Built from /root/reference where present."""
    from oracle import refsim
    if not refsim.available():
        if not os.path.isdir(refsim.REFERENCE_ROOT):
            pytest.skip("oracle/_ref not built and /root/reference absent")
        refsim.build()
    refsim.init_from_mem(mesh_sets)
    return refsim


@pytest.fixture(scope="session")
def hostlib(mesh_sets):
    """Host-compiled check build of the device code (tests only)."""
    import parity_util as pu
    from rocketsim_b200 import lib
    L = lib.load(pu.build_hostcheck())
    lib.init_from_mem(mesh_sets, L)
    return L


@pytest.fixture(scope="session")
def cudalib(mesh_sets):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    from rocketsim_b200 import lib
    L = lib.load()  # rocketsim_b200/librsb.so; raises if it was not built: no fallback
    lib.init_from_mem(mesh_sets, L)
    return L

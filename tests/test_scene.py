"""Static-scene builder: the quantised BVH, subtree headers and triangle
info map built by rsb_init_from_mem must equal the reference's bit for bit, for every mesh."""
import ctypes as C

import numpy as np
import pytest

from rocketsim_b200 import types as T


def my_dump(L, mode, mesh):
    cnt = (C.c_int * 5)()
    assert L.rsb_scene_counts(mode, mesh, cnt) == 0
    n_nodes, n_hdr, n_tris, n_verts, n_info = list(cnt)
    nodes = np.zeros(n_nodes * 16, np.uint8)
    hdrs = np.zeros(max(n_hdr, 1) * 20, np.uint8)
    quant = np.zeros(9, np.float32)
    info = np.zeros(n_tris * 16, np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.rsb_scene_dump(mode, mesh, vp(nodes), vp(hdrs), quant.ctypes.data_as(C.POINTER(C.c_float)), vp(info)) == 0
    return dict(counts=list(cnt), nodes=nodes, headers=hdrs[: n_hdr * 20], quant=quant, info=info)


@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_HEATSEEKER, T.MODE_SNOWDAY])
def test_bvh_and_edge_info_match_reference(oracle, hostlib, mode):
    n = oracle.lib().ref_num_meshes(mode)
    assert n == hostlib.rsb_scene_num_meshes(mode) and n > 0
    for mesh in range(n):
        ref = oracle.bvh_dump(mode, mesh)
        mine = my_dump(hostlib, mode, mesh)
        # the reference allocates 2N node slots and uses m_curNodeIndex = 2N-1 of them (btQuantizedBvh.cpp:147-192)
        n_used = 2 * len(ref["triinfo"]) - 1
        assert mine["counts"][0] == n_used and len(ref["nodes"]) in (n_used, n_used + 1)
        ref["nodes"] = ref["nodes"][:n_used]
        assert mine["counts"][1] == len(ref["headers"])
        assert mine["counts"][3] == ref["n_verts"] and mine["counts"][4] == ref["n_info"]
        assert mine["nodes"].tobytes() == ref["nodes"].tobytes()
        assert mine["headers"].tobytes() == ref["headers"].tobytes()
        np.testing.assert_array_equal(mine["quant"], ref["quant"])
        info_dt = np.dtype({"names": ["flags", "angles"], "formats": ["<i4", ("<f4", 3)], "itemsize": 16})
        mi = mine["info"].view(info_dt)
        np.testing.assert_array_equal(mi["flags"], ref["triinfo"]["flags"])
        # edge angles come out of atan2f on slightly different operand roundings: a few ulp
        np.testing.assert_allclose(mi["angles"], ref["triinfo"]["angles"], rtol=0, atol=2e-6)


def test_cuda_library_builds_the_same_scene(hostlib, mesh_sets):
    """rsb_init_from_mem is host code: the CUDA library must build the identical scene (no GPU needed)."""
    from rocketsim_b200 import lib
    L = lib.load()
    lib.init_from_mem(mesh_sets, L)
    for mode in (T.MODE_SOCCAR, T.MODE_HOOPS):
        for mesh in range(hostlib.rsb_scene_num_meshes(mode)):
            a, b = my_dump(hostlib, mode, mesh), my_dump(L, mode, mesh)
            assert a["nodes"].tobytes() == b["nodes"].tobytes() and a["info"].tobytes() == b["info"].tobytes()

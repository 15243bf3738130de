"""ctypes binding of the C ABI in ``include/rsb_c_api.h`` (the CUDA shared library).

The product library is ``rocketsim_b200/librsb.so``, built in-tree for sm_100a by
``rocketsim_b200.build`` (``__graft_entry__.build()``).  There is no CPU fallback: if the
library is missing this module raises, and every batch call fails without a CUDA device.

``load(path)`` with an explicit path exists so that the CPU-only test tier can load the
host-compiled check build (``tests/_hostbuild``) of the same device code; the package itself
never does that.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librsb.so")

EXPORTS = [
    "rsb_last_error", "rsb_sizeof", "rsb_init_from_mem", "rsb_batch_create", "rsb_batch_destroy", "rsb_batch_info",
    "rsb_set_stream", "rsb_set_controls", "rsb_set_controls_device", "rsb_step", "rsb_step_with_controls_device",
    "rsb_sync", "rsb_get_car_states", "rsb_set_car_states", "rsb_get_ball_states", "rsb_set_ball_states",
    "rsb_get_pad_states", "rsb_set_pad_states", "rsb_get_tick_counts", "rsb_is_ball_scored", "rsb_get_events",
    "rsb_get_error_flags", "rsb_seed_rng", "rsb_get_observations", "rsb_observation_floats",
    "rsb_get_car_internals", "rsb_set_car_internals", "rsb_get_ball_internals", "rsb_set_ball_internals",
    "rsb_reset_to_random_kickoff", "rsb_scene_counts", "rsb_scene_dump", "rsb_scene_num_meshes", "rsb_scene_cells",
    "rsb_bench_fill_controls_device", "rsb_debug_get_wheels", "rsb_step_with_controls", "rsb_set_auto_reset", "rsb_debug_get_manifolds", "rsb_debug_get_solver_rows", "rsb_debug_profile", "rsb_debug_get_solver_bodies", "rsb_debug_get_misc", "rsb_car_config_preset", "rsb_default_mutator_config", "rsb_set_event_tracker", "rsb_is_ball_probably_going_in", "rsb_predict_ball", "rsb_serialize_arena", "rsb_deserialize_arena", "rsb_clone_arena", "rsb_set_arena_offset", "rsb_batch_create_ex", "rsb_default_arena_config",
    "rsb_get_dropshot_tiles", "rsb_set_dropshot_tiles", "rsb_debug_set_detect_queue_cap",
]


class RsbError(RuntimeError):
    pass


def _declare(L):
    vp, i32, u32, u64, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
    L.rsb_last_error.restype = C.c_char_p
    L.rsb_sizeof.argtypes = [C.c_char_p]
    L.rsb_init_from_mem.argtypes = [i32, vp, vp, vp]
    L.rsb_batch_create.argtypes = [i32, i32, i32, i32, vp, vp, vp, f32, u32, vp]
    L.rsb_batch_create_ex.argtypes = [i32, i32, vp, i32, i32, vp, vp, vp, f32, u32, vp]
    L.rsb_default_arena_config.argtypes = [vp]
    L.rsb_default_arena_config.restype = None
    L.rsb_batch_destroy.argtypes = [vp]
    L.rsb_batch_info.argtypes = [vp, vp, vp, vp, vp]
    L.rsb_set_stream.argtypes = [vp, vp]
    L.rsb_set_controls.argtypes = [vp, vp]
    L.rsb_set_controls_device.argtypes = [vp, vp]
    L.rsb_step.argtypes = [vp, i32]
    L.rsb_step_with_controls_device.argtypes = [vp, vp, i32]
    L.rsb_sync.argtypes = [vp]
    L.rsb_step_with_controls.argtypes = [vp, vp, i32]
    L.rsb_set_auto_reset.argtypes = [vp, i32, u32]
    L.rsb_set_arena_offset.argtypes = [vp, u32]
    for n in ("car_states", "ball_states", "pad_states", "tick_counts", "error_flags", "observations", "car_internals",
              "ball_internals"):
        getattr(L, "rsb_get_" + n).argtypes = [vp, vp]
    L.rsb_is_ball_scored.argtypes = [vp, vp]
    L.rsb_set_car_states.argtypes = [vp, vp, vp, i32, vp]
    L.rsb_set_car_internals.argtypes = [vp, vp, vp, i32, vp]
    L.rsb_set_ball_states.argtypes = [vp, vp, i32, vp]
    L.rsb_set_ball_internals.argtypes = [vp, vp, i32, vp]
    L.rsb_set_pad_states.argtypes = [vp, vp, i32, vp]
    if hasattr(L, "rsb_get_dropshot_tiles"):  # (absent from older developer A/B builds; test_abi checks the product library exports it)
        L.rsb_get_dropshot_tiles.argtypes = [vp, vp]
        L.rsb_set_dropshot_tiles.argtypes = [vp, vp, i32, vp]
    L.rsb_get_events.argtypes = [vp, vp, i32, vp]
    L.rsb_seed_rng.argtypes = [vp, u32, u32]
    L.rsb_observation_floats.argtypes = [vp]
    L.rsb_reset_to_random_kickoff.argtypes = [vp, vp, i32, vp]
    L.rsb_scene_counts.argtypes = [i32, i32, vp]
    L.rsb_scene_dump.argtypes = [i32, i32, vp, vp, vp, vp]
    L.rsb_scene_num_meshes.argtypes = [i32]
    L.rsb_scene_cells.argtypes = [i32, vp, i32, vp, i32, vp, vp]
    L.rsb_bench_fill_controls_device.argtypes = [vp, vp, i32, u32, u64]
    L.rsb_debug_get_wheels.argtypes = [vp, i32, vp]
    L.rsb_debug_get_solver_rows.argtypes = [vp, i32, i32, vp, i32, vp]
    L.rsb_debug_get_solver_bodies.argtypes = [vp, i32, vp, i32, vp]
    L.rsb_debug_profile.argtypes = [vp, i32]
    L.rsb_debug_get_misc.argtypes = [vp, i32, vp]
    if hasattr(L, "rsb_debug_set_detect_queue_cap"):
        L.rsb_debug_set_detect_queue_cap.argtypes = [vp, i32]
    L.rsb_predict_ball.argtypes = [vp, i32, vp]
    L.rsb_clone_arena.argtypes = [vp, u32, u32]
    L.rsb_serialize_arena.argtypes = [vp, u32, vp, C.c_size_t, vp]
    L.rsb_deserialize_arena.argtypes = [vp, u32, vp, C.c_size_t]
    L.rsb_set_event_tracker.argtypes = [vp, i32, vp]
    L.rsb_is_ball_probably_going_in.argtypes = [vp, f32, f32, vp, vp]
    L.rsb_car_config_preset.argtypes = [i32, vp]
    L.rsb_default_mutator_config.argtypes = [i32, vp]
    L.rsb_debug_get_manifolds.argtypes = [vp, i32, vp, i32, vp]
    return L


_libs = {}


def load(path: str | None = None):
    """Load the C-ABI library.  Default = the CUDA build; raises if it has not been built."""
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise RsbError(
                f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "rocketsim_b200 has no CPU fallback.")
        _libs[path] = _declare(C.CDLL(path))
    return _libs[path]


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def check(L, rc):
    if rc != 0:
        raise RsbError(L.rsb_last_error().decode(errors="replace"))


def init_from_mem(mesh_sets: dict, lib=None):
    """``RocketSim::InitFromMem`` replacement: ``{game_mode: [cmf bytes, ...]}``."""
    L = lib or load()
    modes, bufs = [], []
    for mode in sorted(mesh_sets):
        for data in mesh_sets[mode]:
            modes.append(mode)
            bufs.append(C.create_string_buffer(data, len(data)))
    n = len(modes)
    m_arr = (C.c_int * max(n, 1))(*modes)
    p_arr = (C.c_void_p * max(n, 1))(*[C.cast(b, C.c_void_p).value for b in bufs])
    s_arr = (C.c_size_t * max(n, 1))(*[len(b) for b in bufs])
    check(L, L.rsb_init_from_mem(n, m_arr, p_arr, s_arr))

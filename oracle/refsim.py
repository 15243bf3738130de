"""ctypes view of the ORACLE (oracle/_ref/librocketsim_ref.so) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  It drives the unmodified reference
RocketSim (compiled by ``oracle/Makefile`` from ``/root/reference``) through the C shim in
``oracle/ref_harness.cpp``.  The product package ``rocketsim_b200`` never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rocketsim_b200 import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "librocketsim_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(quiet: bool = True) -> bool:
    """Compile the reference into oracle/_ref (only possible where /root/reference exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return os.path.exists(LIB_PATH)
    r = subprocess.run(["make", "-C", _HERE, "-j8"], capture_output=quiet, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return True


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(LIB_PATH)
        vp, u32, i32, f32, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_uint64
        L.ref_init_from_mem.argtypes = [i32, vp, vp, vp]; L.ref_init_from_mem.restype = i32
        L.ref_last_error.restype = C.c_char_p
        L.ref_arena_create.argtypes = [i32, f32]; L.ref_arena_create.restype = vp
        L.ref_arena_create_cfg.argtypes = [i32, f32, i32, i32, vp]; L.ref_arena_create_cfg.restype = vp
        L.ref_arena_destroy.argtypes = [vp]
        L.ref_add_car.argtypes = [vp, i32, i32]; L.ref_add_car.restype = u32
        L.ref_run_bench_inputs.argtypes = [vp, i32, C.c_uint64, i32, i32]; L.ref_run_bench_inputs.restype = C.c_double
        L.ref_num_cars.argtypes = [vp]; L.ref_num_pads.argtypes = [vp]
        L.ref_seed_thread_rng.argtypes = [u32]
        L.ref_reset_kickoff.argtypes = [vp, i32]
        L.ref_set_controls.argtypes = [vp, u32, vp]
        L.ref_step.argtypes = [vp, i32]
        L.ref_tick_count.argtypes = [vp]; L.ref_tick_count.restype = u64
        L.ref_is_ball_scored.argtypes = [vp]; L.ref_is_ball_scored.restype = i32
        L.ref_enable_event_tracker.argtypes = [vp, vp]
        L.ref_ball_pred.argtypes = [vp, i32, vp]; L.ref_ball_pred.restype = i32
        L.ref_clone.argtypes = [vp]; L.ref_clone.restype = vp
        L.ref_serialize.argtypes = [vp, vp, C.c_size_t]; L.ref_serialize.restype = C.c_size_t
        L.ref_deserialize_new.argtypes = [vp, C.c_size_t]; L.ref_deserialize_new.restype = vp
        L.ref_is_ball_probably_going_in.argtypes = [vp, f32, f32, vp]; L.ref_is_ball_probably_going_in.restype = i32
        for n in ("car_state", "car_internal"):
            getattr(L, "ref_get_" + n).argtypes = [vp, u32, vp]
            getattr(L, "ref_set_" + n).argtypes = [vp, u32, vp]
        for n in ("ball_state", "ball_internal", "pad_states", "mutator_config"):
            getattr(L, "ref_get_" + n).argtypes = [vp, vp]
            getattr(L, "ref_set_" + n).argtypes = [vp, vp]
        L.ref_get_pad_positions.argtypes = [vp, vp]
        L.ref_get_dropshot_tiles.argtypes = [vp, vp]; L.ref_set_dropshot_tiles.argtypes = [vp, vp]
        L.ref_get_dropshot_tile_pos.argtypes = [i32, i32, vp]
        L.ref_default_mutator_config.argtypes = [i32, vp]
        L.ref_get_car_config.argtypes = [i32, vp]
        L.ref_drain_events.argtypes = [vp, vp, i32]; L.ref_drain_events.restype = i32
        L.ref_num_meshes.argtypes = [i32]; L.ref_num_meshes.restype = i32
        L.ref_bvh_counts.argtypes = [i32, i32, vp]; L.ref_bvh_counts.restype = i32
        L.ref_bvh_dump.argtypes = [i32, i32, vp, vp, vp, vp]; L.ref_bvh_dump.restype = i32
        L.ref_get_wheel_debug.argtypes = [vp, u32, vp]
        L.ref_get_car_body.argtypes = [vp, u32, vp]
        L.ref_get_manifolds.argtypes = [vp, vp, i32]; L.ref_get_manifolds.restype = i32
        L.ref_get_solver_rows.argtypes = [vp, i32, i32, vp]
        L.ref_get_solver_bodies.argtypes = [vp, i32, vp]
        L.ref_bench.argtypes = [i32, i32, i32, i32, i32, u64, vp]; L.ref_bench.restype = C.c_double
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def init_from_mem(mesh_sets: dict) -> None:
    """mesh_sets: {game_mode: [cmf bytes, ...]}.  Idempotent per process (like the reference)."""
    modes, bufs = [], []
    for mode in sorted(mesh_sets):
        for data in mesh_sets[mode]:
            modes.append(mode)
            bufs.append(C.create_string_buffer(data, len(data)))
    n = len(modes)
    m_arr = (C.c_int * max(n, 1))(*modes)
    p_arr = (C.c_void_p * max(n, 1))(*[C.cast(b, C.c_void_p).value for b in bufs])
    s_arr = (C.c_size_t * max(n, 1))(*[len(b) for b in bufs])
    if lib().ref_init_from_mem(n, m_arr, p_arr, s_arr) != 0:
        raise RuntimeError(lib().ref_last_error().decode())


class RefArena:
    def __init__(self, mode: int, tick_rate: float = 120.0, no_ball_rot: bool = True, custom_pads=None):
        """custom_pads: None (default pads) or an array n x (x, y, z, isBig) -> ArenaConfig::customBoostPads"""
        self.L = lib()
        if no_ball_rot and custom_pads is None:
            self.h = self.L.ref_arena_create(mode, tick_rate)
        else:
            pads = None if custom_pads is None else np.ascontiguousarray(custom_pads, dtype=np.float32).reshape(-1, 4)
            self.h = self.L.ref_arena_create_cfg(mode, tick_rate, int(no_ball_rot), -1 if pads is None else len(pads), None if pads is None else _ptr(pads))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self.mode = mode
        self.car_ids = []

    def clone(self):
        c = RefArena.__new__(RefArena)
        c.L = self.L; c.h = self.L.ref_clone(self.h); c.mode = self.mode; c.car_ids = list(self.car_ids)
        return c

    def serialize(self) -> bytes:
        n = self.L.ref_serialize(self.h, None, 0)
        buf = (C.c_uint8 * n)()
        self.L.ref_serialize(self.h, buf, n)
        return bytes(buf)

    @classmethod
    def deserialize_new(cls, data: bytes, mode: int, n_cars: int):
        self = cls.__new__(cls)
        self.L = lib()
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        self.h = self.L.ref_deserialize_new(buf, len(data))
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self.mode = mode
        self.car_ids = list(range(1, n_cars + 1))
        return self

    def close(self):
        if self.h:
            self.L.ref_arena_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_car(self, team: int, preset: int = 0) -> int:
        i = self.L.ref_add_car(self.h, team, preset)
        self.car_ids.append(i)
        return i

    @property
    def num_pads(self):
        return self.L.ref_num_pads(self.h)

    def seed_rng(self, seed: int):
        self.L.ref_seed_thread_rng(seed)

    def reset_kickoff(self, seed: int):
        self.L.ref_reset_kickoff(self.h, seed)

    def set_controls(self, car_id: int, ctrl: np.ndarray):
        c = np.ascontiguousarray(ctrl, dtype=T.CarControls)
        self.L.ref_set_controls(self.h, car_id, _ptr(c))

    def step(self, n: int = 1):
        self.L.ref_step(self.h, n)

    def run_bench_inputs(self, ticks: int, seed: int, gid: int, tick0: int = 0) -> float:
        """`ticks` x Arena::Step with the pinned README-bench input stream of global arena `gid`, in C; returns the wall seconds"""
        return float(self.L.ref_run_bench_inputs(self.h, ticks, seed, gid, tick0))

    @property
    def tick_count(self):
        return self.L.ref_tick_count(self.h)

    def ball_pred(self, n_ticks):
        out = np.zeros(n_ticks, dtype=T.BallState)
        n = lib().ref_ball_pred(self.h, n_ticks, _ptr(out))
        return out[:n]

    def enable_event_tracker(self, cfg7=None):
        arr = None if cfg7 is None else np.asarray(cfg7, np.float32)
        lib().ref_enable_event_tracker(self.h, None if arr is None else _ptr(arr))

    def is_ball_probably_going_in(self, max_time=2.0, extra_margin=0.0):
        t = C.c_int()
        r = lib().ref_is_ball_probably_going_in(self.h, max_time, extra_margin, C.byref(t))
        return bool(r), t.value

    def is_ball_scored(self):
        return bool(self.L.ref_is_ball_scored(self.h))

    def _get(self, name, dtype, car_id=None, n=1):
        out = np.zeros(n, dtype=dtype)
        f = getattr(self.L, "ref_get_" + name)
        if car_id is None:
            f(self.h, _ptr(out))
        else:
            f(self.h, car_id, _ptr(out))
        return out

    def car_state(self, car_id):
        return self._get("car_state", T.CarState, car_id)[0]

    def car_internal(self, car_id):
        return self._get("car_internal", T.CarInternal, car_id)[0]

    def ball_state(self):
        return self._get("ball_state", T.BallState)[0]

    def ball_internal(self):
        return self._get("ball_internal", T.BallInternal)[0]

    def pad_states(self):
        return self._get("pad_states", T.BoostPadState, n=max(self.num_pads, 1))[: self.num_pads]

    def mutator_config(self):
        return self._get("mutator_config", T.MutatorConfig)[0]

    def set_car_state(self, car_id, s):
        a = np.array([s], dtype=T.CarState)
        self.L.ref_set_car_state(self.h, car_id, _ptr(a))

    def set_car_internal(self, car_id, s):
        a = np.array([s], dtype=T.CarInternal)
        self.L.ref_set_car_internal(self.h, car_id, _ptr(a))

    def set_ball_state(self, s):
        a = np.array([s], dtype=T.BallState)
        self.L.ref_set_ball_state(self.h, _ptr(a))

    def set_ball_internal(self, s):
        a = np.array([s], dtype=T.BallInternal)
        self.L.ref_set_ball_internal(self.h, _ptr(a))

    def set_pad_states(self, s):
        a = np.ascontiguousarray(s, dtype=T.BoostPadState)
        if len(a):
            self.L.ref_set_pad_states(self.h, _ptr(a))

    def set_mutator_config(self, m):
        a = np.array([m], dtype=T.MutatorConfig)
        self.L.ref_set_mutator_config(self.h, _ptr(a))

    def dropshot_tiles(self):
        out = np.zeros(140, dtype=np.uint8)
        self.L.ref_get_dropshot_tiles(self.h, _ptr(out))
        return out

    def set_dropshot_tiles(self, states):
        a = np.ascontiguousarray(states, dtype=np.uint8).reshape(140)
        self.L.ref_set_dropshot_tiles(self.h, _ptr(a))

    def pad_positions(self):
        out = np.zeros((max(self.num_pads, 1), 4), dtype=np.float32)
        self.L.ref_get_pad_positions(self.h, _ptr(out))
        return out[: self.num_pads]

    def drain_events(self, cap=256):
        out = np.zeros(cap, dtype=T.Event)
        n = self.L.ref_drain_events(self.h, _ptr(out), cap)
        return out[:n]

    def wheel_debug(self, car_id):
        out = np.zeros((4, 14), dtype=np.float32)
        self.L.ref_get_wheel_debug(self.h, car_id, _ptr(out))
        return out

    def manifolds(self, cap=64):
        out = np.zeros((cap, 16), dtype=np.float32)
        n = self.L.ref_get_manifolds(self.h, _ptr(out), cap)
        return out[:n]

    def solver_rows(self, which, n):
        out = np.zeros((n, 24), dtype=np.float32)
        self.L.ref_get_solver_rows(self.h, which, n, _ptr(out))
        return out

    def solver_bodies(self, n):
        out = np.zeros((n, 12), dtype=np.float32)
        self.L.ref_get_solver_bodies(self.h, n, _ptr(out))
        return out

    def car_body(self, car_id):
        out = np.zeros(18, dtype=np.float32)
        self.L.ref_get_car_body(self.h, car_id, _ptr(out))
        return out


def dropshot_tile_pos(team: int, idx: int):
    out = np.zeros(3, dtype=np.float32)
    lib().ref_get_dropshot_tile_pos(team, idx, _ptr(out))
    return out


def default_mutator_config(mode: int):
    out = np.zeros(1, dtype=T.MutatorConfig)
    lib().ref_default_mutator_config(mode, _ptr(out))
    return out[0]


def car_config(preset: int):
    out = np.zeros(1, dtype=T.CarConfig)
    lib().ref_get_car_config(preset, _ptr(out))
    return out[0]


def bvh_dump(mode: int, mesh: int):
    L = lib()
    cnt = (C.c_int * 5)()
    if L.ref_bvh_counts(mode, mesh, cnt) != 0:
        raise IndexError("no such mesh")
    n_nodes, n_hdr, n_tris, n_verts, n_info = list(cnt)
    node_dt = np.dtype({"names": ["qmin", "qmax", "idx"], "formats": [("<u2", 3), ("<u2", 3), "<i4"], "itemsize": 16})
    hdr_dt = np.dtype({"names": ["qmin", "qmax", "root", "size"], "formats": [("<u2", 3), ("<u2", 3), "<i4", "<i4"],
                       "offsets": [0, 6, 12, 16], "itemsize": 20})
    info_dt = np.dtype({"names": ["flags", "angles"], "formats": ["<i4", ("<f4", 3)], "itemsize": 16})
    nodes = np.zeros(n_nodes, dtype=node_dt)
    hdrs = np.zeros(n_hdr, dtype=hdr_dt)
    quant = np.zeros(9, dtype=np.float32)
    info = np.zeros(n_tris, dtype=info_dt)
    L.ref_bvh_dump(mode, mesh, _ptr(nodes), _ptr(hdrs), _ptr(quant), _ptr(info))
    return dict(nodes=nodes, headers=hdrs, quant=quant, triinfo=info, n_verts=n_verts, n_info=n_info)


def bench(mode: int, cars_per_team: int, n_threads: int, arenas_per_thread: int, ticks: int, seed: int = 1234):
    wall = C.c_double(0)
    v = lib().ref_bench(mode, cars_per_team, n_threads, arenas_per_thread, ticks, seed, C.byref(wall))
    return v, wall.value

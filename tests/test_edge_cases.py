"""Edge cases of the boundary: loader errors (the reference's CollisionMeshFile validation, src/CollisionMeshFile/
CollisionMeshFile.cpp:22-55), empty and degenerate calls, partial resets, event buffer limits, and non-default tick rates
(Arena.cpp:419 allows 15..120 Hz; below 80 Hz the vehicle's rolling-friction rounding is active, btVehicleRL.cpp:355-359)."""
import ctypes as C
import struct
import subprocess
import sys

import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import lib as rlib, types as T
from rocketsim_b200.arena_batch import ArenaBatch

LOADER = r"""
import sys, ctypes as C
sys.path.insert(0, {root!r})
from rocketsim_b200 import lib as rlib
L = rlib.load({lib!r})
data = bytes.fromhex({hexdata!r})
buf = C.create_string_buffer(data, len(data))
rc = L.rsb_init_from_mem(1, (C.c_int * 1)(0), (C.c_void_p * 1)(C.cast(buf, C.c_void_p).value), (C.c_size_t * 1)(len(data)))
print(rc, L.rsb_last_error().decode())
"""


def load_in_fresh_process(data: bytes):
    src = LOADER.format(root=pu.ROOT, lib=pu.HOSTCHECK_LIB, hexdata=data.hex())
    r = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, timeout=120)
    rc, _, msg = r.stdout.strip().partition(" ")
    return int(rc), msg


def cmf(tris, verts):
    out = struct.pack("<ii", len(tris), len(verts))
    for t in tris:
        out += struct.pack("<iii", *t)
    for v in verts:
        out += struct.pack("<fff", *v)
    return out


def test_invalid_mesh_files_are_rejected_like_the_reference(hostlib):
    good = cmf([(0, 1, 2)], [(0, 0, 0), (1, 0, 0), (0, 1, 0)])
    assert load_in_fresh_process(good)[0] == 0
    rc, msg = load_in_fresh_process(struct.pack("<ii", 0, 3) + b"\0" * 36)
    assert rc != 0 and "bad triangle/vertex count" in msg
    rc, msg = load_in_fresh_process(struct.pack("<ii", 2_000_000, 3))
    assert rc != 0 and "bad triangle/vertex count" in msg
    rc, msg = load_in_fresh_process(good[:-5])
    assert rc != 0 and "overflown" in msg
    rc, msg = load_in_fresh_process(cmf([(0, 1, 7)], [(0, 0, 0), (1, 0, 0), (0, 1, 0)]))
    assert rc != 0 and "bad triangle vertex index" in msg


def test_degenerate_calls(hostlib):
    b = ArenaBatch(T.MODE_SOCCAR, 3, [0, 1], lib=hostlib)
    before = b.get_car_states().copy()
    b.step(0)
    b.step(-4)
    assert np.all(b.get_tick_counts() == 0) and pu.struct_equal(b.get_car_states(), before)
    b.set_car_states([], [], np.zeros(0, T.CarState))
    b.set_ball_states([], np.zeros(0, T.BallState))
    # partial reset: only the listed arena changes
    b.reset_to_random_kickoff([5], arena_idx=[1])
    after = b.get_car_states()
    assert pu.struct_equal(after[0], before[0]) and pu.struct_equal(after[2], before[2]) and not pu.struct_equal(after[1], before[1])
    with pytest.raises(rlib.RsbError):
        b.reset_to_random_kickoff([5], arena_idx=[3])
    with pytest.raises(rlib.RsbError):
        b.set_car_states([0], [2], np.zeros(1, T.CarState))
    # zero-car batch: every per-car call is a no-op, the ball still steps
    z = ArenaBatch(T.MODE_SOCCAR, 2, [], lib=hostlib)
    z.set_controls(np.zeros((2, 0), T.CarControls))
    z.step(3)
    assert z.get_car_states().shape == (2, 0) and np.all(z.get_tick_counts() == 3)
    z.close(); b.close()


def test_event_buffer_cap(hostlib):
    b = ArenaBatch(T.MODE_SOCCAR, 4, [0, 1], lib=hostlib)
    b.reset_to_random_kickoff(np.arange(4, dtype=np.int32))
    bs = b.get_ball_states()
    bs["pos"]["x"] = 600; bs["pos"]["y"] = 4900; bs["pos"]["z"] = 200; bs["vel"]["y"] = 5000
    b.set_ball_states(np.arange(4), bs)
    b.step(30)
    # Without a reset the goal fires on EVERY tick the ball stays in the goal, as in the reference (Arena.cpp:744-752 calls the
    # goal callback whenever IsBallScored()).  Fewer slots than events: the C call hands out `cap` at a time and reports that more
    # are pending (rc 1) -- nothing is dropped; the Python front end loops.  Sorted by (arena, tick).
    import ctypes as C
    out = np.zeros(2, T.Event); n = C.c_int()
    assert hostlib.rsb_get_events(b.h, rlib.ptr(out), 2, C.byref(n)) == 1 and n.value == 2 and b"pending" in hostlib.rsb_last_error()
    assert list(out["arena"]) == [0, 0] and np.all(out["type"] == T.EVENT_GOAL) and out["tick"][1] == out["tick"][0] + 1
    ev = b.get_events(cap=7)
    assert len(ev) > 7 and np.all(ev["type"] == T.EVENT_GOAL) and ev["tick"][0] == out["tick"][1] + 1
    assert np.all(np.diff(ev["arena"].astype(np.int64)) >= 0) and sorted(set(ev["arena"])) == [0, 1, 2, 3]
    assert len(b.get_events()) == 0    # the buffer is drained by the call
    assert np.all(b.is_ball_scored())  # no auto reset: the ball stays in the goal (the reference leaves resets to the caller)
    b.close()


def run_tickrate(lib, oracle, rate, ticks=240):
    teams = [0, 1]
    ref = oracle.RefArena(T.MODE_SOCCAR, tick_rate=rate)
    for t in teams:
        ref.add_car(t)
    batch = ArenaBatch(T.MODE_SOCCAR, 2, teams, tick_rate=rate, lib=lib)
    ref.seed_rng(9); ref.reset_kickoff(9)
    sched = pu.ControlSchedule(2, 9)
    worst, mism = 0.0, []
    for t in range(ticks):
        for a in range(2):
            pu.sync_from_ref(batch, ref, a); batch.seed_rng(a, 9 + t)
        ref.seed_rng(9 + t)
        c = sched.at(t)
        if t < 60:
            c["throttle"] = 1; c["boost"] = 1; c["steer"] = 0
        if 100 < t < 160:
            c["throttle"] = 0; c["boost"] = 0; c["handbrake"] = 0; c["jump"] = 0  # coasting + braking: the sub-80 Hz rolling-friction path
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, c[i])
        batch.set_controls(np.broadcast_to(c, (2, 2)))
        ref.step(1); batch.step(1)
        cs = batch.get_car_states()
        ill = any(np.any((ref.wheel_debug(cid)[:, 0] != 0) & (ref.wheel_debug(cid)[:, 1] == 0)) for cid in ref.car_ids)
        for i, cid in enumerate(ref.car_ids):
            bad, errs = pu.compare_car(cs[1, i], ref.car_state(cid))
            if bad:
                mism.append((t, i, bad))
            if not ill:
                worst = max(worst, max(errs.values()))
    err = int(batch.get_error_flags().max())
    batch.close(); ref.close()
    assert err == 0 and not mism, (err, mism[:3])
    return worst


@pytest.mark.parametrize("rate", [60.0, 30.0, 15.0])
def test_tick_rates_hostcheck(hostlib, oracle, rate):
    assert run_tickrate(hostlib, oracle, rate) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [90.0, 60.0, 15.0])
def test_tick_rates_cuda(cudalib, oracle, rate):
    assert run_tickrate(cudalib, oracle, rate) <= 1e-4

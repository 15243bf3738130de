"""Size-independent properties of the batched step at BASELINE.json's full sizes (4,096 arenas on one GPU).

The oracle steps one arena at a time, so at full size parity is carried by properties that tie the big batch back to
the small, oracle-checked cases:
  * slot independence: arena i of a 4,096-arena batch evolves bit-identically to the same arena alone in a batch of 1
    (different chunk, different SM, different thread mapping);
  * replicas: arenas given the same seed and controls stay bit-identical to each other;
  * launch granularity: one launch of n ticks == n launches of one tick, bit for bit;
  * determinism run to run; exact tick counters (no skipped work); no device error flag; states finite and inside the arena.
"""
import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import lib as lib_mod, types as T
from rocketsim_b200.arena_batch import ArenaBatch

N_FULL = 4096


def bench_controls(n_ticks, n_arenas, n_cars, seed, arena0=0):
    """README-bench control streams, one independent stream per (arena, car)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_ticks, n_arenas, n_cars), T.CarControls)
    for a in range(n_arenas):
        sched = pu.ControlSchedule(n_cars, seed * 7919 + arena0 + a)
        for t in range(n_ticks):
            out[t, a] = sched.at(t)
    del rng
    return out


def run(lib, mode, teams, n_arenas, seeds, ctrl, chunks):
    """kickoff from `seeds`, then step through ctrl [T, n_arenas, n_cars] in launches of the given sizes."""
    b = ArenaBatch(mode, n_arenas, teams, lib=lib)
    b.reset_to_random_kickoff(np.asarray(seeds, np.int32))
    t = 0
    for n in chunks:
        b.step_with_controls(ctrl[t:t + n])
        t += n
    res = dict(cars=b.get_car_states(), ball=b.get_ball_states(), pads=b.get_pad_states() if b.n_pads else None,
               ticks=b.get_tick_counts(), err=b.get_error_flags(), ci=b.get_car_internals())
    b.close()
    return res


def same(a, b):
    return pu.struct_equal(a, b)


@pytest.mark.gpu
def test_full_size_replicas_and_slot_independence(cudalib):
    mode, teams = T.MODE_SOCCAR, [0, 0, 1, 1]
    n_ticks, n_distinct = 240, 8
    # 4,096 arenas = 512 replicas of 8 distinct (seed, control stream) arenas, interleaved over all chunks
    base = bench_controls(n_ticks, n_distinct, len(teams), seed=5)
    for t in range(150):  # drive at the ball first: touches, bumps and pad pickups inside the window
        base[t]["throttle"] = 1; base[t]["boost"] = 1; base[t]["steer"] = 0; base[t]["jump"] = 0; base[t]["handbrake"] = 0
    idx = np.arange(N_FULL) % n_distinct
    ctrl = np.ascontiguousarray(base[:, idx])
    seeds = 1000 + idx
    big = run(cudalib, mode, teams, N_FULL, seeds, ctrl, [n_ticks])
    assert int(big["err"].max()) == 0
    assert np.all(big["ticks"] == n_ticks)
    for k in range(n_distinct):
        rep = np.nonzero(idx == k)[0]
        for key in ("cars", "ball", "pads", "ci"):
            first = big[key][rep[0]]
            assert all(same(big[key][r], first) for r in rep[1:]), f"replicas of arena {k} diverged in {key}"
    # the same arenas alone in a batch of n_distinct (other chunking, other SMs): bit-identical
    small = run(cudalib, mode, teams, n_distinct, 1000 + np.arange(n_distinct), base, [n_ticks])
    for k in range(n_distinct):
        for key in ("cars", "ball", "pads", "ci"):
            assert same(small[key][k], big[key][k]), f"arena {k}: batch-of-{n_distinct} != batch-of-{N_FULL} in {key}"
    # sanity: the arenas did something and stayed inside the field
    pos = big["cars"]["pos"]
    assert np.isfinite(pos["x"]).all() and np.isfinite(big["ball"]["pos"]["z"]).all()
    assert np.abs(pos["x"]).max() < 4300 and np.abs(pos["y"]).max() < 6100 and pos["z"].max() < 2100
    assert len({big["cars"][k].tobytes() for k in range(n_distinct)}) == n_distinct


@pytest.mark.gpu
@pytest.mark.parametrize("mode,cpt", [(T.MODE_SOCCAR, 3), (T.MODE_HOOPS, 2), (T.MODE_HEATSEEKER, 1), (T.MODE_SNOWDAY, 3), (T.MODE_THE_VOID, 1), (T.MODE_DROPSHOT, 3)],
                         ids=["soccar_3v3", "hoops_2v2", "heatseeker_1v1", "snowday_3v3", "void_1v1", "dropshot_3v3"])
def test_launch_granularity_and_determinism(cudalib, mode, cpt):
    teams = [0] * cpt + [1] * cpt
    n_arenas, n_ticks = 300, 120
    ctrl = bench_controls(n_ticks, n_arenas, len(teams), seed=11)
    for t in range(90):
        ctrl[t]["throttle"] = 1; ctrl[t]["boost"] = 1; ctrl[t]["steer"] = 0; ctrl[t]["jump"] = 0
    seeds = 77 + np.arange(n_arenas)
    one = run(cudalib, mode, teams, n_arenas, seeds, ctrl, [n_ticks])
    again = run(cudalib, mode, teams, n_arenas, seeds, ctrl, [n_ticks])
    split = run(cudalib, mode, teams, n_arenas, seeds, ctrl, [1] * 30 + [7, 13, 70])
    assert int(one["err"].max()) == 0 and np.all(one["ticks"] == n_ticks)
    for key in ("cars", "ball", "ci"):
        assert same(one[key], again[key]), f"run-to-run difference in {key}"
        assert same(one[key], split[key]), f"1 launch of {n_ticks} ticks != split launches in {key}"


@pytest.mark.gpu
def test_events_and_auto_reset_full_size(cudalib):
    """README bench loop on device: goals are followed by a kickoff reset inside the launch; every goal produces exactly one
    event and the scoring arena restarts from a kickoff (ball back at the centre spot within a few ticks of the event)."""
    mode, teams = T.MODE_SOCCAR, [0, 0, 1, 1]
    b = ArenaBatch(mode, N_FULL, teams, seed=9, lib=cudalib)
    b.reset_to_random_kickoff((9 + np.arange(N_FULL)).astype(np.int32))
    b.set_auto_reset(True, 9)
    # fire every ball at the orange goal
    bs = b.get_ball_states()
    bs["vel"]["y"] = 5000.0; bs["pos"]["z"] = 200.0; bs["pos"]["x"] = 600.0  # x = 600 clears every kickoff spot
    b.set_ball_states(np.arange(N_FULL), bs)
    b.step(240)
    ev = b.get_events(cap=4 * N_FULL)
    goals = ev[ev["type"] == 1]
    assert len(goals) >= N_FULL  # every arena scored at least once
    assert len(np.unique(goals["arena"])) == N_FULL
    assert np.all(goals["a"] == 0)  # ball crossed +y: team 0 (blue) scores
    assert int(b.get_error_flags().max()) == 0
    assert np.all(b.get_tick_counts() == 240)
    b.close()


def _queue_cap_run(lib, mode, teams, n_arenas, cap, ticks=150):
    """Cars driven into each other, the ball and the walls (GJK / EPA / box-box records every tick), with the usable capacity of the
    per-block detector queue set to `cap` (None = the product's 512)."""
    b = ArenaBatch(mode, n_arenas, teams, seed=5, lib=lib)
    if cap is not None:
        lib_mod.check(lib, lib.rsb_debug_set_detect_queue_cap(b.h, cap))
    b.reset_to_random_kickoff((5 + np.arange(n_arenas)).astype(np.int32))
    cs = b.get_car_states()
    n = len(teams)
    for i in range(n):  # a column of cars heading for the +x wall, the rear ones faster: they run into each other, then into the ramp
        cs[:, i]["pos"]["x"] = 1800.0 + 270.0 * i; cs[:, i]["pos"]["y"] = (np.arange(n_arenas) % 7) * 9.0 + 25.0 * (i % 2); cs[:, i]["pos"]["z"] = 17.0
        cs[:, i]["rotMat"]["forward"] = (1, 0, 0, 0); cs[:, i]["rotMat"]["right"] = (0, 1, 0, 0); cs[:, i]["rotMat"]["up"] = (0, 0, 1, 0)
        cs[:, i]["vel"]["x"] = 2200.0 - 300.0 * i; cs[:, i]["vel"]["y"] = 0.0
    ai = np.repeat(np.arange(n_arenas), n); ci = np.tile(np.arange(n), n_arenas)
    b.set_car_states(ai, ci, cs.ravel())
    bs = b.get_ball_states()
    bs["pos"]["x"] = 3500.0; bs["pos"]["y"] = 20.0; bs["pos"]["z"] = 120.0; bs["vel"]["x"] = 600.0
    b.set_ball_states(np.arange(n_arenas), bs)
    ctrl = np.zeros((ticks, n_arenas, n), T.CarControls); ctrl["throttle"] = 1; ctrl["boost"] = 1
    b.step_with_controls(ctrl)
    res = dict(cars=b.get_car_states(), ball=b.get_ball_states(), err=b.get_error_flags(), ev=len(b.get_events(cap=1 << 20)))
    b.close()
    return res


def _check_queue_overflow(lib, n_arenas):
    """The detector-queue overflow sweep (run_phase_block<PH_NARROW_DETECT> / arena_tick): with room for ONE record per block, all
    the others of a tick go through the sweep -- and nothing may change."""
    for mode, teams in ((T.MODE_SOCCAR, [0, 0, 0, 1, 1, 1]), (T.MODE_SNOWDAY, [0, 0, 1, 1])):
        full = _queue_cap_run(lib, mode, teams, n_arenas, None)
        tiny = _queue_cap_run(lib, mode, teams, n_arenas, 1)
        assert int(full["err"].max()) == 0 and int(tiny["err"].max()) == 0
        assert full["ev"] > 0, "no bump in the whole run: the scenario produces no car-car records"
        assert same(full["cars"], tiny["cars"]) and same(full["ball"], tiny["ball"]), f"mode {mode}: results depend on the detector queue capacity"


def test_hostcheck_detector_queue_overflow_sweep(hostlib):
    _check_queue_overflow(hostlib, 3)


@pytest.mark.gpu
def test_cuda_detector_queue_overflow_sweep(cudalib):
    _check_queue_overflow(cudalib, 300)

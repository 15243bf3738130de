"""Car presets and mutators through the boundary: every CarConfig preset (src/Sim/Car/CarConfig/CarConfig.cpp:20-100) and a
non-default MutatorConfig, single-step parity against the live oracle (resynchronised every tick)."""
import ctypes as C

import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import lib as rlib, types as T

TOL = TOL_ILLCOND = 1e-4


def preset_cfg(L, k):
    c = np.zeros(1, T.CarConfig)
    rlib.check(L, L.rsb_car_config_preset(k, rlib.ptr(c)))
    return c[0]


def test_constant_tables_match_reference(hostlib, oracle):
    for k in range(7):
        assert pu.struct_equal(preset_cfg(hostlib, k), oracle.car_config(k)), f"preset {k}"
    for mode in (T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_HEATSEEKER, T.MODE_SNOWDAY, T.MODE_DROPSHOT, T.MODE_THE_VOID):
        m = np.zeros(1, T.MutatorConfig)
        rlib.check(hostlib, hostlib.rsb_default_mutator_config(mode, rlib.ptr(m)))
        assert pu.struct_equal(m[0], oracle.default_mutator_config(mode)), f"mutator of mode {mode}"


def run_presets(lib, oracle, presets, ticks, seed, mutate=False):
    teams = [i % 2 for i in range(len(presets))]
    ref = oracle.RefArena(T.MODE_SOCCAR)
    for t, p in zip(teams, presets):
        ref.add_car(t, p)
    mut = None
    if mutate:
        mut = oracle.default_mutator_config(T.MODE_SOCCAR).copy()
        mut["gravity"]["z"] = -400.0; mut["boostAccelGround"] *= 1.5; mut["boostAccelAir"] *= 1.5
        mut["jumpImmediateForce"] *= 1.2; mut["ballDrag"] = 0.05; mut["bumpCooldownTime"] = 0.1
        mut["ballWorldRestitution"] = 0.8; mut["unlimitedFlips"] = 1; mut["boostUsedPerSecond"] = 10.0
        ref.set_mutator_config(mut)
    cfgs = np.array([preset_cfg(lib, p) for p in presets], T.CarConfig)
    batch = pu.ArenaBatch(T.MODE_SOCCAR, 2, teams, car_configs=cfgs, mutator_config=mut, lib=lib)
    ref.seed_rng(seed); ref.reset_kickoff(seed)
    sched = pu.ControlSchedule(len(teams), seed)
    worst, worst_ill, mism = 0.0, 0.0, []
    for t in range(ticks):
        for a in range(2):
            pu.sync_from_ref(batch, ref, a); batch.seed_rng(a, seed + t)
        ref.seed_rng(seed + t)
        c = sched.at(t)
        if t < 100:
            c["throttle"] = 1; c["boost"] = 1; c["steer"] = 0; c["handbrake"] = 0
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, c[i])
        batch.set_controls(np.broadcast_to(c, (2, len(teams))))
        ref.step(1); batch.step(1)
        ill = any(np.any((ref.wheel_debug(cid)[:, 0] != 0) & (ref.wheel_debug(cid)[:, 1] == 0)) for cid in ref.car_ids)
        cs, bs = batch.get_car_states(), batch.get_ball_states()
        e = 0.0
        for i, cid in enumerate(ref.car_ids):
            bad, errs = pu.compare_car(cs[1, i], ref.car_state(cid))
            if bad:
                mism.append((t, i, bad))
            e = max(e, max(errs.values()))
        e = max(e, max(pu.compare_ball(bs[1], ref.ball_state()).values()))
        if ill:
            worst_ill = max(worst_ill, e)
        else:
            worst = max(worst, e)
    err = int(batch.get_error_flags().max())
    batch.close(); ref.close()
    assert err == 0 and not mism, (err, mism[:4])
    assert worst <= TOL, worst
    assert worst_ill <= pu.illcond_tol(oracle, TOL_ILLCOND), worst_ill


def test_presets_hostcheck(hostlib, oracle):
    run_presets(hostlib, oracle, [1, 2, 3, 4, 5, 6], ticks=200, seed=3)


def test_mutators_hostcheck(hostlib, oracle):
    run_presets(hostlib, oracle, [0, 1], ticks=200, seed=5, mutate=True)


@pytest.mark.gpu
def test_presets_cuda(cudalib, oracle):
    run_presets(cudalib, oracle, [1, 2, 3, 4, 5, 6], ticks=300, seed=3)


@pytest.mark.gpu
def test_mutators_cuda(cudalib, oracle):
    run_presets(cudalib, oracle, [0, 6, 3, 2], ticks=300, seed=5, mutate=True)

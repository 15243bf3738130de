"""Non-default ArenaConfig through the boundary (src/Sim/Arena/ArenaConfig/ArenaConfig.h:18-50, Arena.cpp:416-545):
noBallRot = false (the ball's angular velocity is integrated into its orientation, Ball.cpp:53-115) and custom boost pads
(Arena.cpp:498-512 pad construction, :724-730 the all-pads check replacing the grid).  Single-step parity against the live
oracle, resynchronised every tick, pads compared bitwise; unsupported fields are rejected with an error naming the field."""
import ctypes as C

import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import arena_batch as ab, lib as rlib, types as T

TOL = 1e-4

# pads on the kickoff approach lines and around the centre circle: the README-style drive picks several of them up
CUSTOM_PADS = np.array([
    [0, -2300, 70, 0], [0, 2300, 70, 0], [-256, -3400, 70, 1], [256, 3400, 70, 1], [-1200, -1500, 70, 0], [1200, 1500, 70, 0],
    [-1800, -2300, 73, 1], [1800, 2300, 73, 1], [600, -600, 70, 0], [-600, 600, 70, 0], [0, -1000, 70, 1],
], dtype=np.float32)


def pad_cfgs(p):
    out = np.zeros(len(p), T.BoostPadConfig)
    out["pos"]["x"], out["pos"]["y"], out["pos"]["z"] = p[:, 0], p[:, 1], p[:, 2]
    out["isBig"] = p[:, 3] != 0
    return out


def run_config(lib, oracle, mode, teams, ticks, seed, no_ball_rot=True, pads=None, aim=False):
    ref = oracle.RefArena(mode, no_ball_rot=no_ball_rot, custom_pads=pads)
    for t in teams:
        ref.add_car(t)
    ac = ab.default_arena_config(lib).copy()
    ac["noBallRot"] = int(no_ball_rot)
    batch = pu.ArenaBatch(mode, 2, teams, lib=lib, arena_config=ac, custom_boost_pads=None if pads is None else pad_cfgs(pads))
    assert batch.n_pads == ref.num_pads
    ref.seed_rng(seed); ref.reset_kickoff(seed)
    if aim:
        pu.aim_cars_at_ball(ref, radius=700.0)
    sched = pu.ControlSchedule(len(teams), seed)
    worst, mism, pickups, spun = 0.0, [], 0, False
    for t in range(ticks):
        for a in range(2):
            pu.sync_from_ref(batch, ref, a); batch.seed_rng(a, seed + t)
        ref.seed_rng(seed + t)
        c = sched.at(t)
        if t < 150:
            c["throttle"] = 1; c["boost"] = 1; c["steer"] = 0; c["handbrake"] = 0; c["jump"] = 0
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
        rb = ref.ball_state()
        e = max(e, max(pu.compare_ball(bs[1], rb).values()))
        if not ill:
            worst = max(worst, e)
        if batch.n_pads:
            mine, theirs = batch.get_pad_states()[1], ref.pad_states()
            if not pu.struct_equal(mine, theirs):
                mism.append((t, "pads"))
            pickups += int(np.sum(theirs["isActive"] == 0))
        rm = pu.field_as_array(rb, "rotMat").reshape(-1)
        spun = spun or abs(float(rm[0]) - 1.0) > 1e-3
    err = int(batch.get_error_flags().max())
    batch.close(); ref.close()
    assert err == 0 and not mism, (err, mism[:4])
    assert worst <= TOL, worst
    return dict(pickups=pickups, spun=spun)


def check_rejections(lib):
    d = ab.default_arena_config(lib)
    assert (int(d["noBallRot"]), int(d["useCustomBroadphase"]), float(d["maxAABBLen"]), int(d["maxObjects"])) == (1, 1, 370.0, 512)
    for field, value in (("memWeightMode", 1), ("maxAABBLen", 500.0), ("useCustomBroadphase", 0)):
        ac = d.copy(); ac[field] = value
        with pytest.raises(RuntimeError, match=field):
            pu.ArenaBatch(T.MODE_SOCCAR, 1, [0, 1], lib=lib, arena_config=ac)
    ac = d.copy(); ac["maxPos"]["z"] = 3000.0
    with pytest.raises(RuntimeError, match="minPos / maxPos"):
        pu.ArenaBatch(T.MODE_SOCCAR, 1, [0, 1], lib=lib, arena_config=ac)
    with pytest.raises(RuntimeError, match="customBoostPads"):
        pu.ArenaBatch(T.MODE_SOCCAR, 1, [0, 1], lib=lib, custom_boost_pads=pad_cfgs(np.zeros((41, 4), np.float32)))


def test_arena_config_defaults_and_rejections_hostcheck(hostlib):
    check_rejections(hostlib)


def test_custom_boost_pads_hostcheck(hostlib, oracle):
    r = run_config(hostlib, oracle, T.MODE_SOCCAR, [0, 1, 0, 1], ticks=360, seed=4, pads=CUSTOM_PADS)
    assert r["pickups"] > 0  # the scenario does collect custom pads


def test_no_custom_pads_at_all_hostcheck(hostlib, oracle):
    run_config(hostlib, oracle, T.MODE_SOCCAR, [0, 1], ticks=120, seed=2, pads=np.zeros((0, 4), np.float32))


def test_ball_rotation_hostcheck(hostlib, oracle):
    r = run_config(hostlib, oracle, T.MODE_SOCCAR, [0, 1], ticks=300, seed=6, no_ball_rot=False, aim=True)
    assert r["spun"]  # the ball was hit and its orientation left identity


@pytest.mark.gpu
def test_arena_config_cuda(cudalib, oracle):
    check_rejections(cudalib)
    assert run_config(cudalib, oracle, T.MODE_SOCCAR, [0, 1, 0, 1], ticks=480, seed=4, pads=CUSTOM_PADS)["pickups"] > 0
    assert run_config(cudalib, oracle, T.MODE_SOCCAR, [0, 1], ticks=300, seed=6, no_ball_rot=False, aim=True)["spun"]
    run_config(cudalib, oracle, T.MODE_HOOPS, [0, 1], ticks=300, seed=6, no_ball_rot=False, aim=True)
    run_config(cudalib, oracle, T.MODE_SOCCAR, [0, 1], ticks=120, seed=2, pads=np.zeros((0, 4), np.float32))

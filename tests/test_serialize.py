"""Wire compatibility with the CPU library: the byte stream of Arena::Serialize
(src/Sim/Arena/Arena.cpp:546-574).

  * reference -> batch: a stream written by the reference loads into an arena of the batch; afterwards the batch steps
    exactly like the reference arena that Arena::DeserializeNew builds from the same bytes (single-step bar);
  * batch -> reference: a stream written by the batch loads with the reference's own Arena::DeserializeNew, reproduces the
    reference's Serialize() of that arena byte for byte, and both sides step alike.
"""
import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import lib as rlib, types as T
from rocketsim_b200.arena_batch import ArenaBatch

TOL = 1e-4


def advance(ref, ticks, seed):
    sched = pu.ControlSchedule(len(ref.car_ids), seed)
    for t in range(ticks):
        c = sched.at(t)
        if t < 120:
            c["throttle"] = 1; c["boost"] = 1; c["steer"] = 0
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, c[i])
        ref.step(1)
    return c


def compare_step(batch, arena, ref, ctrl, ticks=30):
    worst = 0.0
    for t in range(ticks):
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, ctrl[i])
        batch.set_controls(np.broadcast_to(ctrl, (batch.n_arenas, len(ref.car_ids))))
        ref.step(1); batch.step(1)
        cs, bs = batch.get_car_states(), batch.get_ball_states()
        for i, cid in enumerate(ref.car_ids):
            bad, errs = pu.compare_car(cs[arena, i], ref.car_state(cid))
            assert not bad, (t, i, bad)
            worst = max(worst, max(errs.values()))
        worst = max(worst, max(pu.compare_ball(bs[arena], ref.ball_state()).values()))
        # free-running from identical (deserialised) states: resync so that every tick is a single-step case
        pu.sync_from_ref(batch, ref, arena)
    return worst


def run_case(lib, oracle, mode, cpt, presets=None):
    teams = [0] * cpt + [1] * cpt
    presets = presets or [0] * len(teams)
    ref = oracle.RefArena(mode)
    for t, p in zip(teams, presets):
        ref.add_car(t, p)
    ref.seed_rng(3); ref.reset_kickoff(3)
    ctrl = advance(ref, 200, seed=3)
    cfgs = np.zeros(len(teams), T.CarConfig)
    for i, p in enumerate(presets):
        rlib.check(lib, lib.rsb_car_config_preset(p, rlib.ptr(cfgs[i:i + 1])))
    batch = ArenaBatch(mode, 3, teams, car_configs=cfgs, lib=lib)

    # ---- reference -> batch: the loaded arena holds exactly what the reference arena holds
    blob = ref.serialize()
    batch.deserialize_arena(1, blob)
    assert int(batch.get_tick_counts()[1]) == ref.tick_count
    cs = batch.get_car_states()
    # CARSTATE_SERIALIZATION_FIELDS (Car.h:122-128) omits the first three; boost is in the stream but Arena::DeserializeNewCar
    # overwrites it with the spawn amount (Car::_BulletSetup runs after the read, Car.cpp:291) -- mirrored
    NOT_IN_STREAM = {"wheelsWithContact", "isSupersonic", "airTime", "boost"}
    for i, cid in enumerate(ref.car_ids):
        bad, errs = pu.compare_car(cs[1, i], ref.car_state(cid))
        assert not (set(bad) - NOT_IN_STREAM), (i, bad)
        # positions travel as UU floats and are stored in Bullet units: one rounding of x50 / x0.02
        assert max(v for k, v in errs.items() if k not in NOT_IN_STREAM) <= 1e-6, (i, errs)
    assert max(pu.compare_ball(batch.get_ball_states()[1], ref.ball_state()).values()) <= 1e-6
    if batch.n_pads:
        assert pu.struct_equal(batch.get_pad_states()[1], ref.pad_states())

    # ---- batch -> reference.  Our stream lists the cars by ascending id.  (The reference writes them in unordered_set order,
    # and its own DeserializeNew loses cars when that order is not ascending: forcing a car's id erases the _carIDMap entry of
    # the car that was loaded under that temporary id, Arena.cpp:606-611 -- so the reference cannot always reload its own stream.)
    mine = batch.serialize_arena(1)
    back = oracle.RefArena.deserialize_new(mine, mode, len(teams))
    # the reference re-serialises in its own (unordered) car order: compare what the stream MEANS by loading it into arena 2
    batch.deserialize_arena(2, back.serialize())
    cs2 = batch.get_car_states()
    for i in range(len(teams)):   # (each UU <-> Bullet-unit conversion may move the last bit of a position)
        bad, errs = pu.compare_car(cs2[2, i], cs2[1, i])
        assert not bad and max(errs.values()) <= 1e-6, (i, bad, errs)
    assert max(pu.compare_ball(batch.get_ball_states()[2], batch.get_ball_states()[1]).values()) <= 1e-6
    for i, cid in enumerate(back.car_ids):
        bad, errs = pu.compare_car(batch.get_car_states()[1, i], back.car_state(cid))
        assert not bad and max(errs.values()) == 0.0
    # both sides now hold a freshly deserialised arena (new physics world, default hidden state): they must step alike
    worst = compare_step(batch, 1, back, ctrl)
    for r in (ref, back):
        r.close()
    batch.close()
    return worst


def run_clone(lib, oracle, mode, cpt):
    """Arena::Clone: clone a mid-game arena on both sides, then the clones must step alike (and the originals too)."""
    teams = [0] * cpt + [1] * cpt
    ref = pu.make_ref_arena(oracle, mode, teams)
    ref.seed_rng(6); ref.reset_kickoff(6)
    ctrl = advance(ref, 150, seed=6)
    batch = ArenaBatch(mode, 3, teams, lib=lib)
    pu.sync_from_ref(batch, ref, 0)
    ref2 = ref.clone()
    batch.clone_arena(0, 2)
    assert int(batch.get_tick_counts()[2]) == int(batch.get_tick_counts()[0])
    cs = batch.get_car_states()
    for i, cid in enumerate(ref2.car_ids):
        bad, errs = pu.compare_car(cs[2, i], ref2.car_state(cid))
        assert not (set(bad) - {"wheelsWithContact", "isSupersonic"}) and max(v for k, v in errs.items() if k != "airTime") <= 1e-6, (i, bad, errs)
    worst = compare_step(batch, 2, ref2, ctrl)
    ref.close(); ref2.close(); batch.close()
    return worst


def test_clone_hostcheck(hostlib, oracle):
    assert run_clone(hostlib, oracle, T.MODE_SOCCAR, 2) <= TOL
    assert run_clone(hostlib, oracle, T.MODE_HOOPS, 1) <= TOL


@pytest.mark.gpu
def test_clone_cuda(cudalib, oracle):
    assert run_clone(cudalib, oracle, T.MODE_SOCCAR, 3) <= TOL
    assert run_clone(cudalib, oracle, T.MODE_SNOWDAY, 1) <= TOL


def test_serialize_hostcheck(hostlib, oracle):
    assert run_case(hostlib, oracle, T.MODE_SOCCAR, 2, presets=[0, 1, 3, 5]) <= TOL
    assert run_case(hostlib, oracle, T.MODE_HOOPS, 1) <= TOL
    assert run_case(hostlib, oracle, T.MODE_THE_VOID, 1) <= TOL
    assert run_case(hostlib, oracle, T.MODE_DROPSHOT, 1) <= TOL   # BallState::dsInfo is in the stream, tile states are not


def test_deserialize_rejects_mismatches(hostlib, oracle):
    ref = oracle.RefArena(T.MODE_SOCCAR)
    ref.add_car(0); ref.add_car(1)
    blob = ref.serialize()
    b3 = ArenaBatch(T.MODE_SOCCAR, 1, [0, 0, 1], lib=hostlib)
    with pytest.raises(rlib.RsbError, match="cars"):
        b3.deserialize_arena(0, blob)
    bh = ArenaBatch(T.MODE_HOOPS, 1, [0, 1], lib=hostlib)
    with pytest.raises(rlib.RsbError, match="game mode"):
        bh.deserialize_arena(0, blob)
    bt = ArenaBatch(T.MODE_SOCCAR, 1, [1, 0], lib=hostlib)
    with pytest.raises(rlib.RsbError, match="team"):
        bt.deserialize_arena(0, blob)
    ok = ArenaBatch(T.MODE_SOCCAR, 1, [0, 1], lib=hostlib)
    with pytest.raises(rlib.RsbError, match="overflown"):
        ok.deserialize_arena(0, blob[:-7])
    with pytest.raises(rlib.RsbError):
        ok.deserialize_arena(0, b"\x05\x00\x00\x00" + blob[4:])   # wrong field count in the arena header
    ok.deserialize_arena(0, blob)
    for b in (b3, bh, bt, ok):
        b.close()
    ref.close()


@pytest.mark.gpu
def test_serialize_cuda(cudalib, oracle):
    assert run_case(cudalib, oracle, T.MODE_SOCCAR, 3, presets=[0, 1, 2, 3, 4, 6]) <= TOL
    assert run_case(cudalib, oracle, T.MODE_SNOWDAY, 2) <= TOL
    assert run_case(cudalib, oracle, T.MODE_HEATSEEKER, 1) <= TOL
    assert run_case(cudalib, oracle, T.MODE_DROPSHOT, 2) <= TOL

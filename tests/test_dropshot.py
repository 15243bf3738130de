"""Dropshot: hexagonal floor tiles as static convex hulls that only the ball collides with
(reference src/Sim/Arena/DropshotTiles/DropshotTiles.cpp:22-130, Arena.cpp:1070-1085), the DROPSHOT_FLOOR plane for cars and
wheel rays (Arena.cpp:1051, Car.cpp:227-233), tile damage / charge levels (Ball.cpp:299-323, 364-409), the deferred tile-body
update (Arena.cpp:741-746), the kickoff launch (Ball.cpp:209-233), the kickoff table (RLConst.h:381-388) and the goal test
(Arena.cpp:916-923).  The golden records tests/golden/dropshot_*.npz are covered by test_golden.py / test_bitexact.py; this
module adds free-running bit-exact rollouts of the Dropshot-specific scenarios, tile API round trips and a few properties."""
import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import types as T
from test_bitexact import free_run_identical, require_reference_platform

DS_VARIANTS = ["ds_drop", "ds_slam", "ds_charge", "walls", False, "charge"]


def _ds_sweep(lib, oracle, cpt, seeds, ticks):
    n = 0
    for s in seeds:
        n += free_run_identical(lib, oracle, T.MODE_DROPSHOT, cpt, 2000 + 31 * s, ticks, DS_VARIANTS[s % len(DS_VARIANTS)])
    return n


def test_hostcheck_dropshot_scenarios_bit_identical(hostlib, oracle):
    """CPU tier: the six Dropshot start variants x 360 free-running ticks, every CarState / BallState (incl. dsInfo) / tile bit-identical."""
    require_reference_platform(oracle)
    assert _ds_sweep(hostlib, oracle, 2, range(6), 360) >= 0.85 * 6 * 360


@pytest.mark.gpu
def test_cuda_dropshot_scenarios_bit_identical(cudalib, oracle):
    """GPU: 12 seeds (each start variant twice) x 520 ticks of 2v2 and 6 seeds x 400 ticks of 3v3."""
    require_reference_platform(oracle)
    n = _ds_sweep(cudalib, oracle, 2, range(12), 520) + _ds_sweep(cudalib, oracle, 3, range(6), 400)
    print(f"\ndropshot: {n} free-running steps bit-identical to the reference")
    assert n >= 0.8 * (12 * 520 + 6 * 400)


def _tile_api(lib):
    b = pu.ArenaBatch(T.MODE_DROPSHOT, 3, [0, 1], lib=lib)
    assert b.n_pads == 0  # no boost pads in Dropshot (Arena.cpp:498)
    assert not b.get_dropshot_tiles().any()
    bs = b.get_ball_states()
    assert (bs["dsInfo"]["chargeLevel"] == 1).all() and not bs["dsInfo"]["hasDamaged"].any()
    st = np.zeros((2, 2, 70), np.uint8)
    st[0, 1, 5] = 2; st[0, 0, 69] = 1; st[1, 1, :] = 2
    b.set_dropshot_tiles([2, 0], st)
    got = b.get_dropshot_tiles()
    assert np.array_equal(got[2], st[0]) and np.array_equal(got[0], st[1]) and not got[1].any()
    # BallState::dsInfo round trip
    bs[1]["dsInfo"]["chargeLevel"] = 3; bs[1]["dsInfo"]["accumulatedHitForce"] = 12345.5; bs[1]["dsInfo"]["yTargetDir"] = -1
    bs[1]["dsInfo"]["hasDamaged"] = 1; bs[1]["dsInfo"]["lastDamageTick"] = 77
    b.set_ball_states([1], bs[1:2])
    assert pu.struct_equal(b.get_ball_states()[1]["dsInfo"], bs[1]["dsInfo"])
    # ResetToRandomKickoff restores every tile and the ball's dsInfo (Arena.cpp:205-217)
    b.reset_to_random_kickoff([5, 6, 7])
    assert not b.get_dropshot_tiles().any()
    d = b.get_ball_states()["dsInfo"]
    assert (d["chargeLevel"] == 1).all() and not d["hasDamaged"].any() and (d["accumulatedHitForce"] == 0).all()
    with pytest.raises(Exception):
        b.set_dropshot_tiles([0], np.full((1, 2, 70), 3, np.uint8))
    # Arena::IsBallProbablyGoingIn (and with it the GameEventTracker) has no Dropshot case in the reference (Arena.cpp:775-889)
    with pytest.raises(Exception):
        b.is_ball_probably_going_in()
    with pytest.raises(Exception):
        b.set_event_tracker(True)
    b.close()
    s = pu.ArenaBatch(T.MODE_SOCCAR, 1, [0, 1], lib=lib)
    with pytest.raises(Exception):
        s.get_dropshot_tiles()
    with pytest.raises(Exception):
        s.set_dropshot_tiles([0], np.zeros((1, 2, 70), np.uint8))
    s.close()


def test_hostcheck_tile_api(hostlib):
    _tile_api(hostlib)


@pytest.mark.gpu
def test_cuda_tile_api(cudalib):
    _tile_api(cudalib)


def _ball_and_tiles(lib):
    """The ball rests on an intact tile, falls through a broken one (= a goal), and cars stand on the 1.5 uu floor plane
    whatever the tile under them looks like."""
    b = pu.ArenaBatch(T.MODE_DROPSHOT, 2, [0, 1], lib=lib)
    st = np.zeros((1, 2, 70), np.uint8); st[0, 1, :] = 2  # arena 1: the whole orange half is broken
    b.set_dropshot_tiles([1], st)
    bs = b.get_ball_states()
    for a in range(2):
        bs[a]["pos"]["x"], bs[a]["pos"]["y"], bs[a]["pos"]["z"] = 300.0, 1500.0, 200.0
        bs[a]["vel"]["z"] = -100.0
    b.set_ball_states([0, 1], bs)
    b.step(240)
    z = b.get_ball_states()["pos"]["z"]
    sc = b.is_ball_scored()
    assert 80.0 < z[0] < 120.0 and not sc[0]          # resting on the tile (radius 91.25 over the tile top at z = 0)
    assert z[1] < -1.75 * 91.25 and sc[1]              # fell through: Arena::IsBallScored (Arena.cpp:916-923)
    cz = b.get_car_states()["pos"]["z"]
    assert np.all(np.abs(cz - cz[0, 0]) < 1e-3) and 15.0 < cz[0, 0] < 20.0
    assert int(b.get_error_flags().max()) == 0
    b.close()


def test_hostcheck_ball_and_tiles(hostlib):
    _ball_and_tiles(hostlib)


@pytest.mark.gpu
def test_cuda_ball_and_tiles(cudalib):
    _ball_and_tiles(cudalib)


def _clone(lib, oracle):
    """Arena::Clone copies BallState::dsInfo through GetState / SetState but NOT the tile states (Arena.cpp:646-674)."""
    teams = [0, 1]
    ref = pu.make_ref_arena(oracle, T.MODE_DROPSHOT, teams)
    ref.seed_rng(9); ref.reset_kickoff(9); pu.canonical_kickoff(ref, teams)
    pu.apply_scenario(ref, T.MODE_DROPSHOT, teams, "ds_slam")
    b = pu.ArenaBatch(T.MODE_DROPSHOT, 2, teams, lib=lib)
    pu.sync_from_ref(b, ref, 0)
    for _ in range(60):
        ref.step(1)
    b.step(60)
    assert np.array_equal(b.get_dropshot_tiles()[0].ravel(), ref.dropshot_tiles()) and ref.dropshot_tiles().any()
    rc = ref.clone()
    b.clone_arena(0, 1)
    assert np.array_equal(b.get_dropshot_tiles()[1].ravel(), rc.dropshot_tiles())
    assert pu.struct_equal(b.get_ball_states()[1], rc.ball_state())
    rc.close(); ref.close(); b.close()


def test_hostcheck_clone(hostlib, oracle):
    require_reference_platform(oracle)
    _clone(hostlib, oracle)

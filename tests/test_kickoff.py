"""Direct parity of the device-side ``ResetToRandomKickoff`` (rsb_reset_to_random_kickoff, csrc/sim/rsb_kickoff.h) against
``Arena::ResetToRandomKickoff`` (reference src/Sim/Arena/Arena.cpp:113-223) for every game mode x team size x seed.

The reference deals the shuffled spawn slots to the cars in the iteration order of an ``unordered_set<Car*>``,
which is not a function of the seed; what IS defined is, per team, the multiset of (position, rotation, boost) handed out,
and the ball.  Those are compared bit for bit.
"""
import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import types as T
from rocketsim_b200.arena_batch import ArenaBatch

MODES = [("soccar", T.MODE_SOCCAR), ("hoops", T.MODE_HOOPS), ("heatseeker", T.MODE_HEATSEEKER), ("snowday", T.MODE_SNOWDAY), ("void", T.MODE_THE_VOID),
         ("dropshot", T.MODE_DROPSHOT)]
SEEDS = list(range(1, 17)) + [1234, 99991, 2 ** 31 - 1, -7, -123456, -2147483648]   # negative seeds: 64-bit sign extension in libstdc++


def _key(st):
    return (st["pos"].tobytes(), st["rotMat"].tobytes(), st["vel"].tobytes(), st["angVel"].tobytes(), np.asarray(st["boost"]).tobytes(),
            bytes([int(st["isOnGround"]), int(st["hasJumped"]), int(st["isDemoed"])]))


def _check_kickoffs(lib, oracle, mode, cpt):
    teams = [0] * cpt + [1] * cpt
    ref = pu.make_ref_arena(oracle, mode, teams)
    batch = ArenaBatch(mode, len(SEEDS), teams, lib=lib)
    # leave the kickoff pose first, so the reset has something to undo (boost used, pads taken, ball moving)
    c = np.zeros((len(SEEDS), len(teams)), T.CarControls); c["throttle"] = 1; c["boost"] = 1
    batch.set_controls(c); batch.step(30)
    batch.reset_to_random_kickoff(np.asarray(SEEDS, np.int32))
    cs, bs = batch.get_car_states(), batch.get_ball_states()
    ps = batch.get_pad_states() if batch.n_pads else None
    for k, seed in enumerate(SEEDS):
        for cid in ref.car_ids:
            ref.set_controls(cid, c[0, 0])
        ref.step(5)
        ref.reset_kickoff(seed)
        for team in (0, 1):
            mine = sorted(_key(cs[k, i]) for i, t in enumerate(teams) if t == team)
            want = sorted(_key(ref.car_state(cid)) for cid, t in zip(ref.car_ids, teams) if t == team)
            assert mine == want, f"mode {mode} {cpt}v{cpt} seed {seed} team {team}"
        rb = ref.ball_state()
        for f in ("pos", "vel", "angVel", "rotMat"):
            assert bs[k][f].tobytes() == rb[f].tobytes(), f"mode {mode} seed {seed} ball {f}"
        assert bs[k]["hsInfo"].tobytes() == rb["hsInfo"].tobytes()
        assert pu.struct_equal(bs[k]["dsInfo"], rb["dsInfo"])
        if ps is not None:
            rp = ref.pad_states()
            assert np.array_equal(ps[k]["isActive"], rp["isActive"]) and np.array_equal(ps[k]["cooldown"], rp["cooldown"])
    batch.close(); ref.close()


@pytest.mark.parametrize("cpt", [1, 2, 3])
@pytest.mark.parametrize("name,mode", MODES, ids=[m[0] for m in MODES])
def test_hostcheck_reset_to_random_kickoff_vs_oracle(hostlib, oracle, name, mode, cpt):
    _check_kickoffs(hostlib, oracle, mode, cpt)


@pytest.mark.gpu
@pytest.mark.parametrize("cpt", [1, 2, 3])
@pytest.mark.parametrize("name,mode", MODES, ids=[m[0] for m in MODES])
def test_cuda_reset_to_random_kickoff_vs_oracle(cudalib, oracle, name, mode, cpt):
    _check_kickoffs(cudalib, oracle, mode, cpt)

"""Ball-only batches (0 cars): the building block of a batched BallPredTracker (the reference's
src/Sim/BallPredTracker/BallPredTracker.cpp:25-71 steps a ball-only shadow arena).  Single-step parity against a reference
arena without cars, for a ball fired at ramps, walls, the ceiling and the goal."""
import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import types as T

LAUNCHES = [((0, 0, 300), (1800, 2600, 900)), ((3000, -2000, 150), (2500, 800, -300)), ((-500, 4000, 600), (-300, 2200, 1500)),
            ((0, 0, 93.15), (0, 0, 0)), ((1000, 1000, 1800), (600, -900, 2000))]


def run(lib, oracle, mode, ticks=300):
    ref = oracle.RefArena(mode)
    batch = pu.ArenaBatch(mode, len(LAUNCHES), [], lib=lib)
    worst = 0.0
    for k, (pos, vel) in enumerate(LAUNCHES):
        bs = ref.ball_state()
        bs["pos"]["x"], bs["pos"]["y"], bs["pos"]["z"] = pos
        bs["vel"]["x"], bs["vel"]["y"], bs["vel"]["z"] = vel
        bs["angVel"]["x"] = 2.0
        ref.set_ball_state(bs)
        for t in range(ticks):
            batch.set_ball_states([k], np.array([ref.ball_state()], T.BallState))
            batch.set_ball_internals([k], np.array([ref.ball_internal()], T.BallInternal))
            ref.step(1)
            batch.step(1)
            mine = batch.get_ball_states()[k]
            worst = max(worst, max(pu.compare_ball(mine, ref.ball_state()).values()))
            assert bool(batch.is_ball_scored()[k]) == bool(ref.is_ball_scored()), (k, t)
    assert int(batch.get_error_flags().max()) == 0
    batch.close(); ref.close()
    return worst


@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_SNOWDAY], ids=["soccar", "hoops", "snowday"])
def test_ball_only_hostcheck(hostlib, oracle, mode):
    assert run(hostlib, oracle, mode, ticks=200) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_HEATSEEKER, T.MODE_SNOWDAY], ids=["soccar", "hoops", "heatseeker", "snowday"])
def test_ball_only_cuda(cudalib, oracle, mode):
    assert run(cudalib, oracle, mode) <= 1e-4


# ---------------------------------------------------------------------------------------------------------------
# Batched BallPredTracker (rsb_predict_ball) against the reference's own BallPredTracker on the same arena
def run_predict(lib, oracle, mode, n_ticks=240):
    teams = [0, 1]
    ref = pu.make_ref_arena(oracle, mode, teams)
    ref.seed_rng(4); ref.reset_kickoff(4)
    batch = pu.ArenaBatch(mode, len(LAUNCHES), teams, lib=lib)
    worst = 0.0
    for k, (pos, vel) in enumerate(LAUNCHES):
        bs = ref.ball_state()
        bs["pos"]["x"], bs["pos"]["y"], bs["pos"]["z"] = pos
        bs["vel"]["x"], bs["vel"]["y"], bs["vel"]["z"] = vel
        ref.set_ball_state(bs)
        pu.sync_from_ref(batch, ref, k)
        expect = ref.ball_pred(n_ticks)
        got = batch.predict_ball(n_ticks)[:, k]
        assert len(expect) == n_ticks
        assert pu.struct_equal(got[0]["pos"], expect[0]["pos"]) and pu.struct_equal(got[0]["vel"], expect[0]["vel"])  # row 0 = now, exactly
        for t in range(n_ticks):
            worst = max(worst, max(pu.compare_ball(got[t], expect[t]).values()))
            # whatever the spin does after a few bounces, the predicted PATH must agree (uu)
            assert np.abs(pu.vec3(got[t]["pos"]) - pu.vec3(expect[t]["pos"])).max() < 0.5, (k, t)
    # the prediction does not disturb the arenas themselves
    assert np.all(batch.get_tick_counts() == 0)
    batch.close(); ref.close()
    return worst


@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS], ids=["soccar", "hoops"])
def test_predict_ball_hostcheck(hostlib, oracle, mode):
    assert run_predict(hostlib, oracle, mode, n_ticks=180) <= 1e-4  # a free 180-tick ball rollout stays within the single-step bar


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_SNOWDAY], ids=["soccar", "hoops", "snowday"])
def test_predict_ball_cuda(cudalib, oracle, mode):
    worst = run_predict(cudalib, oracle, mode)
    # free 240-tick rollouts: the sphere stays within 1e-3 on every field; the Snowday puck (a 40-point hull bouncing on edges)
    # amplifies ULP differences in its spin after several bounces, so only its path is gated (inside run_predict)
    if mode != T.MODE_SNOWDAY:
        assert worst <= 1e-3

"""Device-side GameEventTracker and Arena::IsBallProbablyGoingIn against the reference's own
implementation (src/Sim/GameEventTracker/GameEventTracker.cpp, src/Sim/Arena/Arena.cpp:771-891): identical shot / goal-credit /
save events (discrete: bit-exact) on scripted plays, and identical going-in predictions on random ball states."""
import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import types as T
from rocketsim_b200.arena_batch import ArenaBatch


def place_car(ref, cid, x, y, yaw, speed=0.0, z=17.0):
    st = ref.car_state(cid)
    st["pos"]["x"], st["pos"]["y"], st["pos"]["z"] = x, y, z
    cy, sy = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
    st["rotMat"]["forward"] = (cy, sy, 0, 0); st["rotMat"]["right"] = (-sy, cy, 0, 0); st["rotMat"]["up"] = (0, 0, 1, 0)
    st["vel"]["x"], st["vel"]["y"], st["vel"]["z"] = speed * cy, speed * sy, 0
    st["angVel"]["x"] = st["angVel"]["y"] = st["angVel"]["z"] = 0
    st["boost"] = 100
    ref.set_car_state(cid, st)


def place_ball(ref, pos, vel=(0, 0, 0)):
    bs = ref.ball_state()
    bs["pos"]["x"], bs["pos"]["y"], bs["pos"]["z"] = pos
    bs["vel"]["x"], bs["vel"]["y"], bs["vel"]["z"] = vel
    bs["angVel"]["x"] = bs["angVel"]["y"] = bs["angVel"]["z"] = 0
    ref.set_ball_state(bs)


# scripted plays (Soccar 2v2; cars 1,2 blue attack +y; cars 3,4 orange defend +y goal)
def play_shot_and_goal(ref):
    ids = ref.car_ids
    place_ball(ref, (0, 2600, 93.15))
    place_car(ref, ids[0], 0, 1700, np.pi / 2, 1400)       # blue shooter straight behind the ball
    place_car(ref, ids[1], -2500, -3000, np.pi / 2)
    place_car(ref, ids[2], 3000, -4000, -np.pi / 2)
    place_car(ref, ids[3], -3000, -4000, -np.pi / 2)


def play_pass_then_shot(ref):
    ids = ref.car_ids
    place_ball(ref, (-900, 1500, 93.15))
    place_car(ref, ids[1], -1900, 1500, 0.0, 1500)           # blue passer pushes the ball across (+x)
    place_car(ref, ids[0], 300, 300, np.pi / 2 - 0.15, 1900)  # blue shooter arrives on the far side
    place_car(ref, ids[2], 3000, -4000, -np.pi / 2)
    place_car(ref, ids[3], -3000, -4000, -np.pi / 2)


def play_save(ref):
    """A rolling shot credited to blue car 1 (its ball touch is scripted through CarState.ballHitInfo) that the orange keeper,
    parked across the goal line, blocks."""
    ids = ref.car_ids
    place_ball(ref, (0, 1500, 93.15), (0, 2600, 0))
    place_car(ref, ids[0], -600, 900, np.pi / 2)
    st = ref.car_state(ids[0])
    st["ballHitInfo"]["isValid"] = 1
    st["ballHitInfo"]["tickCountWhenHit"] = 0
    ref.set_car_state(ids[0], st)
    place_car(ref, ids[1], -2500, -3000, np.pi / 2)
    place_car(ref, ids[2], 0, 4850, 0.0)                      # orange keeper, side-on in the goal mouth
    place_car(ref, ids[3], -3000, -4000, -np.pi / 2)


PLAYS = [("shot_goal", play_shot_and_goal, 260, 1.0), ("pass_shot", play_pass_then_shot, 300, 1.0), ("save", play_save, 260, 0.0)]


def run_play(lib, oracle, setup, ticks, throttle, mode=T.MODE_SOCCAR):
    teams = [0, 0, 1, 1]
    ref = pu.make_ref_arena(oracle, mode, teams)
    ref.enable_event_tracker()
    ref.seed_rng(1); ref.reset_kickoff(1)
    setup(ref)
    batch = ArenaBatch(mode, 2, teams, lib=lib)
    batch.set_event_tracker(True)
    ctrl = np.zeros(4, T.CarControls)
    ctrl["throttle"] = throttle; ctrl["boost"] = throttle > 0
    ref_ev, my_ev = [], []
    for t in range(ticks):
        for a in range(2):
            pu.sync_from_ref(batch, ref, a)
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, ctrl[i])
        batch.set_controls(np.broadcast_to(ctrl, (2, 4)))
        ref.step(1); batch.step(1)
        gi_ref = ref.is_ball_probably_going_in(2.0, 0.0)
        gi_mine, team_mine = batch.is_ball_probably_going_in(2.0, 0.0)
        assert bool(gi_mine[1]) == gi_ref[0], f"tick {t}: going-in mismatch"
        if gi_ref[0]:
            assert int(team_mine[1]) == gi_ref[1]
        for e in ref.drain_events():
            if int(e["type"]) >= T.EVENT_SHOT:
                ref_ev.append((int(e["tick"]), int(e["type"]), int(e["a"]), int(e["b"])))
        for e in batch.get_events():
            if int(e["arena"]) == 1 and int(e["type"]) >= T.EVENT_SHOT:
                my_ev.append((int(e["tick"]), int(e["type"]), int(e["a"]), int(e["b"])))
        if ref.is_ball_scored():
            break
    batch.close(); ref.close()
    return ref_ev, my_ev


def check_plays(lib, oracle):
    seen = set()
    for name, setup, ticks, throttle in PLAYS:
        ref_ev, my_ev = run_play(lib, oracle, setup, ticks, throttle)
        assert my_ev == ref_ev, f"{name}: events differ\n ref {ref_ev}\n mine {my_ev}"
        seen |= {e[1] for e in ref_ev}
    # the scripted plays must actually exercise every tracker event
    assert {T.EVENT_SHOT, T.EVENT_GOAL_CREDIT, T.EVENT_SAVE} <= seen, f"plays produced only {seen}"


def check_going_in_random(lib, oracle, mode, n=400):
    rng = np.random.default_rng(5)
    ref = oracle.RefArena(mode)
    batch = ArenaBatch(mode, n, [], lib=lib)
    ex, ey, h = (2966.0, 3581.0, 1820.0) if mode == T.MODE_HOOPS else (4096.0, 5120.0, 2044.0)
    states = np.zeros(n, T.BallState)
    exp_in, exp_team = np.zeros(n, bool), np.zeros(n, np.uint8)
    for k in range(n):
        bs = ref.ball_state()
        bs["pos"]["x"], bs["pos"]["y"], bs["pos"]["z"] = rng.uniform(-ex, ex) * 0.5, rng.uniform(-ey, ey), rng.uniform(100, h * 0.8)
        bs["vel"]["x"], bs["vel"]["y"], bs["vel"]["z"] = rng.uniform(-800, 800), rng.uniform(-4000, 4000), rng.uniform(-1500, 1500)
        ref.set_ball_state(bs)
        states[k] = ref.ball_state()
        r, t = ref.is_ball_probably_going_in(2.0, 10.0)
        exp_in[k], exp_team[k] = r, t
    batch.set_ball_states(np.arange(n), states)
    got, team = batch.is_ball_probably_going_in(2.0, 10.0)
    assert np.array_equal(got, exp_in)
    assert np.array_equal(team[exp_in], exp_team[exp_in])
    assert 0.05 < exp_in.mean() < 0.95  # both outcomes covered
    batch.close(); ref.close()


def test_event_tracker_hostcheck(hostlib, oracle):
    check_plays(hostlib, oracle)


@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_SNOWDAY], ids=["soccar", "hoops", "snowday"])
def test_going_in_hostcheck(hostlib, oracle, mode):
    check_going_in_random(hostlib, oracle, mode)


def test_tracker_rejects_unsupported_modes(hostlib):
    from rocketsim_b200 import lib as rlib
    b = ArenaBatch(T.MODE_THE_VOID, 1, [0, 1], lib=hostlib)
    with pytest.raises(rlib.RsbError):
        b.set_event_tracker(True)
    with pytest.raises(rlib.RsbError):
        b.is_ball_probably_going_in()
    b.close()


@pytest.mark.gpu
def test_event_tracker_cuda(cudalib, oracle):
    check_plays(cudalib, oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [T.MODE_SOCCAR, T.MODE_HOOPS, T.MODE_SNOWDAY], ids=["soccar", "hoops", "snowday"])
def test_going_in_cuda(cudalib, oracle, mode):
    check_going_in_random(cudalib, oracle, mode)

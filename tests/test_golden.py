"""Single-step parity against the committed golden records (tests/golden/*.npz, generated from
the unmodified reference by tests/golden/make_golden.py).

Bar (BASELINE.json north_star): discrete state bit-exact per single step; float state within
1e-4 relative, `max|a-b| / max(||b||, eps_field)` with eps pos 1 uu, vel 1 uu/s, angVel 0.01 rad/s,
rotMat 1.
"""
import glob
import os

import numpy as np
import pytest

import parity_util as pu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
NAMES = [os.path.basename(p)[:-4] for p in GOLDEN]

# float tolerance (relative, see parity_util.rel_err), every field, every well-conditioned tick
TOL = 1e-4
# Ticks in which a suspension ray rests on another car / the ball ("wheel_on_dynamic" in the records) go through
# btSubsimplexConvexCast.  Round 1 gave them a 2e-2 carve-out; the cast is now bit-exact against Bullet's
# (tests/test_native_units.py), so they are held to the same bar and only reported separately.
TOL_ILLCOND = TOL


def load(path):
    with np.load(path) as z:
        return pu.upgrade_record({k: z[k] for k in z.files})


def check(res, name):
    assert res["error_flags"] == 0, f"{name}: device error flags {res['error_flags']}"
    assert not res["discrete_mismatches"], f"{name}: discrete mismatches {res['discrete_mismatches'][:5]}"
    for k, v in res["worst"].items():
        assert v <= TOL, f"{name}: {k} rel err {v:.3g} > {TOL}"
    for k, v in res["worst_illcond"].items():
        assert v <= TOL_ILLCOND, f"{name}: {k} rel err {v:.3g} > {TOL_ILLCOND} (ill-conditioned tick)"


def test_golden_files_present():
    assert len(GOLDEN) >= 17


def test_golden_records_cover_demos_goals_supersonic():
    """The records must actually contain the events they are meant to pin (no committed record held a demo)."""
    tot = dict(demo_events=0, bump_events=0, goal_events=0, demoed=0, supersonic=0, scored=0, respawns=0, demo_ticks_with_rot_change=0)
    for path in GOLDEN:
        rec = load(path)
        ev = rec["events"]
        if len(ev):
            tot["goal_events"] += int((ev[:, 1] == 1).sum()); tot["bump_events"] += int((ev[:, 1] == 2).sum()); tot["demo_events"] += int((ev[:, 1] == 3).sum())
        pre, post = rec["car_pre"], rec["car_post"]
        tot["demoed"] += int(post["isDemoed"].sum()); tot["supersonic"] += int(post["isSupersonic"].sum()); tot["scored"] += int(rec["scored"].sum())
        tot["respawns"] += int(((pre["isDemoed"] != 0) & (post["isDemoed"] == 0)).sum())   # the record spans the respawn tick (Car.cpp:69-87)
        # the tick of the demo itself: the victim's body still integrates, CarState.rotMat must stay at its pre-tick value (Car.cpp:133-138)
        hit = (pre["isDemoed"] == 0) & (post["isDemoed"] != 0)
        for t, i in zip(*np.nonzero(hit)):
            assert pre["rotMat"][t, i].tobytes() == post["rotMat"][t, i].tobytes()
            tot["demo_ticks_with_rot_change"] += 1
        if "demo" in os.path.basename(path):
            assert post["isDemoed"].any() and post["isSupersonic"].any(), path
    assert tot["demo_events"] >= 8 and tot["bump_events"] >= 10 and tot["goal_events"] >= 2, tot
    assert tot["demoed"] >= 360 * 8 and tot["respawns"] >= 8 and tot["supersonic"] >= 500 and tot["scored"] >= 50, tot
    assert tot["demo_ticks_with_rot_change"] >= 8, tot


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_oracle_reproduces_golden(oracle, path):
    """Pins the fixtures: the oracle built here regenerates the committed records bit for bit."""
    rec = load(path)
    teams = [int(x) for x in rec["teams"]]
    ref = pu.make_ref_arena(oracle, int(rec["mode"]), teams)
    T_ = 60
    for t in range(T_):
        for i, cid in enumerate(ref.car_ids):
            ref.set_car_state(cid, rec["car_pre"][t, i])
            ref.set_car_internal(cid, rec["carint_pre"][t, i])
            ref.set_controls(cid, rec["controls"][t, i])
        ref.set_ball_state(rec["ball_pre"][t])
        ref.set_ball_internal(rec["ballint_pre"][t])
        ref.set_pad_states(rec["pads_pre"][t])
        ref.step(1)
        for i, cid in enumerate(ref.car_ids):
            got = ref.car_state(cid)
            for f in ("pos", "vel", "angVel", "rotMat", "boost"):
                assert got[f].tobytes() == rec["car_post"][t, i][f].tobytes(), (t, i, f)
        assert ref.ball_state()["pos"].tobytes() == rec["ball_post"][t]["pos"].tobytes()
    ref.close()


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_hostcheck_single_step_vs_golden(hostlib, path):
    """CPU tier: the device code compiled for the host (test build) against the golden records."""
    rec = load(path)
    res = pu.replay_record(rec, hostlib, n_arenas=1)
    check(res, os.path.basename(path))
    ev_ref = [tuple(int(v) for v in e) for e in rec["events"]]
    assert sorted(res["events"]) == sorted(ev_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_cuda_single_step_vs_golden(cudalib, path):
    """GPU tier: the CUDA kernels through the C ABI; 33 arenas replay the record (first and last checked)."""
    rec = load(path)
    res = pu.replay_record(rec, cudalib, n_arenas=33)
    check(res, os.path.basename(path))
    ev_ref = [tuple(int(v) for v in e) for e in rec["events"]]
    assert sorted(res["events"]) == sorted(ev_ref)

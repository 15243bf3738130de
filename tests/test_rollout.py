"""Live-oracle parity on the GPU box (oracle/_ref travels with the snapshot; nothing here reads /root/reference).

1. Single-step parity along a README-bench rollout for every configuration of BASELINE.json (mode x team size):
   the CUDA batch is re-synchronised from the reference every tick, so every tick is an independent single-step case
   with seeded inputs (discrete state bit-exact, floats within 1e-4 relative; north_star).
2. Free-running rollouts: divergence from the reference REPORTED at 1 / 10 / 120 ticks (the reference itself notes that
   tiny errors accumulate, README.md:13-21); only the first tick is gated at the single-step tolerance, later
   checkpoints carry loose sanity bounds.
"""
import json
import os

import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import types as T

TOL = 1e-4          # single-step float tolerance (relative, parity_util.rel_err)
TOL_ILLCOND = 2e-2  # ticks with a suspension ray resting on a dynamic body

# BASELINE.json configs: VOID 1v1, Soccar 2v2, Soccar 3v3 (demos/bumps), Hoops 2v2, Heatseeker 1v1, Snowday 3v3
CONFIGS = [("void_1v1", T.MODE_THE_VOID, 1), ("soccar_2v2", T.MODE_SOCCAR, 2), ("soccar_3v3", T.MODE_SOCCAR, 3),
           ("hoops_2v2", T.MODE_HOOPS, 2), ("heatseeker_1v1", T.MODE_HEATSEEKER, 1), ("snowday_3v3", T.MODE_SNOWDAY, 3)]
IDS = [c[0] for c in CONFIGS]


def _wheel_on_dynamic(ref):
    for cid in ref.car_ids:
        wd = ref.wheel_debug(cid)
        if np.any((wd[:, 0] != 0) & (wd[:, 1] == 0)):
            return True
    return False


def _resync_rollout(lib, oracle, mode, cpt, seed, ticks, n_arenas=5):
    teams = [0] * cpt + [1] * cpt
    ref = pu.make_ref_arena(oracle, mode, teams)
    ref.seed_rng(seed)
    ref.reset_kickoff(seed)
    pu.canonical_kickoff(ref, teams)
    batch = pu.ArenaBatch(mode, n_arenas, teams, lib=lib)
    for a in range(n_arenas):
        pu.sync_from_ref(batch, ref, a)
        batch.seed_rng(a, seed)
    sched = pu.ControlSchedule(len(teams), seed)
    worst, worst_ill, mism, n_ill = {}, {}, [], 0
    for t in range(ticks):
        c = sched.at(t)
        if t < 150:  # drive at the ball first so that touches, bumps and pad pickups happen inside the window
            c["throttle"] = 1; c["boost"] = 1; c["steer"] = 0; c["jump"] = 0; c["handbrake"] = 0
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, c[i])
        batch.set_controls(np.broadcast_to(c, (n_arenas, len(teams))))
        ref.step(1)
        batch.step(1)
        ill = _wheel_on_dynamic(ref)
        n_ill += ill
        wd = worst_ill if ill else worst
        cs, bs = batch.get_car_states(), batch.get_ball_states()
        ps = batch.get_pad_states() if batch.n_pads else None
        for a in (0, n_arenas - 1):
            for i, cid in enumerate(ref.car_ids):
                bad, errs = pu.compare_car(cs[a, i], ref.car_state(cid))
                if bad:
                    mism.append((t, a, i, bad))
                for k, v in errs.items():
                    wd[k] = max(wd.get(k, 0.0), v)
            for k, v in pu.compare_ball(bs[a], ref.ball_state()).items():
                wd["ball_" + k] = max(wd.get("ball_" + k, 0.0), v)
            if ps is not None:
                rp = ref.pad_states()
                if not (np.array_equal(ps[a]["isActive"], rp["isActive"]) and np.array_equal(ps[a]["prevLockedCarID"], rp["prevLockedCarID"])):
                    mism.append((t, a, -1, ["pads"]))
        if bool(batch.is_ball_scored()[0]) != bool(ref.is_ball_scored()):
            mism.append((t, 0, -1, ["scored"]))
        if ref.is_ball_scored():
            ref.seed_rng(seed + t)
            ref.reset_kickoff(seed + t)
            pu.canonical_kickoff(ref, teams)
        for a in range(n_arenas):
            pu.sync_from_ref(batch, ref, a)
            batch.seed_rng(a, seed + t)
        ref.seed_rng(seed + t)
    err = int(batch.get_error_flags().max())
    batch.close(); ref.close()
    return worst, worst_ill, mism, n_ill, err


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode,cpt", CONFIGS, ids=IDS)
def test_cuda_single_step_vs_live_oracle(cudalib, oracle, name, mode, cpt):
    ticks = 360
    worst, worst_ill, mism, n_ill, err = _resync_rollout(cudalib, oracle, mode, cpt, seed=101 + cpt, ticks=ticks)
    assert err == 0, f"{name}: device error flags {err}"
    assert not mism, f"{name}: discrete mismatches {mism[:5]}"
    for k, v in worst.items():
        assert v <= TOL, f"{name}: {k} rel err {v:.3g} > {TOL}"
    for k, v in worst_ill.items():
        assert v <= TOL_ILLCOND, f"{name}: {k} rel err {v:.3g} > {TOL_ILLCOND} (ill-conditioned tick)"
    # The scripted drive-at-the-ball opening piles cars up, so up to 10 % of ill-conditioned ticks is normal here.  The bound only
    # guards against the looser tolerance becoming the rule.  (Kickoffs are re-dealt canonically: parity_util.canonical_kickoff.)
    assert n_ill <= 0.10 * ticks


def _free_rollout(lib, oracle, mode, cpt, seed, checkpoints=(1, 10, 120)):
    teams = [0] * cpt + [1] * cpt
    ref = pu.make_ref_arena(oracle, mode, teams)
    ref.seed_rng(seed)
    ref.reset_kickoff(seed)
    pu.canonical_kickoff(ref, teams)
    batch = pu.ArenaBatch(mode, 3, teams, lib=lib)
    for a in range(3):
        pu.sync_from_ref(batch, ref, a)
        batch.seed_rng(a, seed)
    sched = pu.ControlSchedule(len(teams), seed)
    out = {}
    for t in range(max(checkpoints)):
        c = sched.at(t)
        if t < 60:
            c["throttle"] = 1; c["boost"] = 1; c["steer"] = 0; c["jump"] = 0; c["handbrake"] = 0
        for i, cid in enumerate(ref.car_ids):
            ref.set_controls(cid, c[i])
        batch.set_controls(np.broadcast_to(c, (3, len(teams))))
        ref.step(1)
        batch.step(1)
        if t + 1 in checkpoints:
            cs, bs = batch.get_car_states(), batch.get_ball_states()
            d = []
            for i, cid in enumerate(ref.car_ids):
                d.append(pu.vec3(cs[0, i]["pos"]) - pu.vec3(ref.car_state(cid)["pos"]))
            d.append(pu.vec3(bs[0]["pos"]) - pu.vec3(ref.ball_state()["pos"]))
            d = np.array(d)
            flags_equal = all(not pu.compare_car(cs[0, i], ref.car_state(cid))[0] for i, cid in enumerate(ref.car_ids))
            out[t + 1] = dict(pos_rms_uu=float(np.sqrt(np.mean(np.sum(d * d, axis=1)))), pos_max_uu=float(np.abs(d).max()),
                              discrete_equal=bool(flags_equal))
    batch.close(); ref.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode,cpt", CONFIGS, ids=IDS)
def test_cuda_rollout_divergence_report(cudalib, oracle, name, mode, cpt):
    """Divergence of a free-running CUDA rollout from the reference at 1/10/120 ticks (positions, uu)."""
    rep = _free_rollout(cudalib, oracle, mode, cpt, seed=31 + cpt)
    print(f"\nrollout divergence {name}: " + json.dumps(rep))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "rollout_divergence.jsonl"), "a") as f:
            f.write(json.dumps({"config": name, "divergence": rep}) + "\n")
    except OSError:
        pass
    assert rep[1]["pos_max_uu"] <= 1e-4 * 6000 and rep[1]["discrete_equal"]  # one tick: the single-step bar
    assert rep[10]["pos_rms_uu"] < 1.0      # sanity bounds only: chaotic contacts amplify ULP differences
    assert rep[120]["pos_rms_uu"] < 400.0

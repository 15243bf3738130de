"""Developer tool: step the oracle and the host-compiled check build side by side and print
where they diverge.  usage: dev_parity.py MODE CARS_PER_TEAM TICKS [--sync] [--seed S]"""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests")); import parity_util as pu
from rocketsim_b200 import types as T, lib as rsb_lib, synthetic_meshes as sm
from oracle import refsim

ap = argparse.ArgumentParser()
ap.add_argument("mode", type=int); ap.add_argument("cpt", type=int); ap.add_argument("ticks", type=int)
ap.add_argument("--sync", action="store_true"); ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--planes", action="store_true"); ap.add_argument("--verbose", action="store_true")
ap.add_argument("--stop", type=float, default=1e-3)
ap.add_argument("--drive", type=int, default=0)
ap.add_argument("--aim", action="store_true", help="place every car on a circle around the ball, facing it")
a = ap.parse_args()

L = rsb_lib.load(pu.build_hostcheck())
ms = sm.mesh_sets(planes_only=a.planes)
refsim.init_from_mem(ms); rsb_lib.init_from_mem(ms, L)
teams = [0] * a.cpt + [1] * a.cpt
ref = refsim.RefArena(a.mode)
for t in teams: ref.add_car(t)
b = pu.ArenaBatch(a.mode, 1, teams, lib=L)
ref.seed_rng(a.seed); b.seed_rng(0, a.seed)
ref.reset_kickoff(a.seed)
if a.aim:
    n = len(teams)
    for i, cid in enumerate(ref.car_ids):
        st = ref.car_state(cid)
        ang = 2 * np.pi * i / n + 0.3
        r = 900.0 + 150 * i
        st['pos']['x'] = r * np.cos(ang); st['pos']['y'] = r * np.sin(ang); st['pos']['z'] = 17.0
        yaw = ang + np.pi
        cy, sy = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
        st['rotMat']['forward'] = (cy, sy, 0, 0); st['rotMat']['right'] = (-sy, cy, 0, 0); st['rotMat']['up'] = (0, 0, 1, 0)
        ref.set_car_state(cid, st)
pu.sync_from_ref(b, ref)
sched = pu.ControlSchedule(len(teams), a.seed)
worst = {}
for tick in range(a.ticks):
    c = sched.at(tick)
    if tick < a.drive:
        c['throttle'] = 1; c['boost'] = 1; c['steer'] = 0; c['jump'] = 0; c['handbrake'] = 0; c['pitch'] = 0; c['yaw'] = 0; c['roll'] = 0
    for i, cid in enumerate(ref.car_ids): ref.set_controls(cid, c[i])
    b.set_controls(c[None])
    ref.step(1); b.step(1)
    mine_c = b.get_car_states()[0]; mine_b = b.get_ball_states()[0]
    bad_any = False
    for i, cid in enumerate(ref.car_ids):
        rc = ref.car_state(cid)
        bad, errs = pu.compare_car(mine_c[i], rc)
        for k, v in errs.items(): worst[k] = max(worst.get(k, 0), v)
        m = max(errs.values())
        if bad or m > a.stop:
            bad_any = True
            print(f"tick {tick} car {i}: discrete mismatch {bad} max err {m:.3g}", {k: f'{v:.2g}' for k, v in errs.items() if v > a.stop})
            if a.verbose:
                for f in ("pos", "vel", "angVel"): print("  ", f, pu.field_as_array(mine_c[i], f), pu.field_as_array(rc, f))
                for f in bad: print("  ", f, mine_c[i][f], rc[f])
    rb = ref.ball_state(); eb = pu.compare_ball(mine_b, rb)
    for k, v in eb.items(): worst["ball_" + k] = max(worst.get("ball_" + k, 0), v)
    if max(eb.values()) > a.stop:
        bad_any = True
        print(f"tick {tick} ball: ", {k: f'{v:.2g}' for k, v in eb.items()}, pu.field_as_array(mine_b, 'pos'), pu.field_as_array(rb, 'pos'))
    if b.n_pads:
        mp = b.get_pad_states()[0]; rp = ref.pad_states()
        if not (np.array_equal(mp['isActive'], rp['isActive']) and np.array_equal(mp['prevLockedCarID'], rp['prevLockedCarID'])):
            bad_any = True; print(f"tick {tick} pads mismatch")
    ef = b.get_error_flags()[0]
    if ef: print(f"tick {tick} error flags {ef}")
    if bad_any and not a.sync: print("diverged; stopping"); break
    if a.sync: pu.sync_from_ref(b, ref)
print("worst rel errs:", {k: f"{v:.2g}" for k, v in worst.items()})
print("ref events:", ref.drain_events()[:10]); print("my events:", b.get_events()[:10])

"""Developer tool (CPU): replay ONE arena of a soak run (global arena index G of `scripts/soak.py`, same seeds, same input stream,
same auto-reset rule) on the host check build, find the tick at which a car leaves the arena shell, then hand the state some
ticks earlier to the reference and step both through the event, comparing every tick bit for bit.
usage: dev_escape_replay.py G TICKS [MODE] [CARS_PER_TEAM]"""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import parity_util as pu
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm, types as T
from oracle import refsim
G, TICKS = int(sys.argv[1]), int(sys.argv[2]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0; cpt = int(sys.argv[4]) if len(sys.argv) > 4 else 2
L = rlib.load(pu.build_hostcheck()); ms = sm.mesh_sets(); rlib.init_from_mem(ms, L); refsim.init_from_mem(ms)
teams = [0] * cpt + [1] * cpt; n = len(teams)
xmax, ymax = (5700, 5000) if mode == 4 else (4400, 6200)

def make():
    b = pu.ArenaBatch(mode, 1, teams, seed=77 + G, lib=L)
    b.set_arena_offset(G)
    b.reset_to_random_kickoff(np.asarray([77 + G], np.int32)); b.set_auto_reset(True, 77)
    if mode in (0, 1, 3): b.set_event_tracker(True)
    return b

gen = pu.ArenaBatch(T.MODE_THE_VOID, G + 1, teams, lib=L)  # only used to draw the input stream of global arena G (the stream is keyed by the arena index inside the batch)

def controls(b, t0, nt):
    seq = np.zeros((nt, G + 1, n), T.CarControls)
    rlib.check(L, L.rsb_bench_fill_controls_device(gen.h, C.c_void_p(seq.ctypes.data), nt, t0, 77))
    return np.ascontiguousarray(seq[:, G:G + 1])

b = make(); CH = 600; first = None
for c in range(TICKS // CH):
    seq = controls(b, c * CH, CH)
    for t in range(CH):
        b.step_with_controls(seq[t:t + 1])
        p = b.get_car_states()[0]["pos"]
        if (np.abs(p["x"]) > xmax).any() or (np.abs(p["y"]) > ymax).any():
            first = c * CH + t; break
    if first is not None: break
print("first tick with a car outside:", first)
if first is None: sys.exit(0)
back = 400; t_sync = max(0, first - back)
b.close(); b = make()
for c in range(0, t_sync, CH):
    nt = min(CH, t_sync - c)
    if nt > 0: b.step_with_controls(controls(b, c, nt))
# hand the state to the reference: car states + hidden state + ball + pads (the reference arena is fresh otherwise)
ref = pu.make_ref_arena(refsim, mode, teams)
cs, ci = b.get_car_states()[0], b.get_car_internals()[0]
for i, cid in enumerate(ref.car_ids): ref.set_car_state(cid, cs[i]); ref.set_car_internal(cid, ci[i])
ref.set_ball_state(b.get_ball_states()[0]); ref.set_ball_internal(b.get_ball_internals()[0])
if b.n_pads: ref.set_pad_states(b.get_pad_states()[0])
pu.sync_from_ref(b, ref, 0)  # both sides now hold exactly the reference's view of that state
b.set_auto_reset(False, 0)
seq = controls(b, t_sync, back + 200)
worst = 0.0; nbad = 0
for t in range(back + 200):
    for i, cid in enumerate(ref.car_ids): ref.set_controls(cid, seq[t, 0, i])
    b.set_controls(seq[t]); ref.step(1); b.step(1)
    mine = b.get_car_states()[0]
    for i, cid in enumerate(ref.car_ids):
        rc = ref.car_state(cid)
        if not pu.struct_equal(mine[i], rc):
            bad, errs = pu.compare_car(mine[i], rc); nbad += 1; worst = max(worst, max(errs.values()))
            if nbad < 5: print("tick", t_sync + t, "car", i, "differs", bad, {k: v for k, v in errs.items() if v})
    if ref.is_ball_scored(): ref.reset_kickoff(5); pu.canonical_kickoff(ref, teams); pu.sync_from_ref(b, ref, 0)
p = np.array([[float(ref.car_state(cid)["pos"][k]) for k in "xyz"] for cid in ref.car_ids])
print("reference car positions at the end:", p.round(1).tolist())
print("outside in the reference too:", bool((np.abs(p[:, 0]) > xmax).any() or (np.abs(p[:, 1]) > ymax).any()), "| differing car-ticks:", nbad, "worst rel err", worst)

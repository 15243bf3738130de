"""Soak run (GPU): many arenas for many ticks with README-bench inputs and kickoff reset on goal; reports device error flags,
non-finite state, goals and throughput.  usage: python scripts/soak.py [arenas] [ticks] [mode] [cars_per_team]"""
import os, sys, json, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm, types as T
from rocketsim_b200.arena_batch import ArenaBatch
A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TICKS = int(sys.argv[2]) if len(sys.argv) > 2 else 36000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cpt = int(sys.argv[4]) if len(sys.argv) > 4 else 2
L = rlib.load(); rlib.init_from_mem(sm.mesh_sets(), L)
b = ArenaBatch(mode, A, [0] * cpt + [1] * cpt, seed=77, lib=L)
b.reset_to_random_kickoff((77 + np.arange(A)).astype(np.int32)); b.set_auto_reset(True, 77)
if mode in (0, 1, 3): b.set_event_tracker(True)
CH = 600
seq = torch.empty(CH * A * 2 * cpt * 24, dtype=torch.uint8, device="cuda")
counts = {}
t0 = time.perf_counter()
for c in range(TICKS // CH):
    rlib.check(L, L.rsb_bench_fill_controls_device(b.h, C.c_void_p(seq.data_ptr()), CH, c * CH, 77))
    rlib.check(L, L.rsb_step_with_controls_device(b.h, C.c_void_p(seq.data_ptr()), CH))
    ev = b.get_events(cap=1 << 20)
    for k, n in zip(*np.unique(ev["type"], return_counts=True)): counts[int(k)] = counts.get(int(k), 0) + int(n)
b.sync(); dt = time.perf_counter() - t0
cs, bs = b.get_car_states(), b.get_ball_states()
finite = bool(np.isfinite(cs["pos"]["x"]).all() and np.isfinite(cs["vel"]["x"]).all() and np.isfinite(bs["pos"]["z"]).all() and np.isfinite(cs["rotMat"]["up"]["z"]).all())
XMAX, YMAX = (5700, 5000) if mode == 4 else (4400, 6200)  # Dropshot: the hexagonal shell of synthetic_meshes.dropshot_meshes
inside = bool(np.abs(cs["pos"]["x"]).max() < XMAX and np.abs(cs["pos"]["y"]).max() < YMAX and cs["pos"]["z"].max() < 2200 and cs["pos"]["z"].min() > -50)
err = b.get_error_flags()
names = {1: "goal", 2: "bump", 3: "demo", 4: "shot", 5: "goal_credit", 6: "save"}
print(json.dumps({"mode": mode, "cars_per_team": cpt, "arenas": A, "ticks": int(b.get_tick_counts().min()), "arena_ticks_per_s": A * (TICKS // CH) * CH / dt,
                  "error_flags_or": int(np.bitwise_or.reduce(err)), "arenas_with_error": int((err != 0).sum()), "all_finite": finite, "inside_arena": inside,
                  "events": {names.get(k, str(k)): v for k, v in sorted(counts.items())}}))
if not inside:
    p = cs["pos"]; k = np.unravel_index(np.argmax(np.abs(p["y"]) + 1.5 * np.abs(p["x"]) + np.abs(p["z"] - 1000)), p.shape)
    print("outlier car", k, [float(p[k][a]) for a in "xyz"], "demoed", int(cs[k]["isDemoed"]), "max|x|", float(np.abs(p["x"]).max()), "max|y|", float(np.abs(p["y"]).max()), "z range", float(p["z"].min()), float(p["z"].max()))

"""Developer tool (CPU, oracle only): do cars ever leave the SYNTHETIC arena shells under the README-bench inputs in the reference
itself?  (The soak runs flag such cars; parity with the reference is bit-exact, so an escape is a property of the stand-in meshes and
the reference's own contact handling, not of the batched tick.)  usage: dev_escape_probe.py MODE ARENAS TICKS"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R)
import numpy as np
from rocketsim_b200 import synthetic_meshes as sm
from oracle import refsim
mode, n, ticks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
refsim.init_from_mem(sm.mesh_sets())
xmax, ymax = (5700, 5000) if mode == 4 else (4400, 6200)
out = 0
for g in range(n):
    ref = refsim.RefArena(mode)
    for t in (0, 0, 1, 1): ref.add_car(t)
    ref.seed_rng(77 + g); ref.reset_kickoff(77 + g)
    ref.run_bench_inputs(ticks, 77, g)
    for cid in ref.car_ids:
        p = ref.car_state(cid)["pos"]
        if abs(float(p["x"])) > xmax or abs(float(p["y"])) > ymax or float(p["z"]) < -50 or float(p["z"]) > 2200:
            out += 1; print("arena", g, "car", cid, [round(float(p[k]), 1) for k in "xyz"], flush=True)
    ref.close()
print(f"mode {mode}: {out} of {4 * n} cars outside the shell after {ticks} ticks")

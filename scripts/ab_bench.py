"""Developer tool (GPU): A/B several builds of the CUDA library on the bench workload in one process.
usage: python scripts/ab_bench.py [--arenas A] [--steps K] [--warmup W] [--mode M] [--cpt C] lib1.so lib2.so ...
Prints arena-ticks/s per build and a hash of the final state: builds that only differ in scheduling must agree bit for bit."""
import argparse, ctypes as C, hashlib, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm
from rocketsim_b200.arena_batch import ArenaBatch
ap = argparse.ArgumentParser()
ap.add_argument("--arenas", type=int, default=4096); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--mode", type=int, default=0); ap.add_argument("--cpt", type=int, default=2); ap.add_argument("--tps", type=int, default=120)
ap.add_argument("libs", nargs="+")
a = ap.parse_args()
SEED = 1234
teams = [0] * a.cpt + [1] * a.cpt
for path in a.libs:
    path = os.path.join(R, path) if not os.path.isabs(path) else path
    L = rlib.load(path); rlib.init_from_mem(sm.mesh_sets(), L)
    b = ArenaBatch(a.mode, a.arenas, teams, seed=SEED, lib=L)
    b.reset_to_random_kickoff((SEED + np.arange(a.arenas)).astype(np.int32)); b.set_auto_reset(True, SEED)
    n = a.warmup + a.steps
    seqs = [torch.empty(a.tps * a.arenas * 2 * a.cpt * 24, dtype=torch.uint8, device="cuda") for _ in range(n)]
    for i, s in enumerate(seqs):
        rlib.check(L, L.rsb_bench_fill_controls_device(b.h, C.c_void_p(s.data_ptr()), a.tps, i * a.tps, SEED))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(a.warmup):
        flush.zero_(); rlib.check(L, L.rsb_step_with_controls_device(b.h, C.c_void_p(seqs[i].data_ptr()), a.tps))
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for k in range(a.steps):
        flush.zero_(); ev[k][0].record()
        rlib.check(L, L.rsb_step_with_controls_device(b.h, C.c_void_p(seqs[a.warmup + k].data_ptr()), a.tps))
        ev[k][1].record()
    torch.cuda.synchronize()
    ms = [x.elapsed_time(y) for x, y in ev]
    cs, bs = b.get_car_states(), b.get_ball_states()
    h = hashlib.sha1(np.ascontiguousarray(cs["pos"]).tobytes() + np.ascontiguousarray(cs["boost"]).tobytes() + np.ascontiguousarray(bs["pos"]).tobytes()).hexdigest()[:12]
    print(f"{os.path.basename(path):28s} {a.arenas * a.tps / (np.mean(ms) * 1e-3) / 1e6:7.2f} M arena-ticks/s  (ms/step mean {np.mean(ms):.2f} min {min(ms):.2f} max {max(ms):.2f})  err {int(b.get_error_flags().max())}  state {h}", flush=True)
    b.close()

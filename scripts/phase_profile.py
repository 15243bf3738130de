"""Developer tool (GPU): per-phase cycle split of the tick kernel, using the -DRSB_PROFILE build.
usage: python scripts/phase_profile.py [arenas] [ticks] [mode] [cars_per_team]"""
import sys, os, ctypes as C, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
from rocketsim_b200 import lib as rlib, build, synthetic_meshes as sm
from rocketsim_b200.arena_batch import ArenaBatch
A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
TICKS = int(sys.argv[2]) if len(sys.argv) > 2 else 240
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cpt = int(sys.argv[4]) if len(sys.argv) > 4 else 2
path = os.path.join(R, "rocketsim_b200", "librsb_prof.so")
L = rlib.load(path); rlib.init_from_mem(sm.mesh_sets(), L)
b = ArenaBatch(mode, A, [0] * cpt + [1] * cpt, seed=1234, lib=L)
b.reset_to_random_kickoff((1234 + np.arange(A)).astype(np.int32)); b.set_auto_reset(True, 1234)
seq = torch.empty(TICKS * A * 2 * cpt * 24, dtype=torch.uint8, device="cuda")
names = ["T0 init", "T1 pretick_begin(lane0)", "T1 wheel rays", "T1 car logic", "T2 pads+ball", "T3 predict", "T3 collide(lane0)",
         "T3 islands(lane0)", "T3 solve(lane0)", "T3 integrate", "T4 car posttick", "T4-6 tail(lane0)"]
out = (C.c_uint64 * 16)()
for rep in range(3):
    rlib.check(L, L.rsb_bench_fill_controls_device(b.h, C.c_void_p(seq.data_ptr()), TICKS, rep * TICKS, 1234))
    L.rsb_debug_profile(out, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rlib.check(L, L.rsb_step_with_controls_device(b.h, C.c_void_p(seq.data_ptr()), TICKS)); b.sync()
    dt = time.perf_counter() - t0
    L.rsb_debug_profile(out, 1)
    tot = sum(out)
    print(f"rep {rep}: {A * TICKS / dt / 1e6:.2f} M arena-ticks/s ({dt * 1e3:.1f} ms)")
    for i, n in enumerate(names):
        print(f"  {n:28s} {100.0 * out[i] / max(tot, 1):5.1f} %   {out[i] / (A * TICKS):9.0f} cycles/arena-tick")

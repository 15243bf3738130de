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
path = os.environ.get("RSB_PROF_LIB") or os.path.join(R, "rocketsim_b200", "librsb_prof.so")
L = rlib.load(path); rlib.init_from_mem(sm.mesh_sets(), L)
b = ArenaBatch(mode, A, [0] * cpt + [1] * cpt, seed=1234, lib=L)
b.reset_to_random_kickoff((1234 + np.arange(A)).astype(np.int32)); b.set_auto_reset(True, 1234)
seq = torch.empty(TICKS * A * 2 * cpt * 24, dtype=torch.uint8, device="cuda")
names = ["INIT", "PRE_BEGIN", "WHEELS", "CAR_LOGIC", "PREDICT", "PAIRS", "NARROW_STATIC", "NARROW_DETECT", "NARROW_COMMIT", "DYN_SOLVE", "FINISH_BODY", "CAR_POST", "PADS_POST", "TAIL", "EXP14", "EXP15"]
out = (C.c_uint64 * 64)()
for rep in range(3):
    rlib.check(L, L.rsb_bench_fill_controls_device(b.h, C.c_void_p(seq.data_ptr()), TICKS, rep * TICKS, 1234))
    L.rsb_debug_profile(out, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rlib.check(L, L.rsb_step_with_controls_device(b.h, C.c_void_p(seq.data_ptr()), TICKS)); b.sync()
    dt = time.perf_counter() - t0
    L.rsb_debug_profile(out, 1)
    tot = sum(out[:16])
    print(f"rep {rep}: {A * TICKS / dt / 1e6:.2f} M arena-ticks/s ({dt * 1e3:.1f} ms)")
    for i, n in enumerate(names):
        nblk = TICKS * min(148, A)  # block-phases executed (one block per SM when the batch fits)
        print(f"  {n:16s} {100.0 * out[i] / max(tot, 1):5.1f} %  wall/block-tick {out[i] / (TICKS * 147):8.0f}  slowest thread {out[48 + i] / (TICKS * 147):8.0f}"
              f"  mean busy thread {out[16 + i] / max(out[32 + i], 1):8.0f}  busy threads/block {out[32 + i] / (TICKS * 147):6.1f}")

"""BASELINE config 1: THE_VOID 1v1, single arena, random inputs (README bench stream), 1 M ticks.
Prints one JSON line with (a) the reference's single-thread CPU ticks/s for the 1 M-tick rollout and (b) the GPU-vs-CPU comparison
of the same rollout at checkpoints 1 / 10 / 120 / 1,000 / ... / 1,000,000 ticks: both sides start from the same kickoff, consume the
same control stream and run FREE (no re-synchronisation); `identical` = every named field of CarState x2 and BallState equal bit
for bit.  usage: python scripts/config1_void.py [ticks=1000000] [arenas=4]   (each arena = its own seed / input stream)"""
import ctypes as C, json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import parity_util as pu
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm, types as T
from oracle import refsim
TICKS = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
NA = int(sys.argv[2]) if len(sys.argv) > 2 else 4
SEED = 1234
L = rlib.load(pu.build_hostcheck()) if os.environ.get("RSB_CONFIG1_HOST") else rlib.load()
ms = sm.mesh_sets(); rlib.init_from_mem(ms, L); refsim.init_from_mem(ms)
mode, teams = T.MODE_THE_VOID, [0, 1]
refs = []
batch = pu.ArenaBatch(mode, NA, teams, lib=L)
for a in range(NA):
    ref = pu.make_ref_arena(refsim, mode, teams)
    ref.seed_rng(SEED + a); ref.reset_kickoff(SEED + a); pu.canonical_kickoff(ref, teams)
    pu.sync_from_ref(batch, ref, a); batch.seed_rng(a, SEED + a); ref.seed_rng(SEED + a)
    refs.append(ref)
checkpoints = sorted({c for c in (1, 10, 120, 1000, 10_000, 100_000, 1_000_000, TICKS) if c <= TICKS})
CH = 2000  # ticks per GPU launch (control sequence generated on the device per launch)
seq = np.zeros(0)
import torch
use_cuda = not os.environ.get("RSB_CONFIG1_HOST")
buf = torch.empty(CH * NA * 2 * 24, dtype=torch.uint8, device="cuda") if use_cuda else np.zeros(CH * NA * 2, T.CarControls)
bufp = C.c_void_p(buf.data_ptr()) if use_cuda else buf.ctypes.data_as(C.c_void_p)
done, cpu_wall, gpu_wall, report, first_diff = 0, 0.0, 0.0, {}, None
for cp in checkpoints:
    n = cp - done
    for a, ref in enumerate(refs):
        w = ref.run_bench_inputs(n, SEED, a, done)
        if a == 0:
            cpu_wall += w
    t0 = time.perf_counter()
    t = done
    while t < cp:
        k = min(CH, cp - t)
        rlib.check(L, L.rsb_bench_fill_controls_device(batch.h, bufp, k, t, SEED))
        rlib.check(L, L.rsb_step_with_controls_device(batch.h, bufp, k))
        t += k
    batch.sync(); gpu_wall += time.perf_counter() - t0
    done = cp
    cs, bs = batch.get_car_states(), batch.get_ball_states()
    same, worst = True, 0.0
    for a, ref in enumerate(refs):
        for i, cid in enumerate(ref.car_ids):
            rc = ref.car_state(cid)
            if not pu.struct_equal(cs[a, i], rc):
                same = False; worst = max(worst, max(pu.compare_car(cs[a, i], rc)[1].values()))
        if not pu.struct_equal(bs[a], ref.ball_state()):
            same = False; worst = max(worst, max(pu.compare_ball(bs[a], ref.ball_state()).values()))
    report[str(cp)] = {"identical": same, "worst_rel_err": worst}
    if not same and first_diff is None:
        first_diff = cp
print(json.dumps({"config": "THE_VOID 1v1, README-bench inputs, free-running rollout", "arenas_compared": NA, "ticks": TICKS,
                  "cpu_reference": {"ticks_per_s_single_thread": TICKS / cpu_wall, "wall_s": cpu_wall, "kind": "oracle/_ref, 1 thread, 1 arena"},
                  "gpu": {"wall_s": gpu_wall, "ticks_per_s_per_arena": TICKS / gpu_wall, "note": f"{NA} arenas advance together; one launch per {CH} ticks"},
                  "divergence_at_ticks": report, "first_checkpoint_with_a_difference": first_diff,
                  "device_error_flags": int(batch.get_error_flags().max()), "library": "host check build" if not use_cuda else "librsb.so (CUDA)"}))

"""Developer tool (GPU): replay a golden record on the CUDA library AND on the host check build, print the
ticks whose single-step error exceeds the tolerance, and diff the two builds' contact/solver probes there.
usage: python scripts/dev_gpu_diag.py tests/golden/soccar_3v3_walls.npz [n_arenas]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ctypes as C
import numpy as np
import parity_util as pu
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm
from rocketsim_b200.arena_batch import ArenaBatch

path = sys.argv[1]
_pos = [x for x in sys.argv[2:] if not x.startswith("--")]
NA = int(_pos[0]) if _pos else 33
with np.load(path) as z:
    rec = {k: z[k] for k in z.files}
ms = sm.mesh_sets()
teams = [int(x) for x in rec["teams"]]; mode = int(rec["mode"])
DUMP = os.path.join(R, "tests", "_hostbuild", os.path.basename(path) + ".hostprobe.npz")
HOST = "--host" in sys.argv
if HOST:  # run here (CPU): dump the host check build's per-tick results + probes for the GPU-side diff
    LH = rlib.load(pu.build_hostcheck()); rlib.init_from_mem(ms, LH)
    bh = ArenaBatch(mode, 1, teams, lib=LH)
else:
    LC = rlib.load(); rlib.init_from_mem(ms, LC)
    bc = ArenaBatch(mode, NA, teams, lib=LC)
    hd = np.load(DUMP, allow_pickle=True) if os.path.exists(DUMP) else None


def probes(b, L, arena):
    out = np.zeros((128, 16), np.float32); n = C.c_int()
    L.rsb_debug_get_manifolds(b.h, arena, rlib.ptr(out), 128, C.byref(n))
    mf = out[: n.value].copy()
    rows = np.zeros((160, 24), np.float32)
    L.rsb_debug_get_solver_rows(b.h, arena, 0, rlib.ptr(rows), 160, C.byref(n))
    rw = rows[: n.value].copy()
    wh = np.zeros((len(teams) * 4, 14), np.float32)
    L.rsb_debug_get_wheels(b.h, arena, rlib.ptr(wh))
    return mf, rw, wh


T_ = rec["controls"].shape[0]
if HOST:
    dump = {}
    for t in range(T_):
        pu.load_pre_state(bh, rec, t, 0); bh.seed_rng(0, int(rec["seed"]))
        bh.set_controls(rec["controls"][t][None]); bh.step(1)
        mf, rw, wh = probes(bh, LH, 0)
        dump[f"mf{t}"], dump[f"rw{t}"], dump[f"wh{t}"] = mf, rw, wh
        dump[f"cs{t}"] = bh.get_car_states()[0]; dump[f"bs{t}"] = bh.get_ball_states()[0]
    np.savez_compressed(DUMP, **dump)
    print("wrote", DUMP); sys.exit(0)

for t in range(T_):
    for a in sorted({0, NA - 1}):
        pu.load_pre_state(bc, rec, t, a); bc.seed_rng(a, int(rec["seed"]))
    bc.set_controls(np.broadcast_to(rec["controls"][t], (NA, len(teams))))
    bc.step(1)
    cs, bs_ = bc.get_car_states(), bc.get_ball_states()
    for a in sorted({0, NA - 1}):
        worst = 0.0; what = None
        for i in range(len(teams)):
            bad, errs = pu.compare_car(cs[a, i], rec["car_post"][t, i])
            for k, v in errs.items():
                if v > worst: worst, what = v, (i, k)
            if bad: print(f"tick {t} arena {a} car {i} DISCRETE {bad}", flush=True)
        for k, v in pu.compare_ball(bs_[a], rec["ball_post"][t]).items():
            if v > worst: worst, what = v, ("ball", k)
        if worst > 1e-4:
            print(f"tick {t} arena {a}: cuda err {worst:.3g} at {what}; illcond={int(rec['wheel_on_dynamic'][t])}", flush=True)
            mc, rc, wc = probes(bc, LC, a)
            mi4 = np.zeros(4, np.int32); LC.rsb_debug_get_misc(bc.h, a, rlib.ptr(mi4)); print('  misc', mi4, np.array([mi4[3]], np.int32).view(np.float32))
            if hd is None: continue
            mh, rh, wh = hd[f"mf{t}"], hd[f"rw{t}"], hd[f"wh{t}"]
            hworst = max(max(pu.compare_car(hd[f"cs{t}"][i], rec["car_post"][t, i])[1].values()) for i in range(len(teams)))
            print(f"  host build err {hworst:.3g}; manifold points cuda {len(mc)} host {len(mh)}; rows cuda {len(rc)} host {len(rh)}")
            if len(mc) == len(mh) and len(mc):
                d = np.abs(mc - mh).max(axis=1)
                for j in np.argsort(-d)[:4]:
                    print(f"   pt {j} maxdiff {d[j]:.3g}\n     cuda {np.array2string(mc[j], precision=5)}\n     host {np.array2string(mh[j], precision=5)}")
            else:
                print("   cuda pts:\n", np.array2string(mc[:, :7], precision=4)); print("   host pts:\n", np.array2string(mh[:, :7], precision=4))
            if len(rc) == len(rh) and len(rc):
                d = np.abs(rc - rh).max(axis=1); print("  solver row maxdiff:", np.array2string(d, precision=3))
            dw = np.abs(wc - wh).max(axis=1)
            print("  wheel probe maxdiff per wheel:", np.array2string(dw, precision=3), flush=True)
print("error flags", bc.get_error_flags().max())

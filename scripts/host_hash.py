"""Developer tool (CPU): bit-exact regression of the tick arithmetic.  Runs the host check build of the device code on the
bench workload (README inputs, auto reset) for several modes and prints a hash of every arena's final state; a restructuring
that only changes WHERE / in which order independent work runs must leave every hash unchanged.
usage: python scripts/host_hash.py [hostcheck.so] [ticks] [arenas]"""
import ctypes as C, hashlib, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm, types as T
from rocketsim_b200.arena_batch import ArenaBatch
path = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "-" else None
if path is None:
    import parity_util as pu
    path = pu.build_hostcheck()
TICKS = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
A = int(sys.argv[3]) if len(sys.argv) > 3 else 12
L = rlib.load(path); rlib.init_from_mem(sm.mesh_sets(), L)
for name, mode, cpt in (("soccar_2v2", 0, 2), ("soccar_3v3", 0, 3), ("hoops_2v2", 1, 2), ("heatseeker_1v1", 2, 1), ("snowday_3v3", 3, 3), ("void_1v1", 5, 1)):
    teams = [0] * cpt + [1] * cpt
    b = ArenaBatch(mode, A, teams, seed=99, lib=L)
    b.reset_to_random_kickoff((99 + np.arange(A)).astype(np.int32)); b.set_auto_reset(True, 99)
    # bench inputs generated on the "device" (host build: plain memory)
    n = TICKS * A * 2 * cpt
    seq = np.zeros(n, T.CarControls)
    rlib.check(L, L.rsb_bench_fill_controls_device(b.h, seq.ctypes.data_as(C.c_void_p), TICKS, 0, 99))
    step = 250
    ev = 0
    for t0 in range(0, TICKS, step):
        nt = min(step, TICKS - t0)
        rlib.check(L, L.rsb_step_with_controls_device(b.h, C.c_void_p(seq.ctypes.data + t0 * A * 2 * cpt * 24), nt))
        ev += len(b.get_events())
    cs, bs, ps = b.get_car_states(), b.get_ball_states(), (b.get_pad_states() if b.n_pads else np.zeros(0))
    h = hashlib.sha1()
    for f in ("pos", "vel", "angVel", "rotMat", "boost", "isDemoed", "isOnGround", "jumpTime", "flipTime", "supersonicTime"):
        h.update(np.ascontiguousarray(cs[f]).tobytes())
    for f in ("pos", "vel", "angVel"):
        h.update(np.ascontiguousarray(bs[f]).tobytes())
    if len(ps):
        h.update(np.ascontiguousarray(ps["cooldown"]).tobytes())
    print(f"{name:16s} {h.hexdigest()[:16]}  events {ev}  err {int(b.get_error_flags().max())}", flush=True)
    b.close()

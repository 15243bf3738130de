"""Developer tool: replay ONE tick of a golden record on the oracle and on the host check build and
print contact points and solver rows side by side.  usage: dev_tick_probe.py NAME TICK"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, ctypes as C, parity_util as pu
from rocketsim_b200 import lib, synthetic_meshes as sm
from rocketsim_b200.arena_batch import ArenaBatch
from oracle import refsim
np.set_printoptions(precision=8, suppress=True, linewidth=220)
L = lib.load(pu.build_hostcheck()); lib.init_from_mem(sm.mesh_sets(), L); refsim.init_from_mem(sm.mesh_sets())
name, t = sys.argv[1], int(sys.argv[2])
z = np.load(os.path.join(R, "tests", "golden", name + ".npz")); rec = {k: z[k] for k in z.files}
teams = [int(x) for x in rec["teams"]]
b = ArenaBatch(int(rec["mode"]), 1, teams, lib=L)
pu.load_pre_state(b, rec, t); b.seed_rng(0, int(rec["seed"])); b.set_controls(rec["controls"][t][None]); b.step(1)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
out = np.zeros((64, 16), np.float32); n = C.c_int()
L.rsb_debug_get_manifolds(b.h, 0, vp(out), 64, C.byref(n)); mine = out[: n.value]
ref = pu.make_ref_arena(refsim, int(rec["mode"]), teams)
for i, cid in enumerate(ref.car_ids):
    ref.set_car_state(cid, rec["car_pre"][t, i]); ref.set_car_internal(cid, rec["carint_pre"][t, i]); ref.set_controls(cid, rec["controls"][t, i])
ref.set_ball_state(rec["ball_pre"][t]); ref.set_ball_internal(rec["ballint_pre"][t]); ref.set_pad_states(rec["pads_pre"][t])
ref.step(1); rm = ref.manifolds()
print("contacts mine/ref", len(mine), len(rm))
if len(mine) == len(rm) and len(rm):
    print("max abs diff contacts", np.abs(mine - rm).max(0))
else:
    print(mine); print(rm)
for which in (0, 1):
    rows = np.zeros((64, 24), np.float32)
    L.rsb_debug_get_solver_rows(b.h, 0, which, vp(rows), 64, C.byref(n)); mr = rows[: n.value]
    rr = ref.solver_rows(which, n.value)
    print("rows", which, n.value)
    for i in range(n.value):
        print(" mine", mr[i, 18:]); print(" ref ", rr[i, 18:]); print(" jac diff", np.abs(mr[i, :18] - rr[i, :18]).max())
sb = np.zeros((16, 12), np.float32); L.rsb_debug_get_solver_bodies(b.h, 0, vp(sb), 16, C.byref(n)); sb = sb[: n.value]
rb = ref.solver_bodies(n.value)
print('solver bodies mine'); print(sb); print('ref'); print(rb)
cs = b.get_car_states()[0]
for i in range(len(teams)):
    print("car", i, "vel diff", pu.field_as_array(cs[i], "vel") - pu.field_as_array(rec["car_post"][t, i], "vel"),
          "angVel diff", pu.field_as_array(cs[i], "angVel") - pu.field_as_array(rec["car_post"][t, i], "angVel"))
wd = np.zeros((len(teams) * 4, 14), np.float32); L.rsb_debug_get_wheels(b.h, 0, vp(wd))
for i, cid in enumerate(ref.car_ids):
    rw = ref.wheel_debug(cid); mw = wd[4 * i: 4 * i + 4]
    d = np.abs(rw - mw).max()
    if d > 1e-6:
        print("car", i, "wheel diff", d); print(mw); print(rw)
for which in (0, 1):
    rows = np.zeros((64, 24), np.float32)
    L.rsb_debug_get_solver_rows(b.h, 0, which, vp(rows), 64, C.byref(n)); mr = rows[: n.value]
    rr = ref.solver_rows(which, n.value)
    for i in range(n.value):
        print("row", which, i, "n1 mine", mr[i, 0:3], "ref", rr[i, 0:3], "rp1xn mine", mr[i, 3:6], "ref", rr[i, 3:6])
print("solver bodies full"); print(sb); print(rb)

"""Developer tool (GPU, under compute-sanitizer): a short mixed workload -- 3v3 Soccar with cars driven at the ball and at the
walls -- so that memcheck / racecheck / initcheck see ground contact, mesh contacts, car-ball, car-car, EPA and the solver."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import parity_util as pu
from rocketsim_b200 import lib as rlib, synthetic_meshes as sm, types as T
from rocketsim_b200.arena_batch import ArenaBatch
L = rlib.load(); rlib.init_from_mem(sm.mesh_sets(), L)
A, TICKS = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 90
for mode, cpt in ((T.MODE_SOCCAR, 3), (T.MODE_SNOWDAY, 2), (T.MODE_DROPSHOT, 2)):
    teams = [0] * cpt + [1] * cpt
    b = ArenaBatch(mode, A, teams, seed=3, lib=L)
    b.reset_to_random_kickoff((3 + np.arange(A)).astype(np.int32)); b.set_auto_reset(True, 3)
    cs = b.get_car_states()
    # half of the arenas: cars next to the walls heading into them (mesh + EPA paths)
    for a in range(0, A, 2):
        for i in range(len(teams)):
            cs[a, i]["pos"]["x"] = 3600.0 * (1 if i % 2 else -1); cs[a, i]["pos"]["y"] = 800.0 * i - 2000; cs[a, i]["vel"]["x"] = 1800.0 * (1 if i % 2 else -1)
    ai = np.repeat(np.arange(A), len(teams)); ci = np.tile(np.arange(len(teams)), A)
    b.set_car_states(ai, ci, cs.ravel())
    if mode == T.MODE_DROPSHOT:  # damaged tiles + a charged ball coming down on them: tile hulls (GJK), damage, the drop-through goal
        tl = np.ones((A, 2, 70), np.uint8); b.set_dropshot_tiles(np.arange(A), tl)
        bs = b.get_ball_states()
        bs["pos"]["x"] = 700.0; bs["pos"]["y"] = 1500.0; bs["pos"]["z"] = 300.0; bs["vel"]["z"] = -1500.0
        bs["dsInfo"]["chargeLevel"] = 3; bs["dsInfo"]["yTargetDir"] = 1.0
        b.set_ball_states(np.arange(A), bs)
    ctrl = np.zeros((TICKS, A, len(teams)), T.CarControls); ctrl["throttle"] = 1; ctrl["boost"] = 1
    ctrl["jump"][::7] = 1; ctrl["pitch"] = -0.3
    b.step_with_controls(ctrl)
    print("mode", mode, "ticks", b.get_tick_counts().min(), "errors", b.get_error_flags().max(), "events", len(b.get_events()))
    b.close()

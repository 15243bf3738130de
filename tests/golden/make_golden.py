"""Generates the golden single-step records under tests/golden/*.npz from the ORACLE
(the unmodified reference compiled into oracle/_ref).  Run in the build container:

    python tests/golden/make_golden.py

Each record holds, for every tick of an oracle rollout driven by the README-style control
schedule, the complete pre-tick state (CarState + hidden carry-over state + BallState + pads),
the controls and the post-tick state, so every tick is an independent single-step parity case
(tests/test_golden.py replays them without needing the oracle).  Synthetic .cmf meshes
(rocketsim_b200.synthetic_meshes) are used on both sides; see SURVEY.md F4.
"""
import os
import sys

import numpy as np

ONLY = sys.argv[1:]  # optional: regenerate only the named cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_util as pu  # noqa: E402
from rocketsim_b200 import synthetic_meshes as sm, types as T  # noqa: E402
from oracle import refsim  # noqa: E402

# name: (mode, cars per team, seed, ticks, forced-drive ticks, aim cars at ball)
CASES = {
    "void_1v1": (T.MODE_THE_VOID, 1, 3, 240, 0, False),
    "soccar_2v2_kickoff": (T.MODE_SOCCAR, 2, 11, 300, 120, False),
    "soccar_2v2_aim": (T.MODE_SOCCAR, 2, 5, 240, 90, True),
    "soccar_3v3_aim": (T.MODE_SOCCAR, 3, 7, 240, 90, True),
    "hoops_2v2_aim": (T.MODE_HOOPS, 2, 9, 240, 60, True),
    "heatseeker_1v1": (T.MODE_HEATSEEKER, 1, 13, 240, 60, False),
    "snowday_3v3_aim": (T.MODE_SNOWDAY, 3, 17, 240, 90, True),
    # cars driven into side walls / back walls / corners, ball fired at a ramp: arena-mesh BVH paths
    "soccar_3v3_walls": (T.MODE_SOCCAR, 3, 21, 300, 150, "walls"),
    "hoops_2v2_walls": (T.MODE_HOOPS, 2, 23, 300, 150, "walls"),
    "snowday_2v2_walls": (T.MODE_SNOWDAY, 2, 25, 300, 150, "walls"),
    # demos: a scripted supersonic ram (the attacker crosses the supersonic threshold on the way in), the demoed ticks and the
    # respawn (3 s = 360 ticks later); tight circles around the ball (bumps, demos once a car is supersonic); goals with cars in play
    "soccar_3v3_demo_ram": (T.MODE_SOCCAR, 3, 31, 520, 140, "ram"),
    "soccar_2v2_demo_charge": (T.MODE_SOCCAR, 2, 11, 500, 150, "charge"),
    "hoops_2v2_demo_charge": (T.MODE_HOOPS, 2, 12, 500, 150, "charge"),
    "soccar_3v3_demo_charge": (T.MODE_SOCCAR, 3, 13, 500, 150, "charge"),
    "snowday_3v3_demo_charge": (T.MODE_SNOWDAY, 3, 14, 500, 150, "charge"),
    "soccar_2v2_goal": (T.MODE_SOCCAR, 2, 41, 180, 30, "goal"),
    "hoops_2v2_goal": (T.MODE_HOOPS, 2, 43, 240, 30, "goal"),
    # Dropshot (tile hulls, DROPSHOT_FLOOR plane, ball launch, tile damage, charge levels, the ball falling through = goal)
    "dropshot_2v2_kickoff": (T.MODE_DROPSHOT, 2, 51, 300, 150, False),
    "dropshot_2v2_charge": (T.MODE_DROPSHOT, 2, 53, 500, 150, "ds_charge"),
    "dropshot_2v2_drop": (T.MODE_DROPSHOT, 2, 59, 340, 60, "ds_drop"),
    "dropshot_3v3_slam": (T.MODE_DROPSHOT, 3, 55, 420, 90, "ds_slam"),
    "dropshot_2v2_walls": (T.MODE_DROPSHOT, 2, 57, 300, 150, "walls"),
}


def main():
    refsim.init_from_mem(sm.mesh_sets())
    for name, (mode, cpt, seed, ticks, drive, aim) in CASES.items():
        if ONLY and name not in ONLY:
            continue
        teams = [0] * cpt + [1] * cpt
        rec = pu.record_rollout(refsim, mode, teams, seed, ticks, drive=drive, aim=aim)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        ev = rec["events"]
        print(f"{name}: {ticks} ticks, {len(ev)} events (goal/bump/demo = {[int((ev[:, 1] == k).sum()) for k in (1, 2, 3)] if len(ev) else 0}), "
              f"demoed car-ticks {int(rec['car_post']['isDemoed'].sum())}, supersonic car-ticks {int(rec['car_post']['isSupersonic'].sum())}, "
              f"scored ticks {int(rec['scored'].sum())}, " + (f"tile changes {int((rec['tiles_pre'] != rec['tiles_post']).sum())}, broken at end {int((rec['tiles_post'][-1] == 2).sum())}, max charge {int(rec['ball_post']['dsInfo']['chargeLevel'].max())}, " if 'tiles_pre' in rec else "") + f" wheel-on-dynamic ticks {int(rec['wheel_on_dynamic'].sum())}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

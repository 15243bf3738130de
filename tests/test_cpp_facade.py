"""The header-only C++ front end (include/rocketsim_b200.hpp: Arena::Create / AddCar / car->controls / Step / GetState /
callbacks, the reference's own vocabulary) against the oracle: a small C++ program drives one arena exactly like
reference user code would; the same script runs on the reference and the dumped states must agree."""
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import lib as rlib, types as T

ROOT = pu.ROOT
SRC = os.path.join(ROOT, "tests", "native", "facade_check.cpp")
TICKS, SEED = 150, 7
# a FREE rollout of 150 ticks (no resynchronisation): ULP differences amplify through jumps and contacts, so this is a
# drift bound, not the single-step bar (that one is gated per tick in test_golden / test_rollout / test_presets)
ROLLOUT_TOL = 5e-3


def build_and_run(lib_path, tmp_path, mesh_sets):
    exe = str(tmp_path / "facade_check")
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.run(["g++", "-std=c++17", "-O1", SRC, "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-l:" + libname,
                    "-Wl,-rpath," + libdir, "-o", exe], check=True, capture_output=True, text=True)
    paths = []
    for i, data in enumerate(mesh_sets[T.MODE_SOCCAR]):
        p = str(tmp_path / f"mesh_{i}.cmf")
        open(p, "wb").write(data)
        paths.append(p)
    r = subprocess.run([exe, str(T.MODE_SOCCAR), str(TICKS), str(SEED)] + paths, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = {}
    for line in r.stdout.splitlines():
        k, _, v = line.partition(" ")
        out[k] = v
    return out


def reference_run(oracle):
    ref = oracle.RefArena(T.MODE_SOCCAR)
    ida, idb = ref.add_car(0, T.CAR_OCTANE), ref.add_car(1, T.CAR_DOMINUS)
    ref.reset_kickoff(SEED)
    for t in range(TICKS):
        ca, cb = np.zeros((), T.CarControls), np.zeros((), T.CarControls)
        ca["throttle"] = 1; ca["boost"] = 1; ca["steer"] = 0.3 if (t % 40 < 20) else -0.2
        cb["throttle"] = 1; cb["jump"] = (t % 50) < 5; cb["pitch"] = -0.5
        ref.set_controls(ida, ca); ref.set_controls(idb, cb)
        ref.step(1)
    res = (ref.car_state(ida), ref.car_state(idb), ref.ball_state())
    ref.close()
    return res


def check(out, oracle):
    ra, rb, rball = reference_run(oracle)
    n = T.CarState.fields["lastControls"][1]
    for tag, ref in (("carA", ra), ("carB", rb)):
        raw = bytes.fromhex(out[tag]) + bytes(T.CarState.itemsize - n)
        mine = np.frombuffer(raw, T.CarState)[0]
        bad, errs = pu.compare_car(mine, ref)
        assert not bad, (tag, bad)
        assert max(errs.values()) <= ROLLOUT_TOL, (tag, errs)
    ball = np.frombuffer(bytes.fromhex(out["ball"]) + bytes(T.BallState.itemsize - 104), T.BallState)[0]
    assert max(pu.compare_ball(ball, rball).values()) <= ROLLOUT_TOL
    assert out["ticks"].split() == [str(TICKS), "scored", "0"]
    g = out["goals"].split()
    assert int(g[0]) == 1 and int(g[2]) == T.TEAM_BLUE  # ball crossed +y: blue scores, exactly one callback
    assert out["errors"] == "3"


def test_cpp_facade_hostcheck(hostlib, oracle, mesh_sets, tmp_path):
    check(build_and_run(pu.HOSTCHECK_LIB, tmp_path, mesh_sets), oracle)


@pytest.mark.gpu
def test_cpp_facade_cuda(cudalib, oracle, mesh_sets, tmp_path):
    check(build_and_run(rlib.LIB_PATH, tmp_path, mesh_sets), oracle)

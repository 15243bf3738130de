"""Shared machinery of the parity tests: drive the oracle (unmodified reference) and an
``ArenaBatch`` (CUDA library, or the host-compiled check build) with identical inputs and
compare them field by field.

Discrete CarState/BallState fields must match bit-exactly; float fields are compared with
``|a-b| / max(|b|, eps_field)`` (eps pos 1 uu, vel 1 uu/s, angVel 0.01, rotMat 1).
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rocketsim_b200 import types as T  # noqa: E402
from rocketsim_b200 import lib as rsb_lib  # noqa: E402
from rocketsim_b200 import synthetic_meshes  # noqa: E402
from rocketsim_b200.arena_batch import ArenaBatch  # noqa: E402

HOSTCHECK_DIR = os.path.join(ROOT, "tests", "_hostbuild")
HOSTCHECK_LIB = os.path.join(HOSTCHECK_DIR, "librsb_hostcheck.so")

DISCRETE_CAR = ["isOnGround", "wheelsWithContact", "hasJumped", "hasDoubleJumped", "hasFlipped", "isFlipping", "isJumping",
                "isBoosting", "isSupersonic", "isAutoFlipping", "isDemoed"]
FLOAT_CAR = {"pos": 1.0, "vel": 1.0, "angVel": 0.01, "rotMat": 1.0, "boost": 1.0, "jumpTime": 1e-2, "flipTime": 1e-2,
             "airTime": 1e-2, "airTimeSinceJump": 1e-2, "timeSinceBoosted": 1e-2, "boostingTime": 1e-2,
             "supersonicTime": 1e-2, "handbrakeVal": 1e-2, "autoFlipTimer": 1e-2, "demoRespawnTimer": 1e-2,
             "flipRelTorque": 1.0}
FLOAT_BALL = {"pos": 1.0, "vel": 1.0, "angVel": 0.01}


def build_hostcheck(force: bool = False) -> str:
    """g++ build of the device code for the CPU-only tier (tests only; see rsb_api.cu header)."""
    src = os.path.join(ROOT, "rocketsim_b200", "csrc", "rsb_api.cu")
    os.makedirs(HOSTCHECK_DIR, exist_ok=True)
    newest = 0.0
    for d, _, files in os.walk(os.path.join(ROOT, "rocketsim_b200", "csrc")):
        for f in files:
            newest = max(newest, os.path.getmtime(os.path.join(d, f)))
    for f in os.listdir(os.path.join(ROOT, "include")):
        newest = max(newest, os.path.getmtime(os.path.join(ROOT, "include", f)))
    if force or not os.path.exists(HOSTCHECK_LIB) or os.path.getmtime(HOSTCHECK_LIB) < newest:
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DRSB_HOSTCHECK", *os.environ.get("RSB_HOSTCHECK_FLAGS", "").split(), "-x", "c++", src,
               "-o", HOSTCHECK_LIB, "-I" + os.path.join(ROOT, "rocketsim_b200", "csrc")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("hostcheck build failed:\n" + r.stderr[-4000:])
    return HOSTCHECK_LIB


def vec3(v):
    return np.array([v["x"], v["y"], v["z"]], dtype=np.float64)


def rel_err(a, b, eps):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), eps))) if a.size else 0.0


def field_as_array(s, name):
    v = s[name]
    if v.dtype == T.Vec:
        return np.stack([v["x"], v["y"], v["z"]], -1)
    if v.dtype == T.RotMat:
        return np.stack([field_as_array(v, k) for k in ("forward", "right", "up")], -1)
    return np.asarray(v)


def compare_car(mine, ref):
    """returns (discrete mismatches list, {field: rel err})"""
    bad = [f for f in DISCRETE_CAR if not np.array_equal(mine[f], ref[f])]
    errs = {f: rel_err(field_as_array(mine, f), field_as_array(ref, f), eps) for f, eps in FLOAT_CAR.items()}
    return bad, errs


def compare_ball(mine, ref):
    return {f: rel_err(field_as_array(mine, f), field_as_array(ref, f), eps) for f, eps in FLOAT_BALL.items()}


def random_controls(rng, n):
    c = np.zeros(n, dtype=T.CarControls)
    for f in ("throttle", "steer", "pitch", "yaw", "roll"):
        c[f] = rng.uniform(-1, 1, n).astype(np.float32)
    c["jump"] = rng.random(n) < 0.25
    c["boost"] = rng.random(n) < 0.5
    c["handbrake"] = rng.random(n) < 0.2
    return c


class ControlSchedule:
    """README-style inputs: each car redraws its controls every 2..60 ticks."""

    def __init__(self, n_cars, seed):
        self.rng = np.random.default_rng(seed)
        self.n = n_cars
        self.next = np.zeros(n_cars, dtype=np.int64)
        self.cur = np.zeros(n_cars, dtype=T.CarControls)

    def at(self, tick):
        for i in range(self.n):
            if tick >= self.next[i]:
                self.cur[i] = random_controls(self.rng, 1)[0]
                self.next[i] = tick + 2 + int(self.rng.integers(0, 59))
        return self.cur.copy()


def sync_from_ref(batch: ArenaBatch, ref, arena=0):
    """Copy the complete reference state (incl. hidden carry-over state) into arena `arena`."""
    n = len(ref.car_ids)
    if n:
        cs = np.array([ref.car_state(i) for i in ref.car_ids], dtype=T.CarState)
        ci = np.array([ref.car_internal(i) for i in ref.car_ids], dtype=T.CarInternal)
        ai = np.full(n, arena, dtype=np.uint32)
        idx = np.arange(n, dtype=np.uint32)
        batch.set_car_states(ai, idx, cs)
        batch.set_car_internals(ai, idx, ci)
    batch.set_ball_states([arena], np.array([ref.ball_state()], dtype=T.BallState))
    batch.set_ball_internals([arena], np.array([ref.ball_internal()], dtype=T.BallInternal))
    if batch.n_pads:
        batch.set_pad_states([arena], ref.pad_states())

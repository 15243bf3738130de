"""Unit checks of device functions against the reference's own classes, compiled as small C++
programs that include the Bullet headers from /root/reference and link oracle/_ref (CPU tier only:
the headers do not exist on the GPU box)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def build(name):
    if not os.path.isdir(REF):
        pytest.skip("/root/reference absent")
    out_dir = os.path.join(ROOT, "tests", "_hostbuild")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, name)
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-w", f"-I{REF}/libsrc/bullet3-3.24", f"-I{ROOT}/rocketsim_b200/csrc",
           os.path.join(ROOT, "tests", "native", name + ".cpp"), "-o", exe, f"-L{ROOT}/oracle/_ref", "-lrocketsim_ref",
           f"-Wl,-rpath,{ROOT}/oracle/_ref"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_ray_vs_convex_cast_matches_bullet(oracle):
    """ray_convex (rsb_ray.h) vs btSubsimplexConvexCast on 20k random box / sphere casts.  The cast stops at
    dist^2 < 1e-4 and the reference normalises with rsqrtss, so a few ulp of input difference can move an
    individual hit by ~1e-2; the check is therefore statistical: same hit/miss decision and fraction within
    1e-3 for > 97 % of the hits."""
    exe = build("ray_convex_check")
    out = subprocess.run([exe, "20000"], capture_output=True, text=True).stdout
    m = re.search(r"hits_ref (\d+) hits_mine (\d+) hit_mismatch (\d+) fraction_diff_gt_1e-3 (\d+)", out)
    assert m, out
    hits_ref, hits_mine, mismatch, loose = map(int, m.groups())
    assert hits_ref > 5000
    assert mismatch <= 0.002 * 20000, out
    assert loose <= 0.03 * hits_ref, out

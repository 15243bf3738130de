"""Developer tool: worst single-step float error per golden record and the share of bit-identical steps, for the CUDA library
(default) or the host check build (`host`).  usage: python scripts/golden_residue.py [host]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import parity_util as pu, test_golden as tg
from rocketsim_b200 import lib, synthetic_meshes as sm
L = lib.load(pu.build_hostcheck()) if (len(sys.argv) > 1 and sys.argv[1] == "host") else lib.load()
lib.init_from_mem(sm.mesh_sets(), L)
tot, steps = 0.0, 0
for p in sorted(tg.GOLDEN):
    rec = tg.load(p)
    r = pu.replay_record(rec, L)
    allw = dict(r["worst"]); 
    for k, v in r["worst_illcond"].items(): allw[k] = max(allw.get(k, 0.0), v)
    w = max(allw.values()); tot = max(tot, w); steps += r["ticks"]
    print(f"{os.path.basename(p):32s} ticks {r['ticks']:4d} worst {w:.2e}  " + ", ".join(f"{k}={v:.1e}" for k, v in sorted(allw.items(), key=lambda kv: -kv[1])[:4]) + (f"  MISMATCH {r['discrete_mismatches'][:2]}" if r["discrete_mismatches"] else ""), flush=True)
print(f"worst over {steps} single steps: {tot:.3g}")

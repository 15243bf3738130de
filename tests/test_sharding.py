"""Multi-GPU host logic on CPU: world-size-2 gloo job (independent shards, no data-path collective).
Each rank steps its shard with the host check build of the device code (tests only) and rank 0 gathers; the result must
equal a single-process run of the whole job bit for bit, i.e. per-arena results do not depend on the sharding."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import parity_util as pu
from rocketsim_b200 import types as T
from rocketsim_b200.sharding import ShardedArenaBatch, shard_bounds

N_TOTAL, TICKS, SEED = 7, 40, 21
TEAMS = [0, 1]


def test_shard_bounds_cover_everything_once():
    for n in (1, 7, 8, 4096, 65536, 32769):
        for w in (1, 2, 3, 4, 8):
            if n < w:
                continue
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _controls():
    out = np.zeros((TICKS, N_TOTAL, len(TEAMS)), T.CarControls)
    for a in range(N_TOTAL):
        s = pu.ControlSchedule(len(TEAMS), 1000 + a)  # keyed by the GLOBAL arena index
        for t in range(TICKS):
            out[t, a] = s.at(t)
    out["throttle"] = 1.0
    return out


def _run_shard(rank, world, lib_path):
    from rocketsim_b200 import lib as rlib, synthetic_meshes as sm
    L = rlib.load(lib_path)
    rlib.init_from_mem(sm.mesh_sets(), L)
    sh = ShardedArenaBatch(T.MODE_SOCCAR, N_TOTAL, TEAMS, seed=SEED, rank=rank, world_size=world, device=0, lib=L)
    sh.reset_to_random_kickoff(SEED)
    ctrl = _controls()
    sh.batch.step_with_controls(np.ascontiguousarray(sh.local_slice(ctrl, axis=1)))
    return sh, sh.batch.get_car_states(), sh.batch.get_ball_states()


def _worker(rank, world, port, lib_path, out_path):
    import torch.distributed as dist
    for p in (pu.ROOT, os.path.join(pu.ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sh, cars, ball = _run_shard(rank, world, lib_path)
    g_cars, g_ball = sh.gather_to_rank0(cars), sh.gather_to_rank0(ball)
    if rank == 0:
        np.save(out_path + ".cars.npy", g_cars); np.save(out_path + ".ball.npy", g_ball)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_job_equals_single_process(hostlib, tmp_path):
    lib_path = pu.HOSTCHECK_LIB
    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, lib_path, out), nprocs=2, join=True)
    _, cars, ball = _run_shard(0, 1, lib_path)
    assert pu.struct_equal(np.load(out + ".cars.npy"), cars)
    assert pu.struct_equal(np.load(out + ".ball.npy"), ball)
    assert len({cars[a].tobytes() for a in range(N_TOTAL)}) == N_TOTAL  # distinct arenas really ran


def test_auto_reset_and_events_use_global_arena_index(hostlib):
    """Device-side auto reset seeds and event arena ids are functions of the GLOBAL arena index (rsb_set_arena_offset): a
    shard gives the same per-arena results and the same event records as the corresponding arenas of the whole job."""
    def run(rank, world):
        sh = ShardedArenaBatch(T.MODE_SOCCAR, 6, TEAMS, seed=SEED, rank=rank, world_size=world, device=0, lib=hostlib)
        sh.reset_to_random_kickoff(SEED)
        sh.set_auto_reset(True, 4242)
        b = sh.batch
        bs = b.get_ball_states()
        bs["pos"]["y"] = 4700; bs["pos"]["z"] = 200; bs["vel"]["y"] = 3000 + 100 * sh.global_ids   # every arena scores, at its own tick
        b.set_ball_states(np.arange(b.n_arenas), bs)
        b.step(60)
        out = sh.batch.get_car_states(), sh.batch.get_ball_states(), sh.get_events()
        sh.close()
        return out
    cars, ball, ev = run(0, 1)
    assert len(ev) == 6 and sorted(ev["arena"]) == list(range(6))
    for rank in (0, 1):
        c, bl, e = run(rank, 2)
        lo, hi = shard_bounds(6, 2, rank)
        assert pu.struct_equal(c, cars[lo:hi]) and pu.struct_equal(bl, ball[lo:hi])
        assert pu.struct_equal(e, ev[(ev["arena"] >= lo) & (ev["arena"] < hi)])

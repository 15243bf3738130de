"""The rest of the C-ABI surface: compact observations, device-pointer variants, stream binding, and the Python
BallPredTracker front end.  CPU tier through the host check build where no device pointer is involved; GPU tier for the rest."""
import ctypes as C

import numpy as np
import pytest

import parity_util as pu
from rocketsim_b200 import lib as rlib, types as T
from rocketsim_b200.arena_batch import ArenaBatch, BallPredTracker


def observations_match_states(lib):
    teams = [0, 0, 1, 1]
    b = ArenaBatch(T.MODE_SOCCAR, 6, teams, seed=2, lib=lib)
    b.reset_to_random_kickoff(np.arange(6, dtype=np.int32))
    ctrl = np.zeros((6, 4), T.CarControls); ctrl["throttle"] = 1; ctrl["boost"] = 1; ctrl["jump"][:, 1] = 1
    b.set_controls(ctrl); b.step(40)
    obs, cs, bs = b.get_observations(), b.get_car_states(), b.get_ball_states()
    assert obs.shape == (6, 9 + 17 * 4)
    for a in range(6):
        assert np.array_equal(obs[a, 0:3], [bs[a]["pos"][k] for k in "xyz"]) and np.array_equal(obs[a, 3:6], [bs[a]["vel"][k] for k in "xyz"])
        for c in range(4):
            q = obs[a, 9 + 17 * c: 9 + 17 * (c + 1)]
            s = cs[a, c]
            assert np.array_equal(q[0:3], [s["pos"][k] for k in "xyz"])
            assert np.array_equal(q[3:6], [s["rotMat"]["forward"][k] for k in "xyz"]) and np.array_equal(q[6:9], [s["rotMat"]["up"][k] for k in "xyz"])
            assert np.array_equal(q[9:12], [s["vel"][k] for k in "xyz"]) and np.array_equal(q[12:15], [s["angVel"][k] for k in "xyz"])
            assert q[15] == s["boost"]
            flags = int(q[16])
            assert bool(flags & 1) == bool(s["isOnGround"]) and bool(flags & 2) == bool(s["isDemoed"]) and bool(flags & 4) == bool(s["isSupersonic"])
    b.close()


def test_observations_hostcheck(hostlib):
    observations_match_states(hostlib)


def test_ball_pred_tracker_front_end(hostlib, oracle):
    """Python mirror of BallPredTracker (UpdatePredFromArena / GetBallStateForTime) on top of rsb_predict_ball."""
    b = ArenaBatch(T.MODE_SOCCAR, 3, [0, 1], lib=hostlib)
    bs = b.get_ball_states()
    bs["vel"]["x"] = 900; bs["vel"]["z"] = 700
    b.set_ball_states(np.arange(3), bs)
    trk = BallPredTracker(b, num_pred_ticks=120)
    trk.update_pred_from_batch()
    assert trk.pred_data.shape == (120, 3)
    s0, s1 = trk.get_ball_state_for_time(0.0), trk.get_ball_state_for_time(0.5)
    assert pu.struct_equal(s0, trk.pred_data[0]) and pu.struct_equal(s1, trk.pred_data[60])
    assert pu.struct_equal(trk.get_ball_state_for_time(99.0), trk.pred_data[119])  # clamped like RS_CLAMP in the reference
    # the prediction is what the arenas then do (ball only: no car is near)
    b.step(60)
    assert max(pu.compare_ball(b.get_ball_states()[1], s1[1]).values()) <= 1e-4
    b.close()


@pytest.mark.gpu
def test_observations_cuda(cudalib):
    observations_match_states(cudalib)


@pytest.mark.gpu
def test_device_pointer_variants_and_stream(cudalib):
    import torch
    teams = [0, 1]
    A, TK = 50, 24
    seeds = np.arange(A, dtype=np.int32)
    rng = np.random.default_rng(0)
    seq = np.zeros((TK, A, 2), T.CarControls)
    seq["throttle"] = rng.uniform(-1, 1, seq.shape).astype(np.float32); seq["steer"] = rng.uniform(-1, 1, seq.shape).astype(np.float32); seq["boost"] = 1
    host = ArenaBatch(T.MODE_SOCCAR, A, teams, lib=cudalib); host.reset_to_random_kickoff(seeds)
    host.step_with_controls(seq)
    dev = ArenaBatch(T.MODE_SOCCAR, A, teams, lib=cudalib); dev.reset_to_random_kickoff(seeds)
    stream = torch.cuda.Stream()
    dev.set_stream(stream.cuda_stream)
    d_seq = torch.from_numpy(seq.view(np.uint8).reshape(-1)).cuda()
    torch.cuda.synchronize()
    rlib.check(cudalib, cudalib.rsb_step_with_controls_device(dev.h, C.c_void_p(d_seq.data_ptr()), TK))
    dev.sync()
    assert pu.struct_equal(host.get_car_states(), dev.get_car_states()) and pu.struct_equal(host.get_ball_states(), dev.get_ball_states())
    # per-tick device controls == per-tick host controls
    one = ArenaBatch(T.MODE_SOCCAR, A, teams, lib=cudalib); one.reset_to_random_kickoff(seeds)
    for t in range(TK):
        d_c = torch.from_numpy(np.ascontiguousarray(seq[t]).view(np.uint8).reshape(-1)).cuda()
        torch.cuda.synchronize()
        rlib.check(cudalib, cudalib.rsb_set_controls_device(one.h, C.c_void_p(d_c.data_ptr())))
        one.step(1)
    assert pu.struct_equal(host.get_car_states(), one.get_car_states())
    for b in (host, dev, one):
        b.close()

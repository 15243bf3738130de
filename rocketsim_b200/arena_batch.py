"""``ArenaBatch``: the batched front end of the reference's ``Arena`` interface.

Mirrors the reference surface for the hot path (reference ``src/Sim/Arena/Arena.h``:
``Create``/``AddCar``/``Step``/``GetMutatorConfig``/``IsBallScored``/``tickCount``;
``Car::controls``/``GetState``/``SetState`` in ``src/Sim/Car/Car.h:149-152``;
``Ball::GetState``/``SetState`` in ``src/Sim/Ball/Ball.h:68-69``), with every call applying to
all arenas of the batch at once.  All compute happens in the CUDA library behind the C ABI
(``include/rsb_c_api.h``); this class only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import lib as _lib
from . import types as T


def default_arena_config(lib=None) -> np.ndarray:
    """``ArenaConfig()`` of the reference as a ``types.ArenaConfig`` record."""
    L = lib or _lib.load()
    out = np.zeros(1, dtype=T.ArenaConfig)
    L.rsb_default_arena_config(_lib.ptr(out))
    return out[0]


class ArenaBatch:
    def __init__(self, game_mode: int, n_arenas: int, teams: Sequence[int], car_configs: Optional[np.ndarray] = None,
                 mutator_config: Optional[np.ndarray] = None, tick_rate: float = 120.0, seed: int = 0, device: int = 0,
                 lib=None, arena_config: Optional[np.ndarray] = None, custom_boost_pads: Optional[np.ndarray] = None):
        """``Arena::Create(game_mode, arena_config, tick_rate)`` x n_arenas, then ``AddCar(team, config)`` for every entry of
        ``teams`` (car ids 1..n in that order).  ``arena_config``: a ``types.ArenaConfig`` record (default: the reference's
        defaults, see ``default_arena_config``); ``custom_boost_pads``: ``types.BoostPadConfig`` array, implies
        ``useCustomBoostPads`` (``ArenaConfig::customBoostPads``, ArenaConfig.h:43-46)."""
        self.L = lib or _lib.load()
        self.n_arenas = int(n_arenas)
        self.n_cars = len(teams)
        self.game_mode = int(game_mode)
        t = np.asarray(teams, dtype=np.int32)
        cfg = None if car_configs is None else np.ascontiguousarray(car_configs, dtype=T.CarConfig)
        if cfg is not None and len(cfg) != self.n_cars:
            raise ValueError("one CarConfig per car slot expected")
        mut = None if mutator_config is None else np.array([mutator_config], dtype=T.MutatorConfig)
        h = C.c_void_p()
        ac = None
        if arena_config is not None or custom_boost_pads is not None:
            ac = np.array([arena_config], dtype=T.ArenaConfig) if arena_config is not None else np.array([default_arena_config(self.L)], dtype=T.ArenaConfig)
            if custom_boost_pads is not None:
                self._pads = np.ascontiguousarray(custom_boost_pads, dtype=T.BoostPadConfig)
                ac["useCustomBoostPads"] = 1; ac["numCustomBoostPads"] = len(self._pads)
                ac["customBoostPads"] = self._pads.ctypes.data if len(self._pads) else 0
        _lib.check(self.L, self.L.rsb_batch_create_ex(device, game_mode, _lib.ptr(ac), n_arenas, self.n_cars, _lib.ptr(t) if self.n_cars else None,
                                                      _lib.ptr(cfg), _lib.ptr(mut), tick_rate, seed, C.byref(h)))
        self.h = h
        np_ = C.c_int()
        _lib.check(self.L, self.L.rsb_batch_info(self.h, None, None, C.byref(np_), None))
        self.n_pads = np_.value

    def close(self):
        if getattr(self, "h", None):
            self.L.rsb_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- controls / stepping
    def set_controls(self, controls: np.ndarray):
        c = np.ascontiguousarray(controls, dtype=T.CarControls).reshape(self.n_arenas, self.n_cars)
        _lib.check(self.L, self.L.rsb_set_controls(self.h, _lib.ptr(c)))

    def step(self, n_ticks: int = 1):
        _lib.check(self.L, self.L.rsb_step(self.h, n_ticks))

    def step_with_controls(self, controls_seq: np.ndarray):
        """``for t: car->controls = controls_seq[t]; Step(1)`` as one launch; controls_seq is
        ``[n_ticks, n_arenas, n_cars]`` host memory."""
        c = np.ascontiguousarray(controls_seq, dtype=T.CarControls).reshape(-1, self.n_arenas, self.n_cars)
        _lib.check(self.L, self.L.rsb_step_with_controls(self.h, _lib.ptr(c), c.shape[0]))

    def set_auto_reset(self, enable: bool, seed: int = 0):
        """Device-side ``if (IsBallScored()) ResetToRandomKickoff(seed + arena + tick)`` after every tick."""
        _lib.check(self.L, self.L.rsb_set_auto_reset(self.h, int(bool(enable)), seed & 0xFFFFFFFF))

    def set_arena_offset(self, global_index_of_arena0: int):
        """Sharded jobs: auto-reset seeds and event arena ids use ``offset + local arena index``."""
        _lib.check(self.L, self.L.rsb_set_arena_offset(self.h, int(global_index_of_arena0)))

    def set_event_tracker(self, enable: bool = True, config=None):
        """Device-side ``GameEventTracker`` (shots, goal credits, saves) updated after every tick; events arrive via get_events."""
        cfg = None if config is None else np.array([config], dtype=T.EventTrackerConfig)
        _lib.check(self.L, self.L.rsb_set_event_tracker(self.h, int(bool(enable)), _lib.ptr(cfg)))

    def is_ball_probably_going_in(self, max_time: float = 2.0, extra_margin: float = 0.0):
        """``Arena::IsBallProbablyGoingIn`` for every arena: (going_in[n_arenas] bool, goal_team[n_arenas])."""
        out = np.zeros(self.n_arenas, np.uint8); team = np.zeros(self.n_arenas, np.uint8)
        _lib.check(self.L, self.L.rsb_is_ball_probably_going_in(self.h, max_time, extra_margin, _lib.ptr(out), _lib.ptr(team)))
        return out.astype(bool), team

    def predict_ball(self, n_ticks: int):
        """Batched ``BallPredTracker``: ``[n_ticks, n_arenas]`` ball states, row 0 = now, row i = after i ticks of a ball-only
        arena started from the current ball state (``BallPredTracker::ForceUpdateAllPred``)."""
        out = np.zeros((int(n_ticks), self.n_arenas), dtype=T.BallState)
        _lib.check(self.L, self.L.rsb_predict_ball(self.h, int(n_ticks), _lib.ptr(out)))
        return out

    def clone_arena(self, src: int, dst: int):
        """``Arena::Clone``: arena ``dst`` becomes a clone of arena ``src``."""
        _lib.check(self.L, self.L.rsb_clone_arena(self.h, src, dst))

    def serialize_arena(self, arena: int) -> bytes:
        """``Arena::Serialize`` of one arena: bytes that RocketSim's ``Arena::DeserializeNew`` loads."""
        n = C.c_size_t()
        _lib.check(self.L, self.L.rsb_serialize_arena(self.h, arena, None, 0, C.byref(n)))
        buf = (C.c_uint8 * n.value)()
        _lib.check(self.L, self.L.rsb_serialize_arena(self.h, arena, buf, n.value, C.byref(n)))
        return bytes(buf)

    def deserialize_arena(self, arena: int, data: bytes):
        """``Arena::DeserializeNew`` into arena ``arena`` (mode, tick rate, car slots and mutators must match the batch)."""
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        _lib.check(self.L, self.L.rsb_deserialize_arena(self.h, arena, buf, len(data)))

    def sync(self):
        _lib.check(self.L, self.L.rsb_sync(self.h))

    def set_stream(self, cuda_stream: int):
        _lib.check(self.L, self.L.rsb_set_stream(self.h, C.c_void_p(cuda_stream)))

    # ---- state access
    def _get(self, fn, dtype, shape):
        out = np.zeros(shape, dtype=dtype)
        if out.size:
            _lib.check(self.L, getattr(self.L, fn)(self.h, _lib.ptr(out)))
        return out

    def get_car_states(self):
        return self._get("rsb_get_car_states", T.CarState, (self.n_arenas, self.n_cars))

    def get_ball_states(self):
        return self._get("rsb_get_ball_states", T.BallState, (self.n_arenas,))

    def get_pad_states(self):
        return self._get("rsb_get_pad_states", T.BoostPadState, (self.n_arenas, self.n_pads))

    def get_dropshot_tiles(self):
        """Arena::GetDropshotTilesState of every arena: damage states (0 full, 1 damaged, 2 broken), [n_arenas, 2, 70]."""
        return self._get("rsb_get_dropshot_tiles", np.uint8, (self.n_arenas, 2, 70))

    def get_car_internals(self):
        return self._get("rsb_get_car_internals", T.CarInternal, (self.n_arenas, self.n_cars))

    def get_ball_internals(self):
        return self._get("rsb_get_ball_internals", T.BallInternal, (self.n_arenas,))

    def get_tick_counts(self):
        return self._get("rsb_get_tick_counts", np.uint64, (self.n_arenas,))

    def get_error_flags(self):
        return self._get("rsb_get_error_flags", np.uint32, (self.n_arenas,))

    def is_ball_scored(self):
        return self._get("rsb_is_ball_scored", np.uint8, (self.n_arenas,)).astype(bool)

    def get_observations(self):
        nf = self.L.rsb_observation_floats(self.h)
        return self._get("rsb_get_observations", np.float32, (self.n_arenas, nf))

    def get_events(self, cap: int = 65536):
        parts = []
        while True:  # rc 1 = more events pending than `cap`: nothing is dropped, fetch the next slice
            out = np.zeros(cap, dtype=T.Event)
            n = C.c_int()
            rc = self.L.rsb_get_events(self.h, _lib.ptr(out), cap, C.byref(n))
            if rc not in (0, 1):
                _lib.check(self.L, rc)
            parts.append(out[: n.value])
            if rc == 0:
                break
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def _idx(self, arena_idx, car_idx=None):
        ai = np.ascontiguousarray(arena_idx, dtype=np.uint32).ravel()
        ci = None if car_idx is None else np.ascontiguousarray(car_idx, dtype=np.uint32).ravel()
        if ci is not None and len(ci) != len(ai):
            raise ValueError(f"arena_idx has {len(ai)} entries, car_idx {len(ci)}")
        return ai, ci

    @staticmethod
    def _payload(arr, n, what, per=1):
        """The C side reads n * per payload elements: refuse shorter arrays instead of letting it read past the buffer."""
        if arr.size != n * per:
            raise ValueError(f"{what}: expected {n * per} elements for {n} index entries, got {arr.size}")
        return arr

    def set_car_states(self, arena_idx, car_idx, states):
        ai, ci = self._idx(arena_idx, car_idx)
        s = self._payload(np.ascontiguousarray(states, dtype=T.CarState).ravel(), len(ai), "set_car_states")
        _lib.check(self.L, self.L.rsb_set_car_states(self.h, _lib.ptr(ai), _lib.ptr(ci), len(ai), _lib.ptr(s)))

    def set_car_internals(self, arena_idx, car_idx, internals):
        ai, ci = self._idx(arena_idx, car_idx)
        s = self._payload(np.ascontiguousarray(internals, dtype=T.CarInternal).ravel(), len(ai), "set_car_internals")
        _lib.check(self.L, self.L.rsb_set_car_internals(self.h, _lib.ptr(ai), _lib.ptr(ci), len(ai), _lib.ptr(s)))

    def set_ball_states(self, arena_idx, states):
        ai, _ = self._idx(arena_idx)
        s = self._payload(np.ascontiguousarray(states, dtype=T.BallState).ravel(), len(ai), "set_ball_states")
        _lib.check(self.L, self.L.rsb_set_ball_states(self.h, _lib.ptr(ai), len(ai), _lib.ptr(s)))

    def set_ball_internals(self, arena_idx, internals):
        ai, _ = self._idx(arena_idx)
        s = self._payload(np.ascontiguousarray(internals, dtype=T.BallInternal).ravel(), len(ai), "set_ball_internals")
        _lib.check(self.L, self.L.rsb_set_ball_internals(self.h, _lib.ptr(ai), len(ai), _lib.ptr(s)))

    def set_pad_states(self, arena_idx, states):
        ai, _ = self._idx(arena_idx)
        s = self._payload(np.ascontiguousarray(states, dtype=T.BoostPadState).ravel(), len(ai), "set_pad_states", self.n_pads)
        if self.n_pads:
            _lib.check(self.L, self.L.rsb_set_pad_states(self.h, _lib.ptr(ai), len(ai), _lib.ptr(s)))

    def set_dropshot_tiles(self, arena_idx, states):
        """Arena::SetDropshotTilesState for the listed arenas; states [len(arena_idx), 2, 70] of 0 / 1 / 2."""
        ai, _ = self._idx(arena_idx)
        s = self._payload(np.ascontiguousarray(states, dtype=np.uint8).ravel(), len(ai), "set_dropshot_tiles", 140)
        _lib.check(self.L, self.L.rsb_set_dropshot_tiles(self.h, _lib.ptr(ai), len(ai), _lib.ptr(s)))

    def seed_rng(self, arena: int, seed: int):
        _lib.check(self.L, self.L.rsb_seed_rng(self.h, arena, seed))

    def reset_to_random_kickoff(self, seeds, arena_idx=None):
        s = np.ascontiguousarray(seeds, dtype=np.int32).ravel()
        if arena_idx is None:
            if len(s) != self.n_arenas:
                raise ValueError("one seed per arena expected")
            _lib.check(self.L, self.L.rsb_reset_to_random_kickoff(self.h, None, self.n_arenas, _lib.ptr(s)))
        else:
            ai, _ = self._idx(arena_idx)
            self._payload(s, len(ai), "reset_to_random_kickoff seeds")
            _lib.check(self.L, self.L.rsb_reset_to_random_kickoff(self.h, _lib.ptr(ai), len(ai), _lib.ptr(s)))


class BallPredTracker:
    """Batched mirror of the reference's ``BallPredTracker`` (``src/Sim/BallPredTracker/BallPredTracker.h``): ``pred_data[i, a]`` is
    the ball of arena ``a`` after ``i`` ticks of a ball-only arena started from its current state.  The reference re-uses old
    predictions when the ball still matches them; on the GPU a full re-prediction of every arena is one launch, so
    ``update_pred_from_batch`` always does ``ForceUpdateAllPred``."""

    def __init__(self, batch: ArenaBatch, num_pred_ticks: int):
        self.batch = batch
        self.num_pred_ticks = int(num_pred_ticks)
        self.pred_data = None
        self.tick_time = None

    def update_pred_from_batch(self):
        self.pred_data = self.batch.predict_ball(self.num_pred_ticks)
        return self.pred_data

    def get_ball_state_for_time(self, pred_time: float, tick_rate: float = 120.0):
        """``GetBallStateForTime``: the predicted states ``[n_arenas]`` at ``pred_time`` seconds (index clamped to the horizon)."""
        if self.pred_data is None:
            raise RuntimeError("BallPredTracker::GetBallStateForTime(): Predicted ball data is empty, update prediction before calling")
        idx = int(min(max(pred_time * tick_rate, 0), self.num_pred_ticks - 1))
        return self.pred_data[idx]

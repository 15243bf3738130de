"""Multi-GPU = independent arena shards: rank r of W owns the contiguous global arena range
``shard_bounds(N, W, r)`` as its own :class:`ArenaBatch` on its own GPU, with its own copy of the static scene.
There is NO collective on the data path -- arenas never interact.  The only communication is the host-side gather of
results a caller may ask for (``gather_to_rank0``), which goes through ``torch.distributed`` objects (works with
the ``nccl`` and ``gloo`` backends alike) and is never inside a step.

Per-arena quantities that must not depend on the sharding (RNG seeds, kickoff seeds, bench control streams) are keyed
by the GLOBAL arena index, so a job gives bit-identical per-arena results for any world size.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .arena_batch import ArenaBatch


def shard_bounds(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split of ``n_total`` arenas; the first ``n_total % world_size`` ranks get one extra arena."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_device(rank: int) -> int:
    """CUDA device of a rank: LOCAL_RANK when the launcher sets it (torchrun), else rank modulo the GPUs of this node --
    never the global rank, which exceeds the device count on multi-node jobs."""
    import os
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:
        n = 0
    return rank % n if n > 0 else 0


class ShardedArenaBatch:
    """This rank's shard of a job of ``n_total`` arenas.  Construction, stepping and state access are purely local."""

    def __init__(self, game_mode: int, n_total: int, teams: Sequence[int], seed: int = 0, rank: int = 0, world_size: int = 1,
                 device: Optional[int] = None, lib=None, **kw):
        self.rank, self.world_size, self.n_total = rank, world_size, int(n_total)
        self.lo, self.hi = shard_bounds(n_total, world_size, rank)
        if self.hi <= self.lo:
            raise ValueError(f"rank {rank} of {world_size} owns no arena of {n_total}")
        # rsb_batch_create seeds arena a of the batch with seed + a: offsetting by lo makes the RNG stream a function of
        # the global arena index only
        self.batch = ArenaBatch(game_mode, self.hi - self.lo, teams, seed=seed + self.lo,
                                device=local_device(rank) if device is None else device, lib=lib, **kw)
        # auto-reset seeds and event arena ids are formed with the global arena index on the device
        self.batch.set_arena_offset(self.lo)

    @property
    def global_ids(self) -> np.ndarray:
        return np.arange(self.lo, self.hi)

    def reset_to_random_kickoff(self, seed: int):
        self.batch.reset_to_random_kickoff((seed + self.global_ids).astype(np.int32))

    def set_auto_reset(self, enable: bool, seed: int = 0):
        """Device-side kickoff reset after a goal, seeded with ``seed + GLOBAL arena index + tick``."""
        self.batch.set_auto_reset(enable, seed)

    def get_events(self, cap: int = 65536):
        """This shard's events; ``arena`` is the GLOBAL arena index."""
        return self.batch.get_events(cap)

    def local_slice(self, global_array: np.ndarray, axis: int = 0) -> np.ndarray:
        """The rows of a job-wide array (e.g. controls [n_total, n_cars]) that belong to this rank."""
        sl = [slice(None)] * global_array.ndim
        sl[axis] = slice(self.lo, self.hi)
        return global_array[tuple(sl)]

    def gather_to_rank0(self, local_array: np.ndarray):
        """Host-side gather of per-arena results, in global arena order, on rank 0 (None elsewhere)."""
        if self.world_size == 1:
            return local_array
        import torch.distributed as dist
        parts = [None] * self.world_size if self.rank == 0 else None
        dist.gather_object(local_array, parts, dst=0)
        return np.concatenate(parts, axis=0) if self.rank == 0 else None

    def close(self):
        self.batch.close()

#!/usr/bin/env python
"""bench.py -- arena-ticks/s of the batched Arena::Step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU Arena::Step on the host cores

Workload (BASELINE.json configs[1]): Soccar 2v2, 4,096 arenas per GPU, README bench inputs (each car
redraws its controls every 2..60 ticks, include/rsb_bench_inputs.h), kickoff reset on goal inside the
measured loop (device-side auto reset; the reference arm does `if IsBallScored() ResetToRandomKickoff`).
One "step" = TICKS_PER_STEP consecutive ticks of every arena = one launch of the fused tick kernel.

  value  : arena-ticks/s with the control sequences already resident in HBM (device timed, CUDA events).
  e2e    : same through the C ABI with HOST buffers: per step one H2D copy of the step's control
           sequence (pinned) + launch + D2H of every CarState and BallState.
  roofline: algorithmic bytes per arena-tick (B = 2*(nCars*308+108+nPads*12+8)+nCars*24)
           x arena-ticks per launch / measured kernel time, against the measured HBM peak.
  cpu_baseline: oracle/_ref (the unmodified reference) on all host cores, bounded sample.

Arenas are independent: N GPUs = N independent batches (weak scaling), no collective on the data path;
torch.distributed is used only for the barrier and the max-over-ranks of the timing.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODE_NAMES = {0: "soccar", 1: "hoops", 2: "heatseeker", 3: "snowday", 4: "dropshot", 5: "the_void"}
SEED = 1234


def algorithmic_bytes(n_cars: int, n_pads: int) -> int:
    return 2 * (n_cars * 308 + 108 + n_pads * 12 + 8) + n_cars * 24


def n_pads_of(mode: int) -> int:
    return {0: 34, 1: 20, 2: 34, 3: 34, 4: 0, 5: 0}[mode]


def ncu_traffic_per_launch(mode, cars_per_team, arenas, ticks):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE tick-kernel launch from the committed `ncu --set full` capture
    (profiles/rNN_summary.json, scripts/gpu_round.sh), valid only for the launch shape it was taken on; else None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not name.endswith("_summary.json"):
            continue
        try:
            d = json.load(open(os.path.join(pdir, name)))
            cfg = d["bench"]["config"]
            if (cfg["arenas_per_gpu"], cfg["ticks_per_step"], cfg["n_cars"]) == (arenas, ticks, 2 * cars_per_team) and \
                    cfg["workload"].startswith(MODE_NAMES[mode]):
                best = float(d["ncu_full"]["dram_traffic_bytes_per_launch"])
        except Exception:
            continue
    return best


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6500.0, "fallback (B200_PROFILING.md: ~6.5 TB/s measured copy bandwidth)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                out = [x.strip() for x in out]
                if len(out) >= 7:
                    self.sm.append(float(out[0]))
                    self.sm_max = float(out[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def workload_config(args, world):
    """`config` of the JSON line: identical for both arms (the driver compares them)."""
    n_cars = 2 * args.cars_per_team
    return {"workload": f"{MODE_NAMES[args.mode]} {args.cars_per_team}v{args.cars_per_team}, README bench inputs (redrawn every 2-60 ticks/car), "
                        f"{args.arenas} arenas/GPU, kickoff reset on goal", "arenas_per_gpu": args.arenas, "ticks_per_step": args.ticks_per_step,
            "n_cars": n_cars, "meshes": "synthetic .cmf (SURVEY F4)", "l2": "256 MiB flush write between timed steps",
            "parallelism": f"replicas x{world} (independent arena shards, no collective)"}


def cpu_reference_sample(args, steps, warmup):
    """The reference's CPU Arena::Step on all host cores, one arena per thread (README style, its fastest configuration): a short
    untimed warm-up, then a bounded sample of 512 x ticks_per_step ticks per step.  Used by BOTH the reference arm and the CUDA
    arm's cpu_baseline leg, so the two numbers of one box agree."""
    from oracle import refsim
    from rocketsim_b200 import synthetic_meshes as sm
    if not refsim.available():
        refsim.build()
    refsim.init_from_mem(sm.mesh_sets())
    threads = os.cpu_count() or 1
    apt = 1
    ticks = 512 * args.ticks_per_step
    if warmup > 0:
        refsim.bench(args.mode, args.cars_per_team, threads, apt, max(1, ticks * warmup // 8), SEED)
    v, wall = refsim.bench(args.mode, args.cars_per_team, threads, apt, ticks * max(1, steps), SEED)
    sample = f"{threads} threads x {apt} arena x {ticks * max(1, steps)} ticks ({wall:.1f} s), oracle/_ref"
    return v, wall, threads, sample


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref) on the host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # README style (README.md:10-11): one arena per thread is the reference's fastest configuration (more arenas per
    # thread only add cache pressure: measured 375k -> 28k arena-ticks/s from 1 to 64 arenas/thread on 8 cores).
    value, wall, threads, sample = cpu_reference_sample(args, args.steps, args.warmup)
    n_cars = 2 * args.cars_per_team
    line = {
        "impl": "reference", "metric": "arena_ticks_per_s", "value": value, "unit": "arena-ticks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "arena-ticks/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "arena-ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # the many-arenas-per-thread variant next to the README-style one (reported, never the headline)
    try:
        from oracle import refsim
        v64, w64 = refsim.bench(args.mode, args.cars_per_team, threads, 64, max(200, int(2.0 * value / (threads * 64))), SEED)
        line["cpu_baseline_64_arenas_per_thread"] = {"value": v64, "unit": "arena-ticks/s", "cores": threads, "wall_s": w64}
    except Exception as e:
        line["cpu_baseline_64_arenas_per_thread"] = {"value": None, "error": str(e)}
    print(json.dumps(line), flush=True)


def run_cuda(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (rocketsim_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from rocketsim_b200 import lib as rlib, synthetic_meshes as sm, types as T
    from rocketsim_b200.arena_batch import ArenaBatch

    L = rlib.load()  # rocketsim_b200/librsb.so, nothing else (developer A/B runs of other builds: scripts/ab_bench.py)
    rlib.init_from_mem(sm.mesh_sets(), L)
    n_cars = 2 * args.cars_per_team
    teams = [0] * args.cars_per_team + [1] * args.cars_per_team
    A, TPS = args.arenas, args.ticks_per_step
    batch = ArenaBatch(args.mode, A, teams, seed=SEED + rank * A, device=local_rank, lib=L)
    stream = torch.cuda.current_stream()
    batch.set_stream(stream.cuda_stream)
    # arena ids are global: rank r owns arenas [r*A, (r+1)*A) (contiguous shard)
    seeds = (SEED + rank * A + np.arange(A)).astype(np.int32)
    batch.reset_to_random_kickoff(seeds)
    batch.set_auto_reset(True, SEED + rank * A)

    n_steps_total = args.warmup + args.steps
    ctrl_bytes = TPS * A * n_cars * 24
    # pre-generated control sequences, resident in HBM, one per step (different ticks -> different data)
    d_seq = [torch.empty(ctrl_bytes, dtype=torch.uint8, device="cuda") for _ in range(n_steps_total)]
    for i, buf in enumerate(d_seq):
        rlib.check(L, L.rsb_bench_fill_controls_device(batch.h, C.c_void_p(buf.data_ptr()), TPS, i * TPS, SEED + rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        rlib.check(L, L.rsb_step_with_controls_device(batch.h, C.c_void_p(d_seq[i].data_ptr()), TPS))

    # ---- device-resident arm ("value")
    for i in range(args.warmup):
        flush.zero_()  # also warms torch's fill kernel: no lazy module load inside the timed region
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record(stream)
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        kev[k][0].record(stream)
        step_device(args.warmup + k)
        kev[k][1].record(stream)
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
    err = batch.get_error_flags()
    ticks_done = batch.get_tick_counts()
    assert int(ticks_done.min()) == n_steps_total * TPS, "tick counter mismatch: work was skipped"

    # ---- end-to-end arm through the C ABI with host buffers.  A FRESH batch with the same seeds, warmed up the same way,
    # so that both arms cover the same tick window [W*TPS, (W+K)*TPS) of the same games (the cost of a tick depends on how
    # far the games have developed).
    h_seq = [torch.empty(ctrl_bytes, dtype=torch.uint8).pin_memory() for _ in range(n_steps_total)]
    for k in range(n_steps_total):
        h_seq[k].copy_(d_seq[k])
    batch.close()
    batch = ArenaBatch(args.mode, A, teams, seed=SEED + rank * A, device=local_rank, lib=L)
    batch.set_stream(stream.cuda_stream)
    batch.reset_to_random_kickoff(seeds)
    batch.set_auto_reset(True, SEED + rank * A)
    # result buffers in PINNED host memory (what a training loop would hand in): numpy views over torch pinned tensors
    car_pin = torch.empty(A * n_cars * T.CarState.itemsize, dtype=torch.uint8).pin_memory()
    ball_pin = torch.empty(A * T.BallState.itemsize, dtype=torch.uint8).pin_memory()
    car_out = car_pin.numpy().view(T.CarState).reshape(A, n_cars)
    ball_out = ball_pin.numpy().view(T.BallState)
    d2h_bytes = car_out.nbytes + ball_out.nbytes

    def step_e2e(k):
        rlib.check(L, L.rsb_step_with_controls(batch.h, C.c_void_p(h_seq[k].data_ptr()), TPS))
        rlib.check(L, L.rsb_get_car_states(batch.h, rlib.ptr(car_out)))
        rlib.check(L, L.rsb_get_ball_states(batch.h, rlib.ptr(ball_out)))

    for k in range(args.warmup):
        step_e2e(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_host0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(args.warmup + k)
    e1.record(stream)
    barrier()
    t_host1 = time.perf_counter()  # after the barrier: every copy has landed
    e2e_ms = max(e0.elapsed_time(e1), (t_host1 - t_host0) * 1e3)  # the slower of the device clock and the host wall clock
    checksum = float(np.abs(car_out["pos"]["x"]).sum())

    # ---- long window: the e2e batch simply keeps playing, device resident, for --long-window-ticks more ticks
    long_ms = None
    if args.long_window_ticks > 0:
        nst = max(1, args.long_window_ticks // TPS)
        args.long_window_ticks = nst * TPS
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for k in range(nst):
            buf = d_seq[k & 1]
            rlib.check(L, L.rsb_bench_fill_controls_device(batch.h, C.c_void_p(buf.data_ptr()), TPS, (n_steps_total + k) * TPS, SEED + rank))
            rlib.check(L, L.rsb_step_with_controls_device(batch.h, C.c_void_p(buf.data_ptr()), TPS))
        l1.record(stream)
        barrier()
        long_ms = l0.elapsed_time(l1)
        err = np.maximum(err, batch.get_error_flags())

    t = torch.tensor([total_ms, e2e_ms, float(np.mean(kern_ms)), long_ms if long_ms is not None else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_mean, long_max = [float(x) for x in t.cpu()]
    long_ms = long_max if long_ms is not None else None
    errs = torch.tensor([int(err.max())], device="cuda")
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)

    if rank == 0:
        arena_ticks_total = world * A * TPS * args.steps
        value = arena_ticks_total / (total_ms * 1e-3)
        e2e = arena_ticks_total / (e2e_ms * 1e-3)
        n_pads = n_pads_of(args.mode)
        Bt = algorithmic_bytes(n_cars, n_pads)
        peak, peak_src = measured_hbm_peak()
        achieved = A * TPS * Bt / (kern_ms_mean * 1e-3) / 1e9
        line = {
            "metric": "arena_ticks_per_s", "value": value, "unit": "arena-ticks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e, "unit": "arena-ticks/s", "h2d_bytes_per_step": ctrl_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps, "launches_per_step": 3 + 2,
                    "path": "rsb_step_with_controls (host control sequence, pinned) + rsb_get_car_states + rsb_get_ball_states (pinned)"},
            "gpu_launches": args.steps,  # one arena_step_kernel launch per step (plus one memset for the L2 flush)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(args.mode, args.cars_per_team, A, TPS), "traffic_unit": "bytes/launch (ncu, profiles/)",
                         "kernel": "arena_step_kernel", "kernel_ms": kern_ms_mean,
                         "bytes_per_arena_tick": Bt, "peak_source": peak_src},
            "clocks": sampler.result(),
            "device_error_flags": int(errs.item()), "checksum": checksum,
        }
        if long_ms is not None:
            line["long_window"] = {"ticks": args.long_window_ticks, "value": world * A * args.long_window_ticks / (long_ms * 1e-3), "unit": "arena-ticks/s",
                                   "note": "device-resident arm continued for a long window (games fully developed); control sequences "
                                           "generated on the device per step inside the timed region"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, wall, cores, sample = cpu_reference_sample(args, 5, 1)
                line["cpu_baseline"] = {"value": v, "unit": "arena-ticks/s", "cores": cores, "kind": "reference", "sample": sample}
            except Exception as e:  # the checker must never break the measurement
                line["cpu_baseline"] = {"value": None, "unit": "arena-ticks/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--cars-per-team", type=int, default=2)
    ap.add_argument("--arenas", type=int, default=4096, help="arenas per GPU")
    ap.add_argument("--ticks-per-step", type=int, default=120)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--long-window-ticks", type=int, default=36000, help="extra device-resident window after the timed steps (0 = off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

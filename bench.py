#!/usr/bin/env python
"""bench.py -- segment x energy-group intersections/s of the attenuation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference CPU arm

One "step" = one sweep of the hot path over this rank's shard of the deterministic segment
stream: config 2 of BASELINE.json (128 energy groups, 1e8 segments, 6750 regions x 5 intervals,
100 segments per track) per GPU, i.e. weak scaling over the global stream of N x 1e8 segments,
followed for N > 1 by the one NCCL all-reduce of the tally deltas (north star item 4).

Timing: per-step CUDA events on the launching stream (the library is switched onto torch's
current stream), max over ranks, L2 flushed between steps.  `value` counts inputs resident in
HBM; `e2e` is the same sweep through the host-buffer C-ABI call (pinned host slabs, H2D + D2H
inside the timed region).  The reference arm times the UNMODIFIED reference CPU run_kernel
(oracle/_ref/libref_ofast.so = /root/reference/src/cpu built with its Makefile's gnu flags) on
all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "segment\u00d7energy-group intersections/sec"   # BASELINE.json metric
UNIT = "intersections/s"
# 10.4 B source rows (2.6 rows avg) + 4 B sigT + 4 B tally RED payload
ALGO_BYTES_PER_INTERSECTION = 18.4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--segments", type=int, default=100_000_000, help="segments per GPU per step")
    ap.add_argument("--egroups", type=int, default=128)
    ap.add_argument("--regions-2d", type=int, default=5000)
    ap.add_argument("--seg-per-track", type=int, default=100)
    ap.add_argument("--exp", default="poly", choices=["poly", "mufu", "glibc", "table"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-segments", type=int, default=50_000_000,
                    help="reference arm: segments per step; default = the reference's own README default run "
                         "(init.c:12: 5e7 segments x 128 groups), BASELINE config 1")
    return ap.parse_args()


def workload_name(a):
    regions = -(-a.regions_2d * 27 // 20)
    return (f"BASELINE config 2: {a.egroups} energy groups, {a.segments:.0e} segments/GPU, "
            f"{regions} regions x 5 fine axial intervals, {a.seg_per_track} segments/track")


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------------------
# clocks during the timed region (pynvml; the recipe's nvidia-smi line as a fallback)
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference CPU run_kernel on the host cores
# --------------------------------------------------------------------------------------
def time_reference(a, steps, warmup, segments, variant="ofast"):
    from oracle.oracle import Oracle, Reference
    groups = a.egroups
    if Reference.available(variant):
        ref = Reference(variant)
        cores = ref.num_procs()
        kind = "reference"
        run = lambda: ref.time_run_kernel(a.regions_2d, groups, segments, cores)  # noqa: E731
        how = ("unmodified /root/reference/src/cpu run_kernel, Makefile gnu flags (-Ofast -msse2 -fopenmp)"
               if variant == "ofast" else f"unmodified /root/reference/src/cpu run_kernel, {variant} build")
    else:  # the reference could not be compiled where this repo was built: time the oracle port
        o = Oracle()
        cores = o.max_threads()
        kind = "port"
        regions = -(-a.regions_2d * 27 // 20)
        src, flux, sig = o.fill(regions, 5, groups, a.seed)

        def run():
            t0 = time.perf_counter()
            o.run(src, flux, sig, segments, a.seg_per_track, a.seed, nthreads=cores)
            return time.perf_counter() - t0
        how = "oracle/smk_oracle.c port (-O2 -ffp-contract=off -fopenmp)"
    for _ in range(warmup):
        run()
    times = [run() for _ in range(steps)]
    total = sum(times)
    value = steps * segments * groups / total
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} x {segments} segments x {groups} groups ({how})",
            "ms_per_step": 1e3 * total / steps}


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_reference(a, a.steps, max(a.warmup, 1), a.ref_segments)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "timed_sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def main_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import smk_b200 as smk
    multi = smk.multi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G = a.egroups
    I = smk.Input(source_2D_regions=a.regions_2d, segments=a.segments * world, egroups=G,
                  seg_per_thread=a.seg_per_track, seed=a.seed, exp_mode=a.exp, math_mode=a.math,
                  device=local_rank).finalize()
    R, F = I.source_3D_regions, I.fine_axial_intervals
    nt = I.n_tracks
    tb, te = multi.shard_tracks(nt, rank, world)
    my_segments = multi.shard_segments(nt, a.seg_per_track, I.segments, rank, world)

    ctx = smk.Context(I)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.fill_device(0.0)
    tally = torch.as_tensor(multi.DevicePointer(ctx.tally_ptr, ctx.padded_elems), device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def step():
        ctx.reset_tallies()
        ctx.run_async(tb, te)
        if world > 1:
            multi.all_reduce_tallies(tally)  # tally deltas, once per sweep, over NVLink

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(a.warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sampler.start()
    barrier()
    for k in range(a.steps):
        flush.zero_()                        # evict the 38 MB working set from L2 between steps
        ev[k][0].record(stream)
        ctx.reset_tallies()
        kev[k][0].record(stream)
        ctx.run_async(tb, te)
        kev[k][1].record(stream)
        if world > 1:
            multi.all_reduce_tallies(tally)
        ev[k][1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0

    step_ms = [s.elapsed_time(e) for s, e in ev]
    kern_ms = [s.elapsed_time(e) for s, e in kev]
    t = torch.tensor([sum(step_ms), sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_total_ms = t.tolist()
    total_intersections = float(I.segments) * G * a.steps
    value = total_intersections / (total_ms * 1e-3)

    # roofline of the dominant kernel (attenuate_tracks): algorithmic bytes / launch duration
    kernel_ms = kernel_total_ms / a.steps
    peak, peak_kind = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_INTERSECTION * my_segments * G / (kernel_ms * 1e-3) / 1e9
    # secondary roofline: the FP32 pipe, which is what physically binds on the L2-resident working set
    # (DESIGN.md section 5.2).  FAST/POLY issues 46 FP32 lane-operations per interior intersection and
    # 30 per edge intersection (csrc/smk_math.cuh), i.e. (46 (F-2) + 30 * 2) / F on average.
    lane_ops = (46.0 * (F - 2) + 30.0 * 2) / F
    sm_hz = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
    fp32_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 128 * sm_hz
    fp32_achieved = lane_ops * my_segments * G / (kernel_ms * 1e-3)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- end to end through the host-buffer API: pinned slabs, H2D + sweep + D2H per step -----
    src = smk.alloc_pinned((R, F, G))
    flux0 = smk.alloc_pinned((R, F, G))
    sig = smk.alloc_pinned((R, G))
    out = smk.alloc_pinned((R, F, G))
    rng = np.random.default_rng(a.seed)          # every rank holds the same replica of the slabs
    src[...] = rng.random(src.shape, dtype=np.float32)
    flux0[...] = rng.random(flux0.shape, dtype=np.float32)
    sig[...] = rng.random(sig.shape, dtype=np.float32)

    def e2e_step():
        ctx.upload(src, flux0, sig)          # H2D (also zeroes the tally deltas)
        ctx.run_async(tb, te)
        if world > 1:
            multi.all_reduce_tallies(tally)
        ctx.download_flux(out)               # D2H of flux0 + tallies (synchronises)

    e2e_step()
    barrier()
    e2e_steps = max(2, min(a.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = float(I.segments) * G * e2e_steps / t.item()
    h2d = src.nbytes + flux0.nbytes + sig.nbytes
    d2h = out.nbytes

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = time_reference(a, steps=2, warmup=1, segments=min(a.ref_segments, 20_000_000))
        cpu.pop("ms_per_step", None)
        try:   # the same sources built for AVX2+FMA (-march=x86-64-v3): the "fair" CPU figure
            from oracle.oracle import Reference
            if Reference.available("v3"):
                cpu["alt_avx2_fma_build"] = time_reference(a, steps=1, warmup=1, segments=min(a.ref_segments, 20_000_000),
                                                           variant="v3")["value"]
        except Exception:
            pass

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "segments_global": I.segments, "egroups": G,
                       "exp_mode": a.exp, "math_mode": a.math, "seed": a.seed,
                       "l2": "flushed between steps (256 MB memset); the 38 MB working set is "
                             "L2-resident within a step by construction of the workload",
                       "sharding": f"tracks split over {world} rank(s); one all-reduce of tally deltas per step"
                       if world > 1 else "single GPU"},
            "ns_per_intersection": 1e9 / value,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "kernel": "attenuate_tracks", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_intersection": ALGO_BYTES_PER_INTERSECTION},
            "compute_roofline": {"bound": "fp32_pipe", "achieved": fp32_achieved / 1e12, "peak": fp32_peak / 1e12,
                                 "unit": "T lane-op/s", "frac": fp32_achieved / fp32_peak,
                                 "lane_ops_per_intersection": lane_ops,
                                 "note": "valid for --math fast --exp poly; peak = SMs x 128 lanes x SM clock under load"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)

    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)

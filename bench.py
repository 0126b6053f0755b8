#!/usr/bin/env python
"""bench.py -- segment x energy-group intersections/s of the attenuation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference CPU arm

Workload (BASELINE.json configs; `config.workload` names it):
  N = 1   config 2: 128 energy groups, 1e8 segments, 6750 regions x 5 intervals, 100 segments per track.
          One step = one sweep of the whole deterministic stream.  Short extra legs (3 timed steps each,
          under `configs`) carry config 3 (7 groups), config 4 (64 groups, 14 regions), the HBM-resident
          regime (432 000 regions), per-segment geometry and config 5's 1e10 segments on one GPU.
  N > 1   config 5 as written: 1e10 segments GLOBAL, 128 groups, tracks sharded by contiguous range over
          the N ranks, ONE all-reduce (NCCL over NVLink) of the tally deltas per step ("scaling":
          "strong").  The weak-scaling figure (1e8 segments per GPU) is the `weak` sub-record, and
          `single_process` is the same weak workload driven by ONE process through smk_multi_* with the
          library's own peer-memory all-reduce kernel.

Timing: per-step CUDA events on the launching stream (the library is switched onto torch's current
stream), max over ranks, L2 flushed between steps.  `value` counts inputs resident in HBM.  `e2e` is the
same sweep through the host-buffer API: pinned host slabs, H2D + sweep + D2H inside the timed region,
two contexts in flight so that the copies of step k+1 overlap the sweep of step k; at N > 1 every rank
uploads / downloads 1/N of the rows and the replicas are completed over NVLink.
Roofline: the BINDING resource per leg -- the FP32 pipe for the L2-resident working sets (38 MB of source
data live in the 126 MB L2, measured DRAM traffic per launch is under `traffic`), HBM for the HBM-resident
regime.  The reference arm times the UNMODIFIED reference CPU run_kernel (oracle/_ref/libref_ofast.so =
/root/reference/src/cpu built with its Makefile's gnu flags) on all host cores; the config 3 and config 4 legs
carry the same reference timed at their own shape (`cpu_reference`).
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

METRIC = "segment×energy-group intersections/sec"   # BASELINE.json metric
UNIT = "intersections/s"
# 10.4 B source rows (2.6 rows avg) + 4 B sigT + 4 B tally RED payload
ALGO_BYTES_PER_INTERSECTION = 18.4
CONFIG5_SEGMENTS = 10_000_000_000
CPU_SAMPLE_SEGMENTS = 100_000_000          # reference arm: at most this many segments per timed step


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--segments", type=int, default=None,
                    help="GLOBAL segments per step (default: 1e8 at N=1 = config 2, 1e10 at N>1 = config 5)")
    ap.add_argument("--egroups", type=int, default=128)
    ap.add_argument("--regions-2d", type=int, default=5000)
    ap.add_argument("--seg-per-track", type=int, default=100)
    ap.add_argument("--exp", default="poly", choices=["poly", "mufu", "glibc", "table"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--geometry", action="store_true", help="per-segment geometry (SMK_FLAG_SEGMENT_GEOMETRY)")
    ap.add_argument("--fit-per-sweep", action="store_true",
                    help="SMK_FLAG_FIT_PER_SWEEP: the axial source fit once per (region, interval, group) per sweep "
                         "instead of once per segment (off by default, as in the reference; <= 64 groups)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra per-config legs / sub-records")
    ap.add_argument("--ref-segments", type=int, default=None,
                    help="reference arm: segments per timed step (default = --segments, capped at 1e8)")
    a = ap.parse_args()
    if a.segments is None:
        a.segments = 100_000_000 if a.gpus == 1 else CONFIG5_SEGMENTS
    return a


def regions_3d(regions_2d):
    return -(-regions_2d * 27 // 20)          # main.c:18-19


def workload_name(egroups, segments, regions_2d, seg_per_track, gpus, geometry=False):
    R = regions_3d(regions_2d)
    if egroups == 128 and regions_2d == 5000 and segments == CONFIG5_SEGMENTS:
        tag = "BASELINE config 5"
    elif egroups == 128 and regions_2d == 5000 and segments == 100_000_000 * gpus:
        tag = "BASELINE config 2" if gpus == 1 else "BASELINE config 2 per GPU (weak scaling)"
    elif egroups == 7 and regions_2d == 5000:
        tag = "BASELINE config 3"
    elif egroups == 64 and R <= 64:
        tag = "BASELINE config 4"
    elif R * 5 * egroups * 4 * 3 > 512e6:
        tag = "HBM-resident regime"
    else:
        tag = "custom"
    return (f"{tag}: {egroups} energy groups, {segments:.0e} segments global over {gpus} GPU(s), "
            f"{R} regions x 5 fine axial intervals, {seg_per_track} segments/track"
            + (", per-segment geometry" if geometry else ""))


def config_dict(a):
    """Identical for both arms (the driver compares them)."""
    return {"workload": workload_name(a.egroups, a.segments, a.regions_2d, a.seg_per_track, a.gpus, a.geometry),
            "segments_global": a.segments, "egroups": a.egroups, "regions_2d": a.regions_2d,
            "seg_per_track": a.seg_per_track, "exp_mode": a.exp, "math_mode": a.math, "seed": a.seed,
            "fit_per_sweep": bool(getattr(a, "fit_per_sweep", False)),
            "l2": "flushed between steps (256 MB memset); within a step the 38 MB working set of the "
                  "reference's default geometry is L2-resident by construction of the workload",
            "sharding": (f"tracks split over {a.gpus} ranks by contiguous range; one all-reduce of tally deltas per step"
                         if a.gpus > 1 else "single GPU")}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def recorded_traffic(key):
    """DRAM bytes per launch of this round's ncu --set full capture (profiles/traffic_r02.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r02.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


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
        src, flux, sig = o.fill(regions_3d(a.regions_2d), 5, groups, a.seed)

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
    # a step = the arm's own workload, bounded to 1e8 segments (config 5's 1e10 would take ~7 minutes per
    # step on 16 cores); the metric is a per-intersection rate, so the sample size does not enter it
    segments = a.ref_segments if a.ref_segments is not None else min(a.segments, CPU_SAMPLE_SEGMENTS)
    r = time_reference(a, a.steps, max(a.warmup, 1), segments)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak" if a.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def lane_ops_per_intersection(smk, egroups, F, geometry, fit_per_sweep=False):
    """FP32 lane-operations of the FAST/POLY arithmetic (csrc/smk_math.cuh, counted in the SASS of the
    default kernels: FFMA2 + FMUL2 + FADD2 per pair of groups): 45 per interior and 29 per edge
    intersection where the segment type is warp-uniform (33..128 groups: one track per warp), 45 for every
    intersection where tracks of different types share a warp (<= 32 groups) or rows are swept in blocks;
    one more with per-segment geometry (the weight is applied per intersection instead of once at the end).
    With --fit-per-sweep (SMK_FLAG_FIT_PER_SWEEP, <= 128 groups) the fit's 8 (interior) / 3 (edge) operations are
    evaluated per (region, interval, group) and sweep instead: 37 / 26 remain in the segment loop."""
    extra = 1.0 if geometry else 0.0
    gp = smk.lib.smk_padded_groups(egroups)
    hoist = fit_per_sweep and gp <= 128 and not geometry
    interior, edge = (37.0, 26.0) if hoist else (45.0, 29.0)
    if gp >= 64:
        return ((interior + extra) * (F - 2) + (edge + extra) * 2) / F
    return interior + extra


class Sweep:
    """One device-resident problem on this rank + the timed sweep loop."""

    def __init__(self, torch, dist, smk, dev, rank, world, *, egroups, segments, regions_2d, seg_per_track, seed,
                 exp, math, geometry=False, fit_per_sweep=False):
        self.torch, self.dist, self.smk, self.dev, self.rank, self.world = torch, dist, smk, dev, rank, world
        self.G = egroups
        self.I = smk.Input(source_2D_regions=regions_2d, segments=segments, egroups=egroups,
                           seg_per_thread=seg_per_track, seed=seed, exp_mode=exp, math_mode=math,
                           device=dev.index, segment_geometry=geometry, fit_per_sweep=fit_per_sweep).finalize()
        self.ctx = smk.Context(self.I)
        self.stream = torch.cuda.current_stream(dev)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ctx.fill_device(0.0)
        nt = self.I.n_tracks
        self.tb, self.te = smk.multi.shard_tracks(nt, rank, world)
        self.my_segments = smk.multi.shard_segments(nt, seg_per_track, segments, rank, world)
        self.tally = (torch.as_tensor(smk.multi.DevicePointer(self.ctx.tally_ptr, self.ctx.padded_elems), device=dev)
                      if world > 1 else None)

    def step(self):
        self.ctx.reset_tallies()
        self.ctx.run_async(self.tb, self.te)
        if self.world > 1:
            self.smk.multi.all_reduce_tallies(self.tally)     # tally deltas, once per sweep, over NVLink

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, steps, warmup, flush, sampler=None):
        """Returns (total_ms, kernel_total_ms, launches), max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        self.barrier()
        launches0 = self.ctx.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if sampler:
            sampler.start()
        self.barrier()
        for k in range(steps):
            flush.zero_()                        # evict the working set from L2 between steps
            ev[k][0].record(self.stream)
            self.ctx.reset_tallies()
            kev[k][0].record(self.stream)
            self.ctx.run_async(self.tb, self.te)
            kev[k][1].record(self.stream)
            if self.world > 1:
                self.smk.multi.all_reduce_tallies(self.tally)
            ev[k][1].record(self.stream)
        self.barrier()
        launches = self.ctx.launch_count - launches0
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in ev), sum(s.elapsed_time(e) for s, e in kev)],
                         dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        total_ms, kernel_ms = t.tolist()
        return total_ms, kernel_ms, launches

    def rooflines(self, kernel_ms_per_step, clocks, hbm_resident):
        """The binding roofline of the attenuation kernel for this leg + the other one for reference."""
        torch = self.torch
        F = self.I.fine_axial_intervals
        peak, peak_kind = measured_peaks()
        inter = float(self.my_segments) * self.G
        sec = kernel_ms_per_step * 1e-3
        algo_gbs = ALGO_BYTES_PER_INTERSECTION * inter / sec / 1e9
        lane_ops = lane_ops_per_intersection(self.smk, self.G, F, self.I.segment_geometry, self.I.fit_per_sweep)
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        fp32_peak = sms * 128 * sm_mhz * 1e6
        fp32 = {"bound": "fp32_pipe", "achieved": lane_ops * inter / sec / 1e12, "peak": fp32_peak / 1e12,
                "unit": "T lane-op/s", "frac": lane_ops * inter / sec / fp32_peak,
                "lane_ops_per_intersection": lane_ops, "sm_mhz_under_load": sm_mhz,
                "peak_kind": f"{sms} SMs x 128 FP32 lanes x SM clock under load",
                "note": "lane-op count valid for --math fast --exp poly (csrc/smk_math.cuh)"}
        hbm = {"bound": "hbm", "achieved": algo_gbs, "peak": peak, "unit": "GB/s", "frac": algo_gbs / peak,
               "peak_kind": peak_kind, "algorithmic_bytes_per_intersection": ALGO_BYTES_PER_INTERSECTION}
        return (hbm, fp32) if hbm_resident else (fp32, hbm)

    def close(self):
        self.ctx.close()


def run_leg(torch, dist, smk, dev, rank, world, flush, a, name, *, steps=3, warmup=3, hbm_resident=False, **kw):
    """One short timed leg for another configuration; returns its record (rank 0) or None."""
    base = dict(egroups=128, segments=100_000_000 * world, regions_2d=5000, seg_per_track=a.seg_per_track,
                seed=a.seed, exp=a.exp, math=a.math, geometry=False, fit_per_sweep=False)
    base.update(kw)
    sw = Sweep(torch, dist, smk, dev, rank, world, **base)
    sampler = ClockSampler(dev.index)
    total_ms, kernel_ms, _ = sw.timed(steps, warmup, flush, sampler)
    clocks = sampler.stop()
    value = float(sw.I.segments) * sw.G * steps / (total_ms * 1e-3)
    binding, other = sw.rooflines(kernel_ms / steps, clocks, hbm_resident)
    rec = {"name": name, "workload": workload_name(sw.G, sw.I.segments, base["regions_2d"], base["seg_per_track"],
                                                   world, base["geometry"]),
           "value": value, "unit": UNIT, "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
           "arithmetic": ("source fit evaluated once per (region, interval, group) per sweep (SMK_FLAG_FIT_PER_SWEEP): "
                          "NOT the reference's per-segment evaluation; results bit-identical" if base["fit_per_sweep"]
                          else "as the reference: every operation of attenuate_segment per segment"),
           "kernel": sw.ctx.kernel_name, "roofline": binding,
           ("compute_roofline" if hbm_resident else "hbm_algorithmic"): other,
           "clocks": {"sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"]}}
    # measured DRAM bytes of one launch of this leg's shape (ncu --set full of this round), where recorded
    t = recorded_traffic({"config3_7_groups": "config3_7_groups", "config4_64_groups_14_regions": "config4_64_groups_14_regions",
                          "hbm_resident_432000_regions": "hbm_resident"}.get(name, ""))
    if t and world == 1:
        binding["traffic"], binding["traffic_source"] = t["dram_bytes_per_launch"], t["source"]
    sw.close()
    return rec if rank == 0 else None


def e2e_run(torch, dist, smk, dev, rank, world, a, steps):
    """End to end through the host-buffer API, two contexts in flight.  Returns (intersections/s over all
    ranks, h2d bytes per step on this rank, d2h bytes per step on this rank)."""
    import numpy as np
    multi = smk.multi
    G = a.egroups
    segments = 100_000_000 * world if a.segments == CONFIG5_SEGMENTS else a.segments   # config 2 (per GPU)
    I = smk.Input(source_2D_regions=a.regions_2d, segments=segments, egroups=G, seg_per_thread=a.seg_per_track,
                  seed=a.seed, exp_mode=a.exp, math_mode=a.math, device=dev.index,
                  segment_geometry=a.geometry, fit_per_sweep=a.fit_per_sweep).finalize()
    R, F = I.source_3D_regions, I.fine_axial_intervals
    rows = R * F
    tb, te = multi.shard_tracks(I.n_tracks, rank, world)
    rb, re_ = multi.shard_rows(rows, rank, world)          # rows of source / flux this rank moves over PCIe
    qb, qe = multi.shard_rows(R, rank, world)              # rows of sigT

    # host side: the caller's slabs in pinned memory (every rank holds the same synthetic replica, a rank
    # only ever touches its own rows of it)
    src = smk.alloc_pinned((rows, G))
    flux0 = smk.alloc_pinned((rows, G))
    sig = smk.alloc_pinned((R, G))
    rng = np.random.default_rng(a.seed)
    src[...] = rng.random(src.shape, dtype=np.float32)
    flux0[...] = rng.random(flux0.shape, dtype=np.float32)
    sig[...] = rng.random(sig.shape, dtype=np.float32)
    sig_bound = float(sig.max())                           # = all-reduce(max) of the ranks' slice maxima
    outs = [smk.alloc_pinned((rows, G)) for _ in range(2)]

    lanes = []
    for _ in range(2):
        st = torch.cuda.Stream(device=dev)
        ctx = smk.Context(I)
        ctx.set_stream(st.cuda_stream)
        Gp = ctx.G_pad
        t_tally = torch.as_tensor(multi.DevicePointer(ctx.tally_ptr, ctx.padded_elems), device=dev)
        t_src = torch.as_tensor(multi.DevicePointer(ctx.source_ptr, rows * Gp), device=dev)
        t_sig = torch.as_tensor(multi.DevicePointer(ctx.sigt_ptr, R * Gp), device=dev)
        lanes.append((st, ctx, t_tally, t_src, t_sig, Gp))
    sliced = world > 1 and lanes[0][1].padded_elems == rows * lanes[0][5]      # no tally replicas

    def enqueue(k):
        st, ctx, t_tally, t_src, t_sig, Gp = lanes[k % 2]
        out = outs[k % 2]
        with torch.cuda.stream(st):
            # (wait_finalized: the other lane's `flux0 + tallies` pass goes in front of this sweep -- a persistent grid
            # that would otherwise starve it, its download and the next upload for a whole sweep, tools/e2e_timeline.py;
            # the H2D copies above it still overlap the other lane's sweep)
            other = lanes[(k + 1) % 2][1]
            if world == 1:
                ctx.upload_async(src, flux0, sig)                  # H2D (also zeroes the tally deltas)
                ctx.wait_finalized(other)
                ctx.run_async(tb, te)
                ctx.download_flux_rows_async(0, rows, out)         # D2H of flux0 + tallies
            else:
                ctx.upload_rows_async(smk.ARRAY_SOURCE, rb, re_ - rb, src[rb:re_])
                ctx.upload_rows_async(smk.ARRAY_SIGT, qb, qe - qb, sig[qb:qe])
                ctx.upload_rows_async(smk.ARRAY_FLUX, rb, re_ - rb, flux0[rb:re_])
                multi.gather_row_slices(t_src, rows, Gp, world)    # complete the replicas over NVLink
                multi.gather_row_slices(t_sig, R, Gp, world)
                ctx.set_sigt_bound(sig_bound)
                ctx.reset_tallies()
                ctx.wait_finalized(other)
                ctx.run_async(tb, te)
                if sliced:
                    multi.reduce_row_slices(t_tally, rows, Gp, world)
                else:
                    multi.all_reduce_tallies(t_tally)
                ctx.download_flux_rows_async(rb, re_ - rb, out[rb:re_])

    def drain(k):
        lanes[k % 2][0].synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for k in range(2):                 # warm both lanes
        enqueue(k)
        drain(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        enqueue(k)
        if k >= 1:
            drain(k - 1)               # the result of step k-1 is on the host; its lane is free for step k+1
    drain(steps - 1)
    barrier()
    sec = time.perf_counter() - t0
    t = torch.tensor([sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = float(I.segments) * G * steps / t.item()
    if world == 1:
        h2d, d2h = src.nbytes + flux0.nbytes + sig.nbytes, outs[0].nbytes
    else:
        h2d = (2 * (re_ - rb) + (qe - qb)) * G * 4
        d2h = (re_ - rb) * G * 4
    launches = sum(l[1].launch_count for l in lanes)
    for l in lanes:
        l[1].close()
    return value, h2d, d2h, launches, I.segments


def single_process_run(torch, smk, a, world, nccl_flux):
    """Rank 0 only: the weak workload driven by ONE process over all N GPUs (smk_multi_*, peer-memory
    all-reduce kernel).  The other ranks idle on a CPU-side barrier meanwhile."""
    import numpy as np
    G = a.egroups
    I = smk.Input(source_2D_regions=a.regions_2d, segments=100_000_000 * world, egroups=G,
                  seg_per_thread=a.seg_per_track, seed=a.seed, exp_mode=a.exp, math_mode=a.math).finalize()
    R, F = I.source_3D_regions, I.fine_axial_intervals
    rec = {}
    try:
        with smk.MultiContext(I, world, "peer") as m:
            m.fill_device(0.0)
            for _ in range(3):
                m.run()
            ks, ts = zip(*(m.run() for _ in range(5)))
            inter = float(I.segments) * G
            rec["value"] = inter / (sum(ts) / len(ts))
            rec["kernel_only"] = inter / (sum(ks) / len(ks))
            f0, f1 = m.download_flux(0), m.download_flux(world - 1)
            rec["flux_bit_identical_across_devices"] = bool(np.array_equal(f0.view(np.uint32), f1.view(np.uint32)))
            if nccl_flux is not None:
                d = f0.astype(np.float64) - nccl_flux
                rec["l2rel_vs_nccl_path"] = float(np.linalg.norm(d) / np.linalg.norm(nccl_flux.astype(np.float64)))
            # end to end with host slabs: ONE H2D + NVLink broadcast, sweep, all-reduce, D2H
            o_src = np.random.default_rng(a.seed).random((R, F, G), dtype=np.float32)
            o_flux = np.zeros((R, F, G), np.float32)
            o_sig = np.random.default_rng(a.seed + 1).random((R, G), dtype=np.float32)
            m.upload(o_src, o_flux, o_sig)
            t_up, t_all = [], []
            for _ in range(3):
                t0 = time.perf_counter()
                m.upload(o_src, o_flux, o_sig)
                t1 = time.perf_counter()
                m.run()
                m.download_flux(0)
                t_all.append(time.perf_counter() - t0)
                t_up.append(t1 - t0)
            rec["e2e"] = inter / (sum(t_all) / len(t_all))
            rec["upload_ms"] = 1e3 * sum(t_up) / len(t_up)
            rec["upload"] = "one H2D to device 0 (pageable host memory) + NVLink peer copies to the other devices"
            rec["allreduce"] = "allreduce_peer_slices (own kernel over NVLink peer memory)"
            rec["unit"] = UNIT
    except Exception as e:  # noqa: BLE001  (e.g. peer access unavailable): report, do not hide
        rec["error"] = str(e)
    return rec


def main_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import smk_b200 as smk

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")

    G = a.egroups
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    # ---- the headline sweep ---------------------------------------------------------------------
    main = Sweep(torch, dist, smk, dev, rank, world, egroups=G, segments=a.segments, regions_2d=a.regions_2d,
                 seg_per_track=a.seg_per_track, seed=a.seed, exp=a.exp, math=a.math, geometry=a.geometry,
                 fit_per_sweep=a.fit_per_sweep)
    sampler = ClockSampler(local_rank)
    total_ms, kernel_total_ms, launches = main.timed(a.steps, a.warmup, flush, sampler)
    clocks = sampler.stop()
    value = float(main.I.segments) * G * a.steps / (total_ms * 1e-3)
    hbm_resident = regions_3d(a.regions_2d) * 5 * lane_pad(smk, G) * 4 * 3 > 512e6
    binding, other = main.rooflines(kernel_total_ms / a.steps, clocks, hbm_resident)
    binding["kernel"] = main.ctx.kernel_name
    binding["kernel_ms"] = kernel_total_ms / a.steps
    # measured DRAM bytes of one launch (ncu --set full of this round, profiles/traffic_r02.json).  L2-resident:
    # the working set is read once per launch whatever the segment count; HBM-resident: valid for 1e8 segments
    rec = recorded_traffic("hbm_resident" if hbm_resident else "config2") if G == 128 else None
    if rec and hbm_resident and main.my_segments != 100_000_000:
        rec = None
    binding["traffic"] = rec["dram_bytes_per_launch"] if rec else None
    binding["traffic_source"] = rec["source"] if rec else None
    binding["why"] = ("working set exceeds the L2: DRAM binds" if hbm_resident else
                      "the 38 MB working set is L2-resident (DRAM traffic per launch under `traffic` vs "
                      f"{ALGO_BYTES_PER_INTERSECTION * main.my_segments * G / 1e9:.0f} GB algorithmic): the FP32 pipe binds")
    main_kernel_name = main.ctx.kernel_name
    main.close()

    # ---- extra legs / sub-records ---------------------------------------------------------------
    legs, weak, single = [], None, None
    if not a.no_legs and world == 1:
        # (each opt-in fit-per-sweep leg right after its default-arithmetic twin: the board sits on its power cap and the
        # clock sags over a run, so distant legs are not comparable)
        for name, kw in (("config3_7_groups", dict(egroups=7)),
                         # NOT the default arithmetic: the source fit hoisted out of the segment loop (same results)
                         ("config3_7_groups_fit_per_sweep", dict(egroups=7, fit_per_sweep=True)),
                         ("config4_64_groups_14_regions", dict(egroups=64, regions_2d=10)),
                         ("config4_64_groups_14_regions_fit_per_sweep", dict(egroups=64, regions_2d=10, fit_per_sweep=True)),
                         ("config2_fit_per_sweep", dict(fit_per_sweep=True)),
                         ("hbm_resident_432000_regions", dict(regions_2d=320000, hbm_resident=True)),
                         ("config2_per_segment_geometry", dict(geometry=True)),
                         ("config5_1e10_segments_one_gpu", dict(segments=CONFIG5_SEGMENTS, steps=1, warmup=1))):
            legs.append(run_leg(torch, dist, smk, dev, rank, world, flush, a, name, **kw))
    nccl_flux = None
    if not a.no_legs and world > 1:
        weak = run_leg(torch, dist, smk, dev, rank, world, flush, a, "weak_1e8_segments_per_gpu", steps=5, warmup=3)
        # the flux of the process-per-GPU (NCCL) path on the weak workload, for the single-process comparison
        sw = Sweep(torch, dist, smk, dev, rank, world, egroups=G, segments=100_000_000 * world,
                   regions_2d=a.regions_2d, seg_per_track=a.seg_per_track, seed=a.seed, exp=a.exp, math=a.math)
        sw.step()
        sw.barrier()
        if rank == 0:
            nccl_flux = sw.ctx.download_flux()
        sw.close()

    # ---- end to end through the host-buffer API ---------------------------------------------------
    e2e_steps = max(4, min(a.steps, 10))
    e2e_value, h2d, d2h, e2e_launches, e2e_segments = e2e_run(torch, dist, smk, dev, rank, world, a, e2e_steps)

    # ---- one process, N GPUs (rank 0; the others wait on the CPU) -----------------------------------
    if not a.no_legs and world > 1:
        torch.cuda.synchronize(dev)
        dist.barrier(group=cpu_group)
        if rank == 0:
            single = single_process_run(torch, smk, a, world, nccl_flux)
        dist.barrier(group=cpu_group)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = time_reference(a, steps=2, warmup=1, segments=20_000_000)
        cpu.pop("ms_per_step", None)
        # the reference CPU beside the other BASELINE shapes too (bounded samples, ~1 s each)
        for leg in legs:
            shape = {"config3_7_groups": dict(egroups=7), "config4_64_groups_14_regions": dict(egroups=64, regions_2d=10)}.get(leg["name"])
            if shape is None:
                continue
            try:
                la = argparse.Namespace(**{**vars(a), **shape})
                r = time_reference(la, steps=1, warmup=1, segments=20_000_000)
                leg["cpu_reference"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                        "sample": r["sample"]}
            except Exception as e:      # the leg's GPU numbers stand on their own
                leg["cpu_reference"] = {"unavailable": str(e)[:200]}
        try:   # the same sources built for AVX2+FMA (-march=x86-64-v3): the "fair" CPU figure
            from oracle.oracle import Reference
            if Reference.available("v3"):
                cpu["alt_avx2_fma_build"] = time_reference(a, steps=1, warmup=1, segments=20_000_000, variant="v3")["value"]
        except Exception:
            pass

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a),
            "ns_per_intersection": 1e9 / value,
            "roofline": binding,
            ("compute_roofline" if hbm_resident else "hbm_algorithmic"): other,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "segments_global": e2e_segments,
                    "how": ("pinned host slabs -> smk_upload_async -> sweep -> smk_download_flux_rows_async, two contexts "
                            "in flight (copies of step k+1 overlap the sweep of step k)" if world == 1 else
                            "every rank uploads 1/N of the rows, replicas completed by NVLink broadcasts, sweep, "
                            "slice-wise reduce of the tallies, every rank downloads 1/N of the flux; two contexts in "
                            "flight; bytes are per rank; workload = 1e8 segments per GPU")},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if legs:
            line["configs"] = legs
        if weak is not None:
            line["weak"] = weak
        if single is not None:
            line["single_process"] = single
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()


def lane_pad(smk, G):
    return smk.lib.smk_padded_groups(G)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)

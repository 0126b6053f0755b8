#!/usr/bin/env python
"""Full-size parity evidence for BASELINE.json configs 2, 3, 4: the GPU flux (default FAST/POLY kernel and
the STRICT/GLIBC verification kernel) against a FULL CPU replay of the same stream by the oracle on all
host cores.  Prints a markdown table (kept as profiles/parity_full_rNN.md).  Test infrastructure: this is
the one place where the oracle replays 1e8 segments (about a minute of host time), so it is a tool and
not part of the pytest suite."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smk_b200 as smk  # noqa: E402
from oracle.oracle import F64ACC, Oracle  # noqa: E402

CONFIGS = [
    # name, 2D regions, G, segments, seed
    ("config 2: 128 groups, 1e8 segments", 5000, 128, 100_000_000, 42),
    ("config 3: 7 groups (C5G7-like), 1e8 segments", 5000, 7, 100_000_000, 42),
    ("config 4: 64 groups, 14 regions (contention), 1e7 segments", 10, 64, 10_000_000, 42),
    ("config 4 at 1e8 segments (1.4e6 fp32 adds per tally element)", 10, 64, 100_000_000, 42),
]


def l2rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def main():
    o = Oracle()
    print("# Full-size parity: GPU flux vs full CPU replay of the same stream (oracle, all host cores)\n")
    print("| configuration | mode | L2-rel vs fp32 CPU replay | L2-rel vs f64-accumulated replay | CPU replay vs its own "
          "f64 accumulation | finite pattern | indexing fingerprint | GPU int/s | CPU replay s |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, r2d, G, N, seed in CONFIGS:
        I = smk.Input(source_2D_regions=r2d, segments=N, egroups=G, seed=seed).finalize()
        R, F = I.source_3D_regions, I.fine_axial_intervals
        src, flux0, sig = o.fill(R, F, G, seed)
        t0 = time.perf_counter()
        want = flux0.copy()
        _, chk = o.run(src, want, sig, N, I.seg_per_thread, seed, nthreads=0)
        cpu_s = time.perf_counter() - t0
        want64 = flux0.copy()
        o.run(src, want64, sig, N, I.seg_per_thread, seed, nthreads=0, flags=F64ACC)
        for math_mode, exp_mode in (("fast", "poly"), ("strict", "glibc")):
            I.math_mode, I.exp_mode = math_mode, exp_mode
            with smk.Context(I) as ctx:
                ctx.upload(src, flux0, sig)
                ctx.run()
                sec = ctx.run() if False else None
                ctx.reset_tallies()
                sec = ctx.run()
                got = ctx.download_flux()
                gchk = ctx.checksum()
            print(f"| {name} | {math_mode}/{exp_mode} | {l2rel(got, want):.2e} | {l2rel(got, want64):.2e} | "
                  f"{l2rel(want, want64):.2e} | {'same' if np.array_equal(np.isfinite(got), np.isfinite(want)) else 'DIFFERENT'} | "
                  f"{'identical' if gchk == chk else 'DIFFERENT'} ({chk:016x}) | {N * G / sec:.3e} | {cpu_s:.1f} |", flush=True)


if __name__ == "__main__":
    main()

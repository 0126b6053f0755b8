#!/usr/bin/env python
"""Writes the CPU oracle's scalar flux for the C driver's --verify option (test infrastructure: the
driver itself never touches the oracle; it only reads the resulting raw float32 file).

    python tools/oracle_replay.py -s 2000000 -e 128 --regions-2d 5000 --seed 42 -o /tmp/flux_cpu.bin
    ./simplemoc-kernel_b200/bin/SimpleMOC-kernel -s 2000000 --verify /tmp/flux_cpu.bin
"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import TABLE, Oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-s", "--segments", type=int, default=50_000_000)
    ap.add_argument("-e", "--egroups", type=int, default=128)
    ap.add_argument("-p", "--seg-per-track", type=int, default=100)
    ap.add_argument("-t", "--threads", type=int, default=0)
    ap.add_argument("--regions-2d", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--sigt-floor", type=float, default=0.0)
    ap.add_argument("--table", action="store_true")
    ap.add_argument("-o", "--output", required=True)
    a = ap.parse_args()
    regions = int(math.ceil(a.regions_2d * 27 / 20))
    o = Oracle()
    src, flux, sig = o.fill(regions, 5, a.egroups, a.seed, a.sigt_floor)
    _, chk = o.run(src, flux, sig, a.segments, a.seg_per_track, a.seed, nthreads=a.threads,
                   flags=TABLE if a.table else 0)
    flux.tofile(a.output)
    print(f"oracle replay: {regions} regions, checksum {chk:016x}, wrote {a.output}")


if __name__ == "__main__":
    main()

"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_strict*.so,
built by oracle/Makefile from /root/reference/src/cpu/{kernel.c,init.c}).

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
Each fixture stores the inputs (so it does not depend on the oracle's fill), the stream
parameters, and the reference's outputs: final scalar flux and each track's outgoing psi,
replayed single-threaded in track order through the reference's own attenuate_segment.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Reference, build  # noqa: E402

CASES = [
    # name,        R,  F, G,   N,    p,  seed, table, sigt_floor
    ("g128_exp",   12, 5, 128, 1500, 100, 42,  False, 0.0),
    ("g7_exp",     9,  5, 7,   2000, 100, 43,  False, 0.0),
    ("g64_few",    3,  5, 64,  1200, 50,  44,  False, 0.0),
    ("g13_tail",   7,  4, 13,  1033, 37,  45,  False, 0.0),
    ("g128_table", 12, 5, 128, 1500, 100, 46,  True,  0.0),
    ("g32_wellcond", 10, 5, 32, 1000, 100, 47, False, 0.1),
]


def main():
    build(ref=True)
    o = Oracle()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, R, F, G, N, p, seed, table, floor_ in CASES:
        ref = Reference("strict_table" if table else "strict")
        src, flux0, sig = o.fill(R, F, G, seed, floor_)
        flux = flux0.copy()
        psi = ref.replay(src, flux, sig, N, p, seed, want_psi=True)
        qsr, fai = o.segment_ids(seed, 0, N, R, F)
        np.savez_compressed(os.path.join(here, name + ".npz"), src=src, flux0=flux0, sigT=sig,
                            flux=flux, psi=psi, qsr=qsr, fai=fai,
                            meta=np.array([R, F, G, N, p, seed, int(table)], np.int64),
                            sigt_floor=np.float32(floor_))
        print(name, "flux L2", np.linalg.norm(flux.astype(np.float64)), "max", flux.max())


if __name__ == "__main__":
    main()

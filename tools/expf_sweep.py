"""Exhaustive sweep of the GPU's exp(-tau) evaluations against the host libm expf, over EVERY
binary32 tau in [2^-33, 0.7] (the reachable range of tau = 0.7 * sigT, SURVEY.md section 7 hard
part 1).  Run on the GPU box:  python tools/expf_sweep.py > gpurun_out/expf_sweep.md
Uses the oracle (libm) as the checker; test/diagnostic tool, not part of the product path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smk_b200 as smk  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    o = Oracle()
    lo = np.array([2.0 ** -33], np.float32).view(np.uint32)[0]
    hi = np.array([0.7], np.float32).view(np.uint32)[0]
    modes = ("glibc", "poly", "mufu")
    stats = {m: {} for m in modes}
    chunk = 1 << 26
    for start in range(int(lo), int(hi) + 1, chunk):
        stop = min(start + chunk, int(hi) + 1)
        tau = np.arange(start, stop, dtype=np.uint32).view(np.float32)
        want = o.expf_neg(tau).view(np.uint32).astype(np.int64)
        binade = (np.arange(start, stop, dtype=np.uint32) >> 23).astype(np.int32) - 127
        for m in modes:
            got = smk.debug_exp(m, tau).view(np.uint32).astype(np.int64)
            d = np.abs(got - want)
            for b in np.unique(binade):
                sel = binade == b
                s = stats[m].setdefault(int(b), [0, 0, 0])
                s[0] += int(sel.sum())
                s[1] += int((d[sel] != 0).sum())
                s[2] = max(s[2], int(d[sel].max()))
    print("# exp(-tau) on the GPU vs host libm expf, every binary32 tau in [2^-33, 0.7]\n")
    print("| tau binade | values | " + " | ".join(f"{m}: mismatches (max ulp)" for m in modes) + " |")
    print("|---|---|" + "---|" * len(modes))
    for b in sorted(stats["glibc"]):
        row = [f"[2^{b}, 2^{b+1})", str(stats["glibc"][b][0])]
        for m in modes:
            n, mis, mx = stats[m][b]
            row.append(f"{mis} ({mx})")
        print("| " + " | ".join(row) + " |")
    for m in modes:
        tot = sum(v[1] for v in stats[m].values())
        n = sum(v[0] for v in stats[m].values())
        print(f"\n{m}: {tot} mismatches of {n} values, max {max(v[2] for v in stats[m].values())} ulp")


if __name__ == "__main__":
    main()

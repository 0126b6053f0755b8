#!/usr/bin/env python
"""CPU emulation (numpy) of the FAST arithmetic of csrc/smk_math.cuh: attenuate_fast2 with every fp32
rounding reproduced (fma = one rounding of the exact a*b+c, done in float64; MUFU.RCP approximated by the
correctly rounded reciprocal).  DEVELOPMENT / TEST INFRASTRUCTURE: answers "would this re-association still
meet the 1e-5 gate against the oracle?" without a GPU.  Vectorised over tracks: step k processes segment k
of every track at once.

    python tools/fast_emulator.py --R 14 --G 64 --N 200000 --seed 3 --geom base --spread 0.4 [--variant cur]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import F64ACC, GEOM, REFERENCE_GEOMETRY, Oracle, geometry7  # noqa: E402

f32 = np.float32
f64 = np.float64


def fma(a, b, c):
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def mul(a, b):
    return (np.asarray(a, f32) * np.asarray(b, f32)).astype(f32)


def add(a, b):
    return (np.asarray(a, f32) + np.asarray(b, f32)).astype(f32)


def sub(a, b):
    return (np.asarray(a, f32) - np.asarray(b, f32)).astype(f32)


def rcp(x):
    return (f64(1.0) / np.asarray(x, f64)).astype(f32)


C5, C4, C3, C2, C1 = (f32(float.fromhex(h)) for h in
                      ("0x1.415ffep-13", "0x1.6336e4p-10", "0x1.10ac84p-7", "0x1.555146p-5", "0x1.555546p-3"))


A2 = f32(float.fromhex("0x1.000088p-1"))      # glibc expf's x^2 coefficient for |x| < ln2/64: C1 * (32/ln2)^2
B1 = f32(float.fromhex("0x1.a1bdd2p-33"))     # 1 - C2 * (32/ln2): glibc's x coefficient is not exactly 1
T_SMALL = f32(2.0 ** -8)


def exp_poly_neg(tau, track_glibc=True):
    """e = exp(-tau) as exp_val2<kExpPoly>; returns (e, tau^2)."""
    p = fma(-C5, tau, C4)
    p = fma(p, tau, -C3)
    p = fma(p, tau, C2)
    p = fma(p, tau, -C1)
    x2 = mul(tau, tau)
    s = fma(tau, f32(-1.0), f32(1.0))
    lost = fma(tau, f32(-1.0), sub(f32(1.0), s))
    if track_glibc:
        p = fma(p, tau, np.where(tau < T_SMALL, A2, f32(0.5)))
        e = add(s, fma(tau, fma(tau, p, -B1), lost))
    else:
        p = fma(p, tau, f32(0.5))
        e = add(s, fma(x2, p, lost))
    return e, x2, s


def attenuate_fast(kind, fc, y1, y2, y3, sigT, psi, variant="cur"):
    """kind: 0 interior, 1 first, 2 last (arrays broadcastable); fc: dict of per-segment coefficient
    columns (q0_d, q0_s, q1_d, q1_s, q2_s, e0, e1, ds, weight).  All segment types are evaluated in the
    interior's form and selected afterwards (numerically the typed edge bodies: q2 = 0 there)."""
    interior = (kind == 0)
    first = (kind == 1)
    # interior
    d_i = sub(y1, y3)
    s_i = fma(y2, f32(-2.0), add(y1, y3))
    q0_i = fma(fc["q0_s"], s_i, fma(fc["q0_d"], d_i, y2))
    Q1_i = fma(fc["q1_s"], s_i, mul(fc["q1_d"], d_i))
    Q2_i = mul(fc["q2_s"], s_i)
    # edges
    d_e = np.where(first, sub(y3, y2), sub(y2, y1))
    q0_e = fma(fc["e0"], d_e, y2)
    Q1_e = mul(fc["e1"], d_e)
    q0 = np.where(interior, q0_i, q0_e)
    Q1 = np.where(interior, Q1_i, Q1_e)
    Q2 = np.where(interior, Q2_i, f32(0.0))

    tau = mul(sigT, fc["ds"])
    e, tau2, one_m_tau = exp_poly_neg(tau)
    ev = sub(f32(1.0), e)
    tme = sub(tau, ev)
    rs = rcp(sigT)
    rs2 = mul(rs, rs)
    E = mul(ev, rs)
    Fc = mul(tme, rs2)
    if variant == "old_reuse":
        reuse = fma(f32(2.0), mul(E, rs2), fma(tau, f32(-2.0), tau2))
    else:
        # tau^2 - 2 tau = (1 - tau)^2 - 1 from the exponential's s = RN(1 - tau): one operation
        reuse = fma(f32(2.0), mul(E, rs2), fma(one_m_tau, one_m_tau, f32(-1.0)))
    if variant == "chain_r02f":
        # accumulation order of the kernels up to gpurun r02h
        fi = fma(Q1, reuse, fma(q0, Fc, mul(psi, E)))
        acc = mul(psi, e)
        cubic = sub(mul(tau, fma(tau, add(tau, f32(-3.0)), f32(6.0))), mul(f32(6.0), ev))
        fi_q = fma(mul(Q2, f32(1.0 / 3.0)), mul(cubic, mul(rs2, rs2)), fi)
        acc_q = fma(Q2, reuse, acc)
        fi = np.where(interior, fi_q, fi)
        acc = np.where(interior, acc_q, acc)
        psi_new = fma(q0, E, fma(Q1, Fc, acc))
    else:
        # the two sums advance in lock step, each pair of FMAs sharing its first operand (psi, Q2, Q1, q0)
        fi = mul(psi, E)
        acc = mul(psi, e)
        if variant in ("cur", "old_reuse"):
            cubic = sub(mul(tau, fma(tau, add(tau, f32(-3.0)), f32(6.0))), mul(f32(6.0), ev))   # as ptxas fuses it
        elif variant == "nofuse":
            cubic = sub(mul(tau, add(mul(tau, add(tau, f32(-3.0))), f32(6.0))), mul(f32(6.0), ev))
        else:
            raise ValueError(variant)
        h3 = mul(mul(cubic, mul(rs2, rs2)), f32(1.0 / 3.0))
        fi = np.where(interior, fma(Q2, h3, fi), fi)
        acc = np.where(interior, fma(Q2, reuse, acc), acc)
        fi = fma(Q1, reuse, fi)
        acc = fma(Q1, Fc, acc)
        fi = fma(q0, Fc, fi)
        psi_new = fma(q0, E, acc)
    tally = mul(fc["weight"], fi)
    return psi_new, tally


def coeffs(geom6, dz0):
    """fit_coeffs_geom_typed: geom6 [n][6] = dz, zin, weight, mu, mu2, ds."""
    k1, k2, inv_dz = f32(1.0 / (2.0 * dz0)), f32(1.0 / (2.0 * dz0 * dz0)), f32(1.0 / dz0)
    zin, weight, mu, mu2, ds = (geom6[:, i].astype(f32) for i in (1, 2, 3, 4, 5))
    kz = mul(k2, zin)
    return dict(q0_d=mul(k1, zin), q0_s=mul(kz, zin), q1_d=mul(mu, k1), q1_s=mul(mul(f32(2.0), mu), kz),
                q2_s=mul(mu2, k2), e0=mul(zin, inv_dz), e1=mul(mu, inv_dz), ds=ds, weight=weight)


def run(o, src, flux0, sig, N, p, seed, g7=None, variant="cur"):
    """Replays the whole stream; returns flux (f64-accumulated, rounded to f32 like finalize_flux64)."""
    R, F, G = src.shape
    q, f = o.segment_ids(seed, 0, N, R, F)
    base = REFERENCE_GEOMETRY if g7 is None else tuple(float(x) for x in g7[:6])
    if g7 is None:
        geom6 = np.tile(np.array(base, f32), (N, 1))
    else:
        geom6 = o.segment_geometry(seed, 0, N, g7)
    fc_all = coeffs(geom6, f64(f32(base[0])))
    T = (N + p - 1) // p
    psi = np.stack([o.track_psi0(seed, t, G) for t in range(T)])
    tally64 = np.zeros((R * F, G), f64)
    for k in range(p):
        idx = np.arange(T) * p + k
        idx = idx[idx < N]
        n = idx.size
        if n == 0:
            break
        qq, ff = q[idx], f[idx]
        kind = np.where(ff == 0, 1, np.where(ff == F - 1, 2, 0))[:, None]
        y2 = src[qq, ff]
        y1 = src[qq, np.maximum(ff - 1, 0)]
        y3 = src[qq, np.minimum(ff + 1, F - 1)]
        fc = {name: col[idx][:, None] for name, col in fc_all.items()}
        new_psi, t = attenuate_fast(kind, fc, y1, y2, y3, sig[qq], psi[:n], variant)
        psi[:n] = new_psi
        np.add.at(tally64, qq * F + ff, t.astype(f64))
    return (flux0.astype(f64) + tally64.reshape(R, F, G)).astype(f32)


def l2rel(a, b):
    a, b = a.astype(f64), b.astype(f64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--R", type=int, default=14)
    ap.add_argument("--F", type=int, default=5)
    ap.add_argument("--G", type=int, default=64)
    ap.add_argument("--N", type=int, default=200000)
    ap.add_argument("--p", type=int, default=100)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--geom", default="none", choices=["none", "ref", "base"])
    ap.add_argument("--spread", type=float, default=0.25)
    ap.add_argument("--variant", default="cur")
    a = ap.parse_args()
    o = Oracle()
    src, flux0, sig = o.fill(a.R, a.F, a.G, a.seed)
    g7 = None
    flags = F64ACC
    if a.geom != "none":
        base = REFERENCE_GEOMETRY if a.geom == "ref" else (0.2, 0.05, 0.8, 0.6, 0.36, 0.45)
        g7 = geometry7(base, a.spread)
        flags |= GEOM
    want = flux0.copy()
    o.run(src, want, sig, a.N, a.p, a.seed, nthreads=0, flags=flags, geom7=g7)
    got = run(o, src, flux0, sig, a.N, a.p, a.seed, g7, a.variant)
    print(f"emulated FAST vs oracle(f64 acc): L2-rel {l2rel(got, want):.3e}")
    err = np.abs(got.astype(f64) - want.astype(f64))
    worst = np.unravel_index(np.argsort(err, axis=None)[-5:], err.shape)
    for r, fa, g in zip(*worst):
        print(f"  region {r} fai {fa} group {g}: sigT {sig[r, g]:.6e} want {want[r, fa, g]:.6e} "
              f"diff {err[r, fa, g]:.3e} ({err[r, fa, g] / np.linalg.norm(want.astype(f64)):.2e} of |want|)")


if __name__ == "__main__":
    main()

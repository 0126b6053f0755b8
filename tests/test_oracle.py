"""CPU tests: the oracle against known answers, the committed golden vectors (produced by the
unmodified reference) and, when /root/reference is mounted, the reference's object code."""
import glob
import math
import os

import numpy as np
import pytest

from oracle.oracle import F64ACC, GEOM, REFERENCE_GEOMETRY, TABLE, Oracle, Reference, geometry7

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def l2rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


# Random123 kat_vectors, philox4x32 10 rounds
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(oracle, ctr, key, want):
    assert list(oracle.philox(ctr, key)) == want


def test_stream_ranges_and_distribution(oracle):
    q, f = oracle.segment_ids(42, 0, 200_000, 6750, 5)
    assert q.min() >= 0 and q.max() < 6750 and f.min() >= 0 and f.max() < 5
    assert abs(np.bincount(f, minlength=5) / f.size - 0.2).max() < 0.01
    # counter-based: any sub-range replays identically
    q2, f2 = oracle.segment_ids(42, 123_456, 100, 6750, 5)
    assert np.array_equal(q2, q[123_456:123_556]) and np.array_equal(f2, f[123_456:123_556])
    src, flux, sig = oracle.fill(50, 5, 128, 42)
    for a in (src, flux, sig):
        assert a.min() >= 0.0 and a.max() <= 1.0 and abs(a.mean() - 0.5) < 0.01
    _, _, sig_floor = oracle.fill(50, 5, 128, 42, 0.1)
    assert sig_floor.min() >= 0.1 and sig_floor.max() <= 1.0


def test_known_answer_flat_source(oracle):
    """y1=y2=y3=1, sigT=1, psi=0, interior interval =>
    c1=c2=0, q0=1, tau=0.7, expVal=1-e^-0.7, tally=0.5*(0.7-expVal), psi_out=expVal."""
    src = np.ones((5, 4), np.float32)
    sig = np.ones(4, np.float32)
    psi = np.zeros(4, np.float32)
    tally = oracle.attenuate_segment(2, src, sig, psi)
    ev = 1.0 - math.exp(-0.7)
    assert np.allclose(tally, 0.5 * (0.7 - ev), rtol=2e-6)
    assert np.allclose(psi, ev, rtol=2e-6)


def test_known_answer_linear_edges(oracle):
    """FAI == 0 uses (y2, y3), FAI == F-1 uses (y1, y2) (kernel.c:111-161), in double by hand."""
    rng = np.random.default_rng(1)
    src = rng.random((5, 8)).astype(np.float32)
    sig = (0.2 + 0.8 * rng.random(8)).astype(np.float32)
    for fai, (lo, hi) in ((0, (0, 1)), (4, (3, 4))):
        psi0 = rng.random(8).astype(np.float32)
        psi = psi0.copy()
        tally = oracle.attenuate_segment(fai, src, sig, psi)
        y2 = src[fai].astype(np.float64)
        c1 = (src[hi].astype(np.float64) - src[lo]) / 0.1
        q0, q1, s = y2 + c1 * 0.3, c1, sig.astype(np.float64)
        tau = s * 0.7
        ev = 1 - np.exp(-tau)
        reuse = tau * (tau - 2) + 2 * ev / s ** 3
        fi = (q0 * tau + (s * psi0 - q0) * ev) / s ** 2 + q1 * 0.9 * reuse
        assert np.allclose(tally, 0.5 * fi, rtol=1e-4)
        pso = q0 * ev / s + q1 * 0.9 * (tau - ev) / s ** 2 + psi0 * (1 - ev)
        assert np.allclose(psi, pso, rtol=1e-4)


def test_table_constants(oracle):
    """init.c:81-117: N = 353, dx = 10/353, maxVal = 10 - dx, entries {-e^-x, 1 + (x-1) e^-x}."""
    n, vals, dx, maxval = oracle.build_table()
    assert n == 353
    assert dx == np.float32(10.0) / np.float32(353.0)
    assert maxval == np.float32(10.0) - np.float32(dx)
    x = np.arange(353) * np.float64(dx)
    assert np.allclose(vals[0::2], -np.exp(-x), rtol=1e-6)
    assert np.allclose(vals[1::2], 1 + (x - 1) * np.exp(-x), rtol=1e-5, atol=1e-7)
    assert oracle.table_lookup(vals, dx, maxval, 20.0) == 1.0   # kernel.c:340-341


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors_bit_exact(oracle, path):
    """The restatement reproduces the reference's flux and psi bit for bit (single thread)."""
    z = np.load(path)
    R, F, G, N, p, seed, table = (int(v) for v in z["meta"])
    flux = z["flux0"].copy()
    psi, _ = oracle.run(z["src"], flux, z["sigT"], N, p, seed, want_psi=True, nthreads=1,
                        flags=TABLE if table else 0)
    assert np.array_equal(bits(flux), bits(z["flux"]))
    assert np.array_equal(bits(psi), bits(z["psi"]))
    q, f = oracle.segment_ids(seed, 0, N, R, F)
    assert np.array_equal(q, z["qsr"]) and np.array_equal(f, z["fai"])
    src, flux0, sig = oracle.fill(R, F, G, seed, float(z["sigt_floor"]))
    assert np.array_equal(bits(src), bits(z["src"])) and np.array_equal(bits(sig), bits(z["sigT"]))
    assert np.array_equal(bits(flux0), bits(z["flux0"]))


@pytest.mark.skipif(not Reference.available("strict"), reason="oracle/_ref not built")
@pytest.mark.parametrize("table", [False, True])
def test_restatement_vs_reference_object_code(oracle, table):
    """Fresh random case against the compiled, unmodified kernel.c (exp and TABLE builds)."""
    ref = Reference("strict_table" if table else "strict")
    assert bool(ref.lib.ref_build_flags() & 1) == table
    R, F, G, N, p, seed = 30, 5, 100, 9000, 100, 2024
    src, flux0, sig = oracle.fill(R, F, G, seed)
    a, b = flux0.copy(), flux0.copy()
    psi_a, _ = oracle.run(src, a, sig, N, p, seed, want_psi=True, nthreads=1, flags=TABLE if table else 0)
    psi_b = ref.replay(src, b, sig, N, p, seed, want_psi=True)
    assert np.array_equal(bits(a), bits(b)) and np.array_equal(bits(psi_a), bits(psi_b))
    if table:
        n, vals, dx, mv = oracle.build_table()
        n2, vals2, dx2, mv2 = ref.table()
        assert (n, dx, mv) == (n2, dx2, mv2) and np.array_equal(bits(vals), bits(vals2))


def test_threads_and_sharding_agree(oracle):
    """OpenMP replay and track-range sharding only reorder the tally additions."""
    R, F, G, N, p, seed = 20, 5, 64, 20_000, 100, 5
    src, flux0, sig = oracle.fill(R, F, G, seed, 0.1)
    serial = flux0.copy()
    psi_s, chk_s = oracle.run(src, serial, sig, N, p, seed, want_psi=True, nthreads=1)
    par = flux0.copy()
    psi_p, chk_p = oracle.run(src, par, sig, N, p, seed, want_psi=True, nthreads=4)
    assert chk_s == chk_p and np.array_equal(bits(psi_s), bits(psi_p))
    assert l2rel(par, serial) < 1e-6   # tallies of mixed sign cancel: norm-wise, not element-wise
    # two shards, tallies added
    nt = (N + p - 1) // p
    z1, z2 = np.zeros_like(flux0), np.zeros_like(flux0)
    _, c1 = oracle.run(src, z1, sig, N, p, seed, 0, nt // 2, nthreads=1)
    _, c2 = oracle.run(src, z2, sig, N, p, seed, nt // 2, nt, nthreads=1)
    assert (c1 + c2) % 2 ** 64 == chk_s
    assert l2rel(flux0 + z1 + z2, serial) < 1e-6
    # f64 accumulation is the error yardstick
    acc = flux0.copy()
    oracle.run(src, acc, sig, N, p, seed, nthreads=2, flags=F64ACC)
    assert l2rel(acc, serial) < 1e-6


def test_rejects_bad_arguments(oracle):
    src, flux, sig = oracle.fill(4, 5, 8, 1)
    with pytest.raises(ValueError):
        oracle.run(src, flux, sig, 100, 0, 1)
    with pytest.raises(ValueError):
        oracle.run(src, flux, sig, 100, 10, 1, track_begin=0, track_end=11)
    # empty input: nothing happens
    before = flux.copy()
    oracle.run(src, flux, sig, 0, 10, 1)
    assert np.array_equal(before, flux)


# ---------------------------------------------------------------------------------------
# per-segment geometry (kernel.c:95-104 made parameters)
# ---------------------------------------------------------------------------------------
def test_geometry_draws(oracle):
    """Four 16-bit fields of stream words 2,3 -> factors in [1-spread, 1+spread); dz is global."""
    g7 = geometry7(spread=0.25)
    g = oracle.segment_geometry(42, 0, 100_000, g7)
    base = np.array(REFERENCE_GEOMETRY, np.float32)
    assert (g[:, 0] == base[0]).all()                                   # dz
    for col in (1, 2, 3, 5):                                            # zin, weight, mu, ds
        f = g[:, col].astype(np.float64) / base[col]
        assert f.min() >= 0.75 - 1e-6 and f.max() < 1.25 + 1e-6 and abs(f.mean() - 1.0) < 5e-3
    f_mu = g[:, 3].astype(np.float64) / base[3]
    assert np.allclose(g[:, 4], base[4] * f_mu ** 2, rtol=3e-7)        # mu2 follows mu^2
    # zin and ds come from different halves of one word, mu and weight of the other: uncorrelated
    assert abs(np.corrcoef(g[:, 1], g[:, 5])[0, 1]) < 0.02 and abs(np.corrcoef(g[:, 3], g[:, 2])[0, 1]) < 0.02
    # counter-based: a sub-range replays identically; spread = 0 is the base, exactly
    assert np.array_equal(bits(oracle.segment_geometry(42, 777, 50, g7)), bits(g[777:827]))
    g0 = oracle.segment_geometry(42, 0, 1000, geometry7(spread=0.0))
    assert np.array_equal(bits(g0), bits(np.tile(base, (1000, 1))))


@pytest.mark.skipif(not Reference.available("strict"), reason="oracle/_ref not built")
def test_geometry_constant_point_vs_reference_object_code(oracle):
    """Pin (i): the parametrised restatement, driven through its per-segment path but with the draws
    mapped onto the constants (spread = 0), is bit-identical to the unmodified kernel.c."""
    ref = Reference("strict")
    R, F, G, N, p, seed = 30, 5, 100, 9000, 100, 77
    src, flux0, sig = oracle.fill(R, F, G, seed)
    b = flux0.copy()
    psi_b = ref.replay(src, b, sig, N, p, seed, want_psi=True)
    a = flux0.copy()
    psi_a, _ = oracle.run(src, a, sig, N, p, seed, want_psi=True, nthreads=1, flags=GEOM,
                          geom7=geometry7(spread=0.0))
    assert np.array_equal(bits(a), bits(b)) and np.array_equal(bits(psi_a), bits(psi_b))
    # Power-of-two gauge: (dz, zin, mu, mu2) -> (2 dz, 2 zin, 2 mu, 4 mu2) leaves q0, q1 mu and q2 mu2
    # unchanged in exact binary arithmetic (c1 halves, c2 quarters, kernel.c:182-189), so a restatement
    # that uses every parameter where kernel.c uses its constant reproduces the reference bit for bit;
    # doubling the weight doubles every tally exactly (kernel.c:262).
    dz, zin, w, mu, mu2, ds = REFERENCE_GEOMETRY
    c = np.zeros_like(flux0)
    psi_c, _ = oracle.run(src, c, sig, N, p, seed, want_psi=True, nthreads=1, flags=GEOM,
                          geom7=geometry7((2 * dz, 2 * zin, 2 * w, 2 * mu, 4 * mu2, ds), 0.0))
    d = np.zeros_like(flux0)
    ref.replay(src, d, sig, N, p, seed)                                 # reference tallies on zero flux
    assert np.array_equal(bits(psi_c), bits(psi_b))
    assert np.array_equal(bits(c), bits(np.float32(2.0) * d))


def test_geometry_known_answer_in_double(oracle):
    """One interior and one edge segment with a non-reference geometry against the formulae of
    SURVEY.md section 3.3 evaluated in double precision."""
    rng = np.random.default_rng(5)
    src = rng.random((5, 8)).astype(np.float32)
    sig = (0.3 + 0.7 * rng.random(8)).astype(np.float32)
    dz, zin, w, mu, mu2, ds = 0.2, 0.05, 0.8, 0.6, 0.36, 0.45
    for fai in (2, 0, 4):
        psi0 = rng.random(8).astype(np.float32)
        psi = psi0.copy()
        tally = oracle.attenuate_segment(fai, src, sig, psi, geom6=(dz, zin, w, mu, mu2, ds))
        y1, y2, y3 = (src[min(max(fai + k, 0), 4)].astype(np.float64) for k in (-1, 0, 1))
        if fai == 0:
            c1, c2 = (y3 - y2) / dz, 0.0
        elif fai == 4:
            c1, c2 = (y2 - y1) / dz, 0.0
        else:
            c1, c2 = (y1 - y3) / (2 * dz), (y1 - 2 * y2 + y3) / (2 * dz * dz)
        q0, q1, q2 = y2 + c1 * zin + c2 * zin * zin, c1 + 2 * c2 * zin, c2
        s = sig.astype(np.float64)
        tau = s * ds
        ev = 1 - np.exp(-tau)
        reuse = tau * (tau - 2) + 2 * ev / s ** 3
        fi = ((q0 * tau + (s * psi0 - q0) * ev) / s ** 2 + q1 * mu * reuse
              + q2 * mu2 * (tau * (tau * (tau - 3) + 6) - 6 * ev) / (3 * s ** 4))
        pso = q0 * ev / s + q1 * mu * (tau - ev) / s ** 2 + q2 * mu2 * reuse + psi0 * (1 - ev)
        assert np.allclose(tally, w * fi, rtol=2e-4, atol=1e-5)
        assert np.allclose(psi, pso, rtol=2e-4, atol=1e-5)


def test_geometry_spread_changes_results_smoothly(oracle):
    """With per-segment draws the sweep differs from the constant one by O(spread), same indexing."""
    R, F, G, N, p, seed = 20, 5, 32, 20_000, 100, 9
    src, flux0, sig = oracle.fill(R, F, G, seed, 0.1)
    a, b, c = flux0.copy(), flux0.copy(), flux0.copy()
    _, chk_a = oracle.run(src, a, sig, N, p, seed, nthreads=1)
    _, chk_b = oracle.run(src, b, sig, N, p, seed, nthreads=1, flags=GEOM, geom7=geometry7(spread=0.01))
    _, chk_c = oracle.run(src, c, sig, N, p, seed, nthreads=1, flags=GEOM, geom7=geometry7(spread=0.3))
    assert chk_a == chk_b == chk_c
    assert 0 < l2rel(b, a) < 0.01 and 5 * l2rel(b, a) < l2rel(c, a) < 0.5      # zero-mean draws average out over a sweep

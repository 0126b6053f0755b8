"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

* ``Oracle``    -- oracle/liboracle.so, the C restatement (oracle/smk_oracle.c) of
  /root/reference/src/cpu/kernel.c:75-361 + init.c:81-117 driven by the
  deterministic stream of DESIGN.md section 3.
* ``Reference`` -- oracle/_ref/libref_*.so, the UNMODIFIED reference sources built
  by oracle/Makefile (present only if they were built while /root/reference was
  mounted; the built files travel to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

TABLE = 1
F64ACC = 2
GEOM = 4        # per-segment geometry from stream words 2,3 (kernel.c:95-104)

# kernel.c:99-104: dz, zin, weight, mu, mu2, ds (+ spread of the per-segment variation)
REFERENCE_GEOMETRY = (0.1, 0.3, 0.5, 0.9, 0.3, 0.7)


def geometry7(base=REFERENCE_GEOMETRY, spread=0.0) -> np.ndarray:
    """{dz, zin, weight, mu, mu2, ds, spread} as the float32 vector the C side takes."""
    g = np.array(list(base) + [spread], np.float32)
    assert g.shape == (7,)
    return g

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (when /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def n_tracks(segments: int, seg_per_track: int) -> int:
    if seg_per_track < 1:
        raise ValueError("seg_per_track must be >= 1")
    return (segments + seg_per_track - 1) // seg_per_track


class Oracle:
    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.smk_oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.smk_oracle_philox4x32_10.restype = None
        L.smk_oracle_segment_ids.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int,
                                             C.c_int, _i32p, _i32p]
        L.smk_oracle_segment_ids.restype = None
        L.smk_oracle_track_psi0.argtypes = [C.c_uint64, C.c_int64, C.c_int, _f32p]
        L.smk_oracle_track_psi0.restype = None
        L.smk_oracle_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_uint64, C.c_float]
        L.smk_oracle_fill.restype = None
        L.smk_oracle_build_table.argtypes = [C.c_float, C.c_float, _f32p, C.c_int,
                                             C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.smk_oracle_build_table.restype = C.c_int
        L.smk_oracle_table_lookup.argtypes = [_f32p, C.c_float, C.c_float, C.c_float]
        L.smk_oracle_table_lookup.restype = C.c_float
        L.smk_oracle_expf.argtypes = [C.c_float]
        L.smk_oracle_expf.restype = C.c_float
        L.smk_oracle_expf_neg_array.argtypes = [_f32p, _f32p, C.c_int64]
        L.smk_oracle_expf_neg_array.restype = None
        L.smk_oracle_attenuate_segment.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, _f32p,
                                                   _f32p, _f32p, C.c_int]
        L.smk_oracle_attenuate_segment.restype = None
        L.smk_oracle_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                     _f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_void_p,
                                     C.POINTER(C.c_uint64), C.c_int, C.c_uint]
        L.smk_oracle_run.restype = C.c_int
        L.smk_oracle_run_geom.argtypes = L.smk_oracle_run.argtypes + [C.c_void_p]
        L.smk_oracle_run_geom.restype = C.c_int
        L.smk_oracle_segment_geometry.argtypes = [C.c_uint64, C.c_int64, C.c_int64, _f32p, _f32p]
        L.smk_oracle_segment_geometry.restype = None
        L.smk_oracle_attenuate_segment_geom.argtypes = L.smk_oracle_attenuate_segment.argtypes + [_f32p]
        L.smk_oracle_attenuate_segment_geom.restype = None
        L.smk_oracle_max_threads.restype = C.c_int

    # -- stream -------------------------------------------------------------
    def philox(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self.lib.smk_oracle_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
        return out

    def segment_ids(self, seed, seg_begin, count, regions, fai):
        q = np.zeros(count, np.int32)
        f = np.zeros(count, np.int32)
        self.lib.smk_oracle_segment_ids(seed, seg_begin, count, regions, fai, q, f)
        return q, f

    def track_psi0(self, seed, track, groups):
        psi = np.zeros(groups, np.float32)
        self.lib.smk_oracle_track_psi0(seed, track, groups, psi)
        return psi

    def fill(self, regions, fai, groups, seed, sigt_floor=0.0):
        src = np.zeros((regions, fai, groups), np.float32)
        flux = np.zeros((regions, fai, groups), np.float32)
        sig = np.zeros((regions, groups), np.float32)
        self.lib.smk_oracle_fill(src.ctypes.data, flux.ctypes.data, sig.ctypes.data,
                                 regions, fai, groups, seed, sigt_floor)
        return src, flux, sig

    # -- table --------------------------------------------------------------
    def build_table(self):
        vals = np.zeros(706, np.float32)
        dx, mv = C.c_float(), C.c_float()
        n = self.lib.smk_oracle_build_table(0.01, 10.0, vals, 706, C.byref(dx), C.byref(mv))
        return n, vals, dx.value, mv.value

    def table_lookup(self, vals, dx, maxval, x):
        return self.lib.smk_oracle_table_lookup(vals, dx, maxval, x)

    def expf(self, x):
        return self.lib.smk_oracle_expf(x)

    def expf_neg(self, tau):
        """libm expf(-tau), elementwise."""
        tau = np.ascontiguousarray(tau, np.float32)
        out = np.empty_like(tau)
        self.lib.smk_oracle_expf_neg_array(tau, out, tau.size)
        return out

    # -- math ---------------------------------------------------------------
    def attenuate_segment(self, fai_id, src_region, sigt_region, psi, use_table=False, geom6=None):
        """src_region [F][G], sigt_region [G], psi [G] (updated in place). Returns tally [G].
        geom6 = (dz, zin, weight, mu, mu2, ds) overrides the constants of kernel.c:99-104."""
        fai_count, groups = src_region.shape
        tally = np.zeros(groups, np.float32)
        args = (groups, fai_count, fai_id, np.ascontiguousarray(src_region, np.float32),
                np.ascontiguousarray(sigt_region, np.float32), psi, tally, int(use_table))
        if geom6 is None:
            self.lib.smk_oracle_attenuate_segment(*args)
        else:
            self.lib.smk_oracle_attenuate_segment_geom(*args, np.asarray(geom6, np.float32))
        return tally

    def segment_geometry(self, seed, seg_begin, count, geom7):
        """(count, 6) array of dz, zin, weight, mu, mu2, ds of each segment."""
        out = np.zeros((count, 6), np.float32)
        self.lib.smk_oracle_segment_geometry(seed, seg_begin, count, np.asarray(geom7, np.float32), out)
        return out

    def run(self, src, flux, sig, segments, seg_per_track, seed, track_begin=0, track_end=None,
            want_psi=False, nthreads=0, flags=0, geom7=None):
        """Replays tracks [track_begin, track_end); flux is updated IN PLACE.
        Returns (psi_final or None, id_checksum).  geom7 (see geometry7()) replaces the constants of
        kernel.c:99-104; with flags & GEOM they additionally vary per segment."""
        regions, fai, groups = src.shape
        nt = n_tracks(segments, seg_per_track)
        track_end = nt if track_end is None else track_end
        psi = np.zeros((track_end - track_begin, groups), np.float32) if want_psi else None
        chk = C.c_uint64(0)
        g7 = None if geom7 is None else np.ascontiguousarray(geom7, np.float32)
        rc = self.lib.smk_oracle_run_geom(regions, fai, groups, segments, seg_per_track, seed,
                                          src, flux, sig, track_begin, track_end,
                                          psi.ctypes.data if want_psi else None, C.byref(chk),
                                          nthreads, flags, None if g7 is None else g7.ctypes.data)
        if rc != 0:
            raise ValueError(f"smk_oracle_run rejected its arguments (code {rc})")
        return psi, chk.value

    def max_threads(self):
        return self.lib.smk_oracle_max_threads()


class Reference:
    """The unmodified reference CPU sources, built into oracle/_ref by oracle/Makefile."""

    VARIANTS = ("strict", "strict_table", "ofast", "ofast_table", "v3")

    def __init__(self, variant: str = "strict"):
        path = os.path.join(REF_DIR, f"libref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.variant = variant
        L = self.lib = C.CDLL(path)
        L.ref_build_flags.restype = C.c_int
        L.ref_replay_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                     _f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_void_p]
        L.ref_replay_run.restype = C.c_int
        L.ref_table.argtypes = [_f32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ref_table.restype = C.c_int
        L.ref_time_run_kernel.argtypes = [C.c_int, C.c_int, C.c_long, C.c_int]
        L.ref_time_run_kernel.restype = C.c_double
        L.ref_num_procs.restype = C.c_int

    @staticmethod
    def available(variant: str = "strict") -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))

    def replay(self, src, flux, sig, segments, seg_per_track, seed, track_begin=0,
               track_end=None, want_psi=False):
        regions, fai, groups = src.shape
        nt = n_tracks(segments, seg_per_track)
        track_end = nt if track_end is None else track_end
        psi = np.zeros((track_end - track_begin, groups), np.float32) if want_psi else None
        self.lib.ref_replay_run(regions, fai, groups, segments, seg_per_track, seed, src, flux,
                                sig, track_begin, track_end,
                                psi.ctypes.data if want_psi else None)
        return psi

    def table(self):
        vals = np.zeros(706, np.float32)
        dx, mv = C.c_float(), C.c_float()
        n = self.lib.ref_table(vals, C.byref(dx), C.byref(mv))
        return n, vals, dx.value, mv.value

    def time_run_kernel(self, regions_2d, groups, segments, nthreads=0) -> float:
        return self.lib.ref_time_run_kernel(regions_2d, groups, segments, nthreads)

    def num_procs(self) -> int:
        return self.lib.ref_num_procs()

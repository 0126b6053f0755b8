"""B200-native SimpleMOC-kernel segment-attenuation path: ctypes binding of libsmk.so.

This package is the Python-side mirror of the C ABI in include/smk.h (the product's
host driver is the plain-C program host/smk_main.c; this module exists for tests,
bench.py and torch.distributed plumbing).  All compute happens in the hand-written
sm_100a kernels of csrc/; there is NO CPU fallback: if lib/libsmk.so is missing or
does not load, importing this package raises.

Reference interface mirrored here
  Input / set_default_input   /root/reference/src/cpu/SimpleMOC-kernel_header.h:24-42,
                              /root/reference/src/cpu/init.c:4-24 (+ streams /
                              seg_per_thread of /root/reference/src/cuda/init.cu:30-44)
  run_kernel(I, S)            /root/reference/src/cpu/kernel.c:3-73
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import multi  # noqa: F401  (sharding + all-reduce plumbing)

HERE = os.path.dirname(os.path.abspath(__file__))
# SMK_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("SMK_LIB") or os.path.join(HERE, "lib", "libsmk.so")

EXP_POLY, EXP_MUFU, EXP_GLIBC, EXP_TABLE = 0, 1, 2, 3
MATH_FAST, MATH_STRICT = 0, 1
FLAG_KEEP_PSI = 1
FLAG_TALLY_F64 = 2
FLAG_SEGMENT_GEOMETRY = 4
FLAG_FIT_PER_SWEEP = 8
ARRAY_SOURCE, ARRAY_FLUX, ARRAY_SIGT = 0, 1, 2
DEBUG_EXP_PACKED, DEBUG_EXP_WIDE, DEBUG_EXP_TRACK = 0x100, 0x200, 0x400
# kernel.c:99-104: dz, zin, weight, mu, mu2, ds
REFERENCE_GEOMETRY = (0.1, 0.3, 0.5, 0.9, 0.3, 0.7)

EXP_MODES = {"poly": EXP_POLY, "mufu": EXP_MUFU, "glibc": EXP_GLIBC, "table": EXP_TABLE}
MATH_MODES = {"fast": MATH_FAST, "strict": MATH_STRICT}


class SmkError(RuntimeError):
    pass


class Params(C.Structure):
    """struct smk_params (include/smk.h)."""
    _fields_ = [
        ("source_3D_regions", C.c_int32),
        ("fine_axial_intervals", C.c_int32),
        ("egroups", C.c_int32),
        ("seg_per_track", C.c_int32),
        ("segments", C.c_int64),
        ("seed", C.c_uint64),
        ("exp_mode", C.c_int32),
        ("math_mode", C.c_int32),
        ("device", C.c_int32),
        ("flags", C.c_int32),
    ]


class Geometry(C.Structure):
    """struct smk_geometry (include/smk.h): base values of kernel.c:99-104 + per-segment spread."""
    _fields_ = [("dz", C.c_float), ("zin", C.c_float), ("weight", C.c_float), ("mu", C.c_float),
                ("mu2", C.c_float), ("ds", C.c_float), ("spread", C.c_float)]


def build(verbose: bool = False) -> str:
    """Compile lib/libsmk.so and bin/SimpleMOC-kernel for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", HERE, "all"] + ([] if verbose else ["-s"])
    subprocess.run(cmd, check=True)
    return LIB_PATH


_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

# every symbol include/smk.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "smk_abi_version", "smk_last_error", "smk_device_count", "smk_device_name",
    "smk_padded_groups", "smk_num_tracks", "smk_create", "smk_destroy", "smk_set_stream",
    "smk_upload", "smk_fill_device", "smk_reset_tallies", "smk_download_flux",
    "smk_download_psi", "smk_download_checksum", "smk_run", "smk_run_async",
    "smk_synchronize", "smk_launch_count", "smk_run_host", "smk_device_tally",
    "smk_device_flux0", "smk_device_source", "smk_device_sigT", "smk_padded_elems",
    "smk_alloc_host", "smk_free_host", "smk_debug_exp", "smk_debug_segment_ids",
    "smk_multi_create", "smk_multi_destroy", "smk_multi_upload", "smk_multi_fill_device",
    "smk_multi_run", "smk_multi_download_flux", "smk_multi_download_checksum",
    "smk_multi_device_count", "smk_multi_set_geometry",
    "smk_set_geometry", "smk_get_geometry", "smk_kernel_name", "smk_upload_async",
    "smk_upload_rows_async", "smk_scan_sigt_max", "smk_download_flux_rows_async",
    "smk_debug_segment_geometry", "smk_set_sigt_bound", "smk_wait_finalized",
)


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is the product and there is no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C simplemoc-kernel_b200`.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.smk_abi_version.restype = i32
    L.smk_last_error.restype = C.c_char_p
    L.smk_device_count.restype = i32
    L.smk_device_name.argtypes = [i32, C.c_char_p, C.c_size_t]
    L.smk_padded_groups.argtypes = [i32]
    L.smk_num_tracks.argtypes = [i64, i32]
    L.smk_num_tracks.restype = i64
    L.smk_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.smk_destroy.argtypes = [vp]
    L.smk_destroy.restype = None
    L.smk_set_stream.argtypes = [vp, vp]
    L.smk_upload.argtypes = [vp, vp, vp, vp]
    L.smk_device_sigT.argtypes = [vp]
    L.smk_fill_device.argtypes = [vp, C.c_float]
    L.smk_reset_tallies.argtypes = [vp]
    L.smk_download_flux.argtypes = [vp, vp]
    L.smk_download_psi.argtypes = [vp, vp, i64]
    L.smk_set_geometry.argtypes = [vp, C.POINTER(Geometry)]
    L.smk_get_geometry.argtypes = [vp, C.POINTER(Geometry)]
    L.smk_kernel_name.argtypes = [vp]
    L.smk_kernel_name.restype = C.c_char_p
    L.smk_upload_async.argtypes = [vp, vp, vp, vp]
    L.smk_upload_rows_async.argtypes = [vp, i32, i64, i64, vp]
    L.smk_scan_sigt_max.argtypes = [vp, C.POINTER(C.c_float)]
    L.smk_set_sigt_bound.argtypes = [vp, C.c_float]
    L.smk_download_flux_rows_async.argtypes = [vp, i64, i64, vp]
    L.smk_wait_finalized.argtypes = [vp, vp]
    L.smk_multi_set_geometry.argtypes = [vp, C.POINTER(Geometry)]
    L.smk_debug_segment_geometry.argtypes = [C.POINTER(Params), C.POINTER(Geometry), i64, i64, _f32p]
    L.smk_download_checksum.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.smk_run.argtypes = [vp, i64, i64, C.POINTER(C.c_double)]
    L.smk_run_async.argtypes = [vp, i64, i64]
    L.smk_synchronize.argtypes = [vp]
    L.smk_launch_count.argtypes = [vp]
    L.smk_launch_count.restype = i64
    L.smk_run_host.argtypes = [C.POINTER(Params), vp, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in ("smk_device_tally", "smk_device_flux0", "smk_device_source", "smk_device_sigT"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.smk_padded_elems.argtypes = [vp]
    L.smk_padded_elems.restype = i64
    L.smk_alloc_host.argtypes = [C.c_size_t]
    L.smk_alloc_host.restype = vp
    L.smk_free_host.argtypes = [vp]
    L.smk_free_host.restype = None
    L.smk_multi_create.argtypes = [C.POINTER(Params), i32, C.POINTER(C.c_int), i32, C.POINTER(vp)]
    L.smk_multi_destroy.argtypes = [vp]
    L.smk_multi_destroy.restype = None
    L.smk_multi_upload.argtypes = [vp, vp, vp, vp]
    L.smk_multi_fill_device.argtypes = [vp, C.c_float]
    L.smk_multi_run.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.smk_multi_download_flux.argtypes = [vp, i32, vp]
    L.smk_multi_download_checksum.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.smk_multi_device_count.argtypes = [vp]
    L.smk_debug_exp.argtypes = [i32, _f32p, _f32p, i64, i32]
    L.smk_debug_segment_ids.argtypes = [C.POINTER(Params), i64, i64, _i32p, _i32p]
    return L


lib = _load()


def _check(rc: int) -> None:
    if rc != 0:
        raise SmkError(f"libsmk error {rc}: {lib.smk_last_error().decode()}")


# ---------------------------------------------------------------------------
# the reference's Input (header.h:24-42) with the CUDA variant's extra fields
# ---------------------------------------------------------------------------
@dataclass
class Input:
    source_2D_regions: int = 5000        # init.c:8
    coarse_axial_intervals: int = 27     # init.c:9
    fine_axial_intervals: int = 5        # init.c:10
    decomp_assemblies_ax: int = 20       # init.c:11
    segments: int = 50_000_000           # init.c:12
    egroups: int = 128                   # init.c:13
    nthreads: int = 0                    # -t: CPU threads of the verification replay only
    seg_per_thread: int = 100            # -p (cuda/init.cu:41): segments per track
    source_3D_regions: int = 0           # derived, main.c:18-19
    # changed subsystems (north star): stream seed, exponential and arithmetic modes, device
    seed: int = 42
    exp_mode: str = "poly"
    math_mode: str = "fast"
    device: int = 0
    tally_f64: bool = False              # diagnostic: f64 tally accumulators (SMK_FLAG_TALLY_F64)
    # per-segment geometry (SMK_FLAG_SEGMENT_GEOMETRY): kernel.c:99-104 as base values + spread
    segment_geometry: bool = False
    # OFF by default: evaluate the axial source fit once per (region, interval, group) per sweep instead of once per
    # segment (SMK_FLAG_FIT_PER_SWEEP; identical results, fewer operations in the segment loop; <= 64 groups)
    fit_per_sweep: bool = False
    geometry: tuple = REFERENCE_GEOMETRY
    geometry_spread: float = 0.25

    def finalize(self) -> "Input":
        """main.c:18-19: source_3D_regions = ceil(2D * coarse / decomp)."""
        self.source_3D_regions = int(math.ceil(
            float(self.source_2D_regions) * self.coarse_axial_intervals / self.decomp_assemblies_ax))
        return self

    @property
    def n_tracks(self) -> int:
        return (self.segments + self.seg_per_thread - 1) // self.seg_per_thread

    def params(self, flags: int = 0) -> Params:
        if self.source_3D_regions == 0:
            self.finalize()
        if self.tally_f64:
            flags |= FLAG_TALLY_F64
        if self.segment_geometry:
            flags |= FLAG_SEGMENT_GEOMETRY
        if self.fit_per_sweep:
            flags |= FLAG_FIT_PER_SWEEP
        return Params(self.source_3D_regions, self.fine_axial_intervals, self.egroups,
                      self.seg_per_thread, self.segments, self.seed, EXP_MODES[self.exp_mode],
                      MATH_MODES[self.math_mode], self.device, flags)


def set_default_input() -> Input:
    """init.c:4-24 / cuda init.cu:30-44."""
    return Input().finalize()


class Context:
    """Device-resident problem: padded source / sigT / flux arrays + tally deltas."""

    def __init__(self, I: Input, keep_psi: bool = False):
        self.I = I
        self.p = I.params(FLAG_KEEP_PSI if keep_psi else 0)
        self._h = C.c_void_p()
        _check(lib.smk_create(C.byref(self.p), C.byref(self._h)))
        self.R, self.F, self.G = self.p.source_3D_regions, self.p.fine_axial_intervals, self.p.egroups
        self.G_pad = lib.smk_padded_groups(self.G)
        self.n_tracks = lib.smk_num_tracks(self.p.segments, self.p.seg_per_track)
        if I.segment_geometry:
            self.set_geometry(I.geometry, I.geometry_spread)

    def close(self):
        if self._h:
            lib.smk_destroy(self._h)
            self._h = C.c_void_p()

    def set_geometry(self, base=REFERENCE_GEOMETRY, spread: float = 0.25):
        g = Geometry(*base, spread)
        _check(lib.smk_set_geometry(self._h, C.byref(g)))

    @property
    def kernel_name(self) -> str:
        return lib.smk_kernel_name(self._h).decode()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(a):
        """host pointer of a numpy array / torch CPU tensor / raw int address (or None)."""
        if a is None:
            return None
        if isinstance(a, int):
            return a
        if isinstance(a, np.ndarray):
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
            return a.ctypes.data
        return a.data_ptr()  # torch tensor

    def upload(self, fine_source, fine_flux, sigT):
        _check(lib.smk_upload(self._h, self._ptr(fine_source), self._ptr(fine_flux), self._ptr(sigT)))

    def upload_async(self, fine_source, fine_flux, sigT):
        """Enqueue only; the host arrays must stay unchanged until synchronize()."""
        _check(lib.smk_upload_async(self._h, self._ptr(fine_source), self._ptr(fine_flux), self._ptr(sigT)))

    def upload_rows_async(self, array: int, row_begin: int, rows: int, host):
        """rows [row_begin, row_begin+rows) of ARRAY_SOURCE / ARRAY_FLUX (R*F rows) or ARRAY_SIGT (R rows)."""
        _check(lib.smk_upload_rows_async(self._h, array, row_begin, rows, self._ptr(host)))

    def scan_sigt_max(self) -> float:
        v = C.c_float(0.0)
        _check(lib.smk_scan_sigt_max(self._h, C.byref(v)))
        return v.value

    def set_sigt_bound(self, bound: float):
        _check(lib.smk_set_sigt_bound(self._h, bound))

    def download_flux_rows_async(self, row_begin: int, rows: int, out):
        _check(lib.smk_download_flux_rows_async(self._h, row_begin, rows, self._ptr(out)))

    def wait_finalized(self, other: "Context"):
        """Everything enqueued on this context from now on waits (on the device) for the flux0 + tallies pass of
        `other`'s last download: keeps that small kernel from starving behind this context's persistent sweep."""
        _check(lib.smk_wait_finalized(self._h, other._h))

    def fill_device(self, sigt_floor: float = 0.0):
        _check(lib.smk_fill_device(self._h, sigt_floor))

    def reset_tallies(self):
        _check(lib.smk_reset_tallies(self._h))

    def set_stream(self, cuda_stream: int):
        _check(lib.smk_set_stream(self._h, cuda_stream))

    def run(self, track_begin: int = 0, track_end: int | None = None) -> float:
        """Synchronous; returns the CUDA-event time of the kernel in seconds."""
        t = C.c_double(0.0)
        te = self.n_tracks if track_end is None else track_end
        _check(lib.smk_run(self._h, track_begin, te, C.byref(t)))
        return t.value

    def run_async(self, track_begin: int = 0, track_end: int | None = None):
        te = self.n_tracks if track_end is None else track_end
        _check(lib.smk_run_async(self._h, track_begin, te))

    def synchronize(self):
        _check(lib.smk_synchronize(self._h))

    def download_flux(self, out=None):
        if out is None:
            out = np.empty((self.R, self.F, self.G), np.float32)
        _check(lib.smk_download_flux(self._h, self._ptr(out)))
        return out

    def download_psi(self, n_tracks: int):
        out = np.empty((n_tracks, self.G), np.float32)
        _check(lib.smk_download_psi(self._h, out.ctypes.data, n_tracks))
        return out

    def checksum(self) -> int:
        v = C.c_uint64(0)
        _check(lib.smk_download_checksum(self._h, C.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        return lib.smk_launch_count(self._h)

    # raw device pointers for torch / NCCL plumbing
    @property
    def tally_ptr(self) -> int:
        return lib.smk_device_tally(self._h)

    @property
    def padded_elems(self) -> int:
        return lib.smk_padded_elems(self._h)

    @property
    def source_ptr(self) -> int:
        return lib.smk_device_source(self._h)

    @property
    def sigt_ptr(self) -> int:
        return lib.smk_device_sigT(self._h)


class MultiContext:
    """One process, several GPUs: tracks sharded by range, one all-reduce of the tally deltas
    (allreduce = "peer": NVLink peer-memory kernel, "nccl": ncclAllReduce)."""

    def __init__(self, I: Input, n_devices: int, allreduce: str = "peer", devices=None):
        self.I = I
        self.p = I.params()
        self._h = C.c_void_p()
        ids = (C.c_int * n_devices)(*devices) if devices else None
        _check(lib.smk_multi_create(C.byref(self.p), n_devices, ids, {"peer": 0, "nccl": 1}[allreduce],
                                    C.byref(self._h)))
        self.R, self.F, self.G = self.p.source_3D_regions, self.p.fine_axial_intervals, self.p.egroups
        if I.segment_geometry:
            g = Geometry(*I.geometry, I.geometry_spread)
            _check(lib.smk_multi_set_geometry(self._h, C.byref(g)))

    def close(self):
        if self._h:
            lib.smk_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, fine_source, fine_flux, sigT):
        _check(lib.smk_multi_upload(self._h, Context._ptr(fine_source), Context._ptr(fine_flux), Context._ptr(sigT)))

    def fill_device(self, sigt_floor: float = 0.0):
        _check(lib.smk_multi_fill_device(self._h, sigt_floor))

    def run(self):
        """Returns (slowest kernel seconds, wall seconds including the all-reduce)."""
        ks, ts = C.c_double(0.0), C.c_double(0.0)
        _check(lib.smk_multi_run(self._h, C.byref(ks), C.byref(ts)))
        return ks.value, ts.value

    def download_flux(self, which: int = 0):
        out = np.empty((self.R, self.F, self.G), np.float32)
        _check(lib.smk_multi_download_flux(self._h, which, out.ctypes.data))
        return out

    def checksum(self) -> int:
        v = C.c_uint64(0)
        _check(lib.smk_multi_download_checksum(self._h, C.byref(v)))
        return v.value


def run_kernel(I: Input, fine_source: np.ndarray, fine_flux: np.ndarray, sigT: np.ndarray):
    """Drop-in for run_kernel(I, S, table) (kernel.c:3-73) with host slabs: fine_flux is
    updated in place.  Returns (kernel_seconds, total_seconds)."""
    p = I.params()
    ks, ts = C.c_double(0.0), C.c_double(0.0)
    _check(lib.smk_run_host(C.byref(p), fine_source.ctypes.data, fine_flux.ctypes.data,
                            sigT.ctypes.data, C.byref(ks), C.byref(ts)))
    return ks.value, ts.value


def debug_exp(exp_mode: str, tau: np.ndarray, device: int = 0, packed: bool = False, wide: bool = False,
              track: bool = False) -> np.ndarray:
    """exp(-tau) as the kernels evaluate it; packed = the FP32x2 form of the FAST kernels,
    wide = POLY's wide-range form (MUFU.EX2 beyond tau = 0.7), track = the packed form of the
    per-segment-geometry kernels (follows libm where libm is not correctly rounded)."""
    tau = np.ascontiguousarray(tau, np.float32)
    out = np.empty_like(tau)
    mode = EXP_MODES[exp_mode] | (DEBUG_EXP_PACKED if packed else 0) | (DEBUG_EXP_WIDE if wide else 0) | \
        (DEBUG_EXP_TRACK if track else 0)
    _check(lib.smk_debug_exp(mode, tau, out, tau.size, device))
    return out


def debug_segment_geometry(I: Input, seg_begin: int, n: int) -> np.ndarray:
    """(n, 6) dz, zin, weight, mu, mu2, ds of segments [seg_begin, seg_begin+n) as the kernels derive them."""
    p = I.params()
    g = Geometry(*I.geometry, I.geometry_spread)
    out = np.empty((n, 6), np.float32)
    _check(lib.smk_debug_segment_geometry(C.byref(p), C.byref(g), seg_begin, n, out))
    return out


def debug_segment_ids(I: Input, seg_begin: int, n: int):
    p = I.params()
    q = np.empty(n, np.int32)
    f = np.empty(n, np.int32)
    _check(lib.smk_debug_segment_ids(C.byref(p), seg_begin, n, q, f))
    return q, f


def device_count() -> int:
    return lib.smk_device_count()


def alloc_pinned(shape, dtype=np.float32) -> np.ndarray:
    """numpy view of cudaMallocHost memory (never freed explicitly: lives for the process)."""
    n = int(np.prod(shape))
    ptr = lib.smk_alloc_host(n * np.dtype(dtype).itemsize)
    if not ptr:
        raise SmkError(lib.smk_last_error().decode())
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)

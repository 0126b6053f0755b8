"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/smk.h declares, and the host-side logic (shapes, argument validation) behaves."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "smk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smk_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(smk):
    assert declared_functions() == sorted(smk.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(smk):
    out = subprocess.run(["nm", "-D", "--defined-only", smk.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    exported = set(re.findall(r" T (smk_\w+)", out))
    missing = [f for f in declared_functions() if f not in exported]
    assert not missing, missing
    for f in declared_functions():
        assert getattr(smk.lib, f) is not None
    assert smk.lib.smk_abi_version() == 2


def test_no_torch_or_cuda_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "smk.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "cudaStream_t" not in code and "torch" not in code and "#include <cuda" not in code


def test_padded_groups(smk):
    # G <= 128: next power of two of ceil(G/4) float4; beyond: whole blocks of 256 groups, any G
    want = {1: 4, 4: 4, 7: 8, 8: 8, 13: 16, 32: 32, 64: 64, 100: 128, 128: 128, 129: 256, 256: 256,
            300: 512, 1024: 1024, 1025: 1280, 2000: 2048, 5000: 5120}
    for g, gp in want.items():
        assert smk.lib.smk_padded_groups(g) == gp
    assert smk.lib.smk_padded_groups(0) < 0
    assert smk.lib.smk_num_tracks(1033, 37) == 28
    assert smk.lib.smk_num_tracks(0, 100) == 0


def test_input_defaults_match_reference(smk):
    """init.c:8-13, cuda init.cu:41, main.c:18-19."""
    I = smk.set_default_input()
    assert (I.source_2D_regions, I.coarse_axial_intervals, I.fine_axial_intervals,
            I.decomp_assemblies_ax, I.segments, I.egroups, I.seg_per_thread) == (5000, 27, 5, 20, 50_000_000, 128, 100)
    assert I.source_3D_regions == 6750
    assert I.n_tracks == 500_000


def test_create_validates_before_touching_cuda(smk):
    h = C.c_void_p()
    bad = [dict(source_3D_regions=0), dict(fine_axial_intervals=1), dict(egroups=0), dict(flags=64),
           dict(seg_per_track=0), dict(segments=-1), dict(exp_mode=9), dict(math_mode=5),
           dict(source_3D_regions=2 ** 24, egroups=2048)]        # row offsets would not fit 32 bits
    for kw in bad:
        base = dict(source_3D_regions=10, fine_axial_intervals=5, egroups=8, seg_per_track=10, segments=100,
                    seed=1, exp_mode=0, math_mode=0, device=0, flags=0)
        base.update(kw)
        p = smk.Params(**base)
        assert smk.lib.smk_create(C.byref(p), C.byref(h)) == -1   # SMK_EINVAL
        assert smk.lib.smk_last_error()


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful without a GPU")
def test_fails_loudly_without_gpu(smk):
    """No CPU fallback: without a device the product path errors out."""
    assert smk.device_count() == 0
    with pytest.raises(smk.SmkError):
        smk.Context(smk.Input(source_2D_regions=10, segments=100).finalize())
    exe = os.path.join(ROOT, "simplemoc-kernel_b200", "bin", "SimpleMOC-kernel")
    r = subprocess.run([exe, "-s", "1000"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stdout


def test_host_driver_cli_usage():
    """Unknown flag -> usage text + exit(1), like io.c:152-153,163-173."""
    exe = os.path.join(ROOT, "simplemoc-kernel_b200", "bin", "SimpleMOC-kernel")
    r = subprocess.run([exe, "-x"], capture_output=True, text=True)
    assert r.returncode == 1
    for opt in ("-t <threads>", "-s <segments>", "-e <energy groups>", "-p <segs per thread>", "-d <CUDA device ID>"):
        assert opt in r.stdout
    r = subprocess.run([exe, "-s"], capture_output=True, text=True)
    assert r.returncode == 1


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing in the product may import, link, dlopen or run it."""
    import glob
    product = glob.glob(os.path.join(ROOT, "simplemoc-kernel_b200", "**", "*.*"), recursive=True) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.join(ROOT, "smk_b200.py")]
    for path in product:
        if path.endswith((".so", ".o", ".cubin", ".sass", ".log", ".pyc")) or "/build/" in path or "/bin/" in path:
            continue
        text = open(path, errors="ignore").read()
        for needle in ("liboracle", "from oracle", "import oracle", "oracle/", "smk_oracle_", "libref_"):
            assert needle not in text, f"{path} references {needle!r}"
    out = subprocess.run(["ldd", os.path.join(ROOT, "simplemoc-kernel_b200", "lib", "libsmk.so")],
                         capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out


def test_multi_and_misc_argument_validation(smk):
    """Argument errors are reported (negative code + message), never a crash, before any CUDA work."""
    h = C.c_void_p()
    p = smk.Params(10, 5, 8, 10, 100, 1, 0, 0, 0, 0)
    assert smk.lib.smk_multi_create(C.byref(p), 0, None, 0, C.byref(h)) == -1
    assert smk.lib.smk_multi_create(C.byref(p), 9, None, 0, C.byref(h)) == -1
    assert smk.lib.smk_multi_create(C.byref(p), 2, None, 7, C.byref(h)) == -1
    assert smk.lib.smk_multi_create(None, 2, None, 0, C.byref(h)) == -1
    assert smk.lib.smk_upload(None, None, None, None) == -1
    assert smk.lib.smk_run(None, 0, 0, None) == -1
    assert smk.lib.smk_download_flux(None, None) == -1
    assert smk.lib.smk_download_psi(None, None, 0) == -1
    assert smk.lib.smk_upload_async(None, None, None, None) == -1
    assert smk.lib.smk_upload_rows_async(None, 0, 0, 0, None) == -1
    assert smk.lib.smk_download_flux_rows_async(None, 0, 0, None) == -1
    assert smk.lib.smk_wait_finalized(None, None) == -1
    assert smk.lib.smk_set_geometry(None, None) == -1
    assert smk.lib.smk_scan_sigt_max(None, None) == -1
    assert smk.lib.smk_kernel_name(None) == b""
    g = smk.Geometry(0.1, 0.3, 0.5, 0.9, 0.3, 0.7, 1.5)        # spread must be < 1
    out = np.zeros((4, 6), np.float32)
    assert smk.lib.smk_debug_segment_geometry(C.byref(p), C.byref(g), 0, 4, out) == -1
    assert smk.lib.smk_multi_run(None, None, None) == -1
    assert smk.lib.smk_launch_count(None) == 0
    assert smk.lib.smk_padded_elems(None) == 0
    smk.lib.smk_destroy(None)          # no-ops
    smk.lib.smk_multi_destroy(None)
    smk.lib.smk_free_host(None)
    assert smk.lib.smk_last_error()

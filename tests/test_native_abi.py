"""CPU-side checks of the drop-in boundary: libgss.so builds in-tree, loads, exports
every symbol include/gss_api.h declares, and rejects bad arguments without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from gan_sass_tf_b200 import build, _native
    build.build()
    _native.lib()
    return _native


def _declared_symbols(experimental=False):
    """symbols include/gss_api.h declares; the GSS_EXPERIMENTAL block counts only for the experimental flavour"""
    src = open(os.path.join(ROOT, "include", "gss_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    if not experimental:
        src = re.sub(r"#ifdef GSS_EXPERIMENTAL.*?#endif", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gss_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native):
    syms = _declared_symbols()
    assert len(syms) >= 18
    h = ctypes.CDLL(native.LIB_PATH)
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/gss_api.h but not exported by libgss.so"
    # and the ctypes table binds exactly the declared surface
    assert sorted(native.SIGNATURES) == syms
    # the product library keeps no process-wide switches: they live in the experimental flavour only
    xs = _declared_symbols(experimental=True)
    extra = sorted(set(xs) - set(syms))
    assert extra == sorted(native.EXPERIMENTAL_SIGNATURES) == ["gss_set_path", "gss_set_synth_variant"]
    for s in extra:
        assert not hasattr(h, s), f"{s} must not be exported by the product library"
    hx = ctypes.CDLL(native.EXPERIMENTAL_LIB_PATH)
    for s in xs:
        assert hasattr(hx, s), f"{s} missing from libgss_experimental.so"


def test_version_and_sizes(native):
    assert native.lib().gss_version() >= 100
    sizes = native.supported_fft_sizes()
    assert 512 in sizes and all(s & (s - 1) == 0 for s in sizes)


@pytest.mark.parametrize("n,N,H,T,nadd", [
    (48000, 256, 128, 376, 0), (48000, 512, 128, 376, 0), (46797, 512, 128, 367, 51), (64000, 512, 128, 501, 0),
    (960000, 1024, 256, 3751, 0), (48000, 4096, 1024, 48, 128), (48000, 256, 64, 751, 0),
])
def test_frame_count_k1(native, n, N, H, T, nadd):
    from oracle import ref_oracle as R
    assert native.frame_count(n, N, H) == (T, nadd) == R.frame_count(n, N, H)


def test_argument_errors_without_gpu(native):
    lib = native.lib()
    with pytest.raises(ValueError):
        native.frame_count(0, 512, 128)
    # bad FFT size / hop / null pointers are rejected before any CUDA call
    assert lib.gss_stft_packed(None, 1, 4000, 4000, 500, 125, 0, 1e-7, None, None) == native.GSS_EUNSUPPORTED
    assert b"power of two" in lib.gss_last_error()
    assert lib.gss_stft_packed(None, 1, 4000, 4000, 512, 100, 0, 1e-7, None, None) == native.GSS_EUNSUPPORTED
    assert lib.gss_stft_packed(None, 1, 4000, 4000, 512, 128, 0, 1e-7, None, None) == native.GSS_EINVAL
    assert lib.gss_istft_packed(None, 1, 10, 512, 128, 0, 1e-7, None, 1152, None) == native.GSS_EINVAL
    assert lib.gss_mask_istft(None, None, 1, 3, 4000, 4000, 512, 128, None, 4096, None) == native.GSS_EINVAL
    with pytest.raises(ValueError):
        native.check(native.GSS_EINVAL)


def test_ops_reject_cpu_tensors(native):
    import torch
    from gan_sass_tf_b200.app import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.stft(torch.zeros(1, 4000), 512, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.to_log_signal(torch.zeros(1, 4, 256))


def test_plain_c_consumer(native, tmp_path):
    """include/gss_api.h compiles as C and a gcc-built program links libgss.so and gets the same answers."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = os.path.join(tmp_path, "abi_smoke")
    libdir = os.path.dirname(native.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe,
                    "-L", libdir, "-lgss", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout

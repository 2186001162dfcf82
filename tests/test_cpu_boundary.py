"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol that
include/toeplitz_b200.h declares, the header and the ctypes binding agree, and the numpy
model of the kernels' index algebra reproduces numpy's FFT."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("tb_build", os.path.join(ROOT, "toeplitzmatrices.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "toeplitz_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol(built_lib):
    import ctypes
    lib = ctypes.CDLL(built_lib)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s


def test_binding_matches_header(built_lib):
    import toeplitz_b200 as tb
    assert sorted(tb.EXPORTS) == header_symbols()
    from toeplitz_b200 import _lib
    L = _lib.lib()
    assert b"sm_100a" in L.tb200_version()
    # argument validation happens before any CUDA call: usable without a GPU
    import ctypes
    out = ctypes.c_void_p()
    st = L.tb200_factorize(99, 1, 4, 4, None, None, None, ctypes.byref(out))
    assert st == _lib.BAD_ARG and b"kind" in L.tb200_last_error()
    st = L.tb200_factorize(_lib.SYMMETRIC, 1, 4, 5, ctypes.c_void_p(16), None, None, ctypes.byref(out))
    assert st == _lib.DIM_MISMATCH
    with pytest.raises(_lib.ArgumentError):
        _lib.check(L.tb200_set_option(b"no_such_option", 1))
    assert L.tb200_launch_count() == 0


def test_missing_library_fails_loudly(monkeypatch):
    from toeplitz_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtoeplitz_b200.so")
    with pytest.raises(_lib.ToeplitzB200Error):
        _lib.lib()


@pytest.mark.parametrize("loge", [3, 4])
@pytest.mark.parametrize("l", [4, 5, 6, 7, 8, 9])
def test_fft_index_model_single_level(l, loge):
    import fft_model
    from fft_model import Slot, unscramble
    fft_model.set_radix(loge)
    E = fft_model.E
    rng = np.random.default_rng(l)
    S = Slot(l)
    x = rng.standard_normal(S.L) + 1j * rng.standard_normal(S.L)
    regs = S.forward(x)
    X = np.fft.fft(x)
    for t in range(S.NT):
        for i in range(E):
            assert abs(regs[t][i] - X[S.freq_of_pos(S.pos_of_reg(i, t))]) < 1e-11
    assert np.allclose(S.inverse(regs) / S.L, x, atol=1e-12)
    assert np.allclose(unscramble(S.spectrum_layout(regs), 0, l), X, atol=1e-11)
    for k in range(S.L):
        assert S.freq_of_pos(S.pos_of_freq(k)) == k


@pytest.mark.parametrize("loge", [3, 4])
@pytest.mark.parametrize("l1,l2", [(4, 4), (6, 5), (5, 7)])
def test_fft_index_model_two_level(l1, l2, loge):
    import fft_model
    from fft_model import two_level_forward, unscramble
    fft_model.set_radix(loge)
    rng = np.random.default_rng(l1 * 10 + l2)
    Np = 1 << (l1 + l2)
    x = rng.standard_normal(Np) + 1j * rng.standard_normal(Np)
    assert np.allclose(unscramble(two_level_forward(x, l1, l2), l1, l2), np.fft.fft(x), atol=1e-10)


def test_every_runtime_option_is_documented_in_the_header():
    """`tb200_set_option` names parsed in csrc/capi.cu all appear in include/toeplitz_b200.h (the `_f64` / `_f32` pairs are
    listed there once as `name_f64/f32`)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "toeplitzmatrices.jl_b200", "csrc", "capi.cu")).read()
    hdr = open(os.path.join(root, "include", "toeplitz_b200.h")).read()
    names = sorted(set(re.findall(r'strcmp\(name, "([a-z0-9_]+)"\)', src)))
    assert len(names) >= 15
    for name in names:
        stem = re.sub(r"_f(32|64)$", "", name)
        assert ('"%s"' % name) in hdr or ('"%s_f64/f32"' % stem) in hdr, "option %s is not documented in the header" % name

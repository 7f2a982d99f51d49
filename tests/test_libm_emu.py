"""The device code re-evaluates the reference's float math with bit-exact restatements of glibc's log2f / exp2f /
log10f (mapad_b200/csrc/libm_emu.cuh).  Here the host build of that header is compared with the libm of this
machine bit for bit — sampled by default (every 4099th float), exhaustively with MAPAD_EXHAUSTIVE=1 — and the
`-m gpu` test does the same for the code running on the device."""
import ctypes as C
import ctypes.util
import os

import numpy as np
import pytest

from emu import emu

libm = C.CDLL(ctypes.util.find_library("m"))
for name in ("log2f", "exp2f", "log10f"):
    getattr(libm, name).restype = C.c_float
    getattr(libm, name).argtypes = [C.c_float]


def host_libm(name, x):
    f = getattr(libm, name)
    return np.array([f(float(v)) for v in x], dtype=np.float32)


def sample_bits(stride, lo=0, hi=0x7F800000):
    return np.arange(lo, hi, stride, dtype=np.uint32).view(np.float32)


def interesting():
    pos = sample_bits(1, 0, 4096)                                 # zero and tiny subnormals
    around_one = np.arange(0x3F7FF000, 0x3F801000, dtype=np.uint32).view(np.float32)
    probs = np.float32(1.0) - np.linspace(0, 1, 3001, dtype=np.float32)
    return np.concatenate([pos, around_one, probs, sample_bits(1, 0x7F7FF000, 0x7F800001)])


STRIDE = 1 if os.environ.get("MAPAD_EXHAUSTIVE") == "1" else 4099 * 7


@pytest.mark.parametrize("fn,name", [(0, "log2f"), (2, "log10f")])
def test_log_functions_match_glibc(fn, name):
    x = np.concatenate([sample_bits(STRIDE), interesting()])
    want = host_libm(name, x)
    got = emu.libm(fn, x)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_exp2f_matches_glibc():
    pos = sample_bits(STRIDE)
    x = np.concatenate([pos, -pos, np.linspace(-160, 2, 20001, dtype=np.float32)])
    want = host_libm("exp2f", x)
    got = emu.libm(1, x)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_powi():
    x = np.array([0.5, 0.475, 0.6, 0.55, 0.9, 0.0, 1.0, 0.3], dtype=np.float32)
    for n in (1, 2, 3, 7, 30, 51, 100, 151):
        got = emu.libm(3, x, iarg=n)
        want = np.ones_like(x)
        # compiler-rt __powisf2: square-and-multiply
        a, b = x.copy(), n
        r = np.ones_like(x)
        while True:
            if b & 1:
                r = (r * a).astype(np.float32)
            b //= 2
            if b == 0:
                break
            a = (a * a).astype(np.float32)
        assert np.array_equal(got.view(np.uint32), r.view(np.uint32)), n


@pytest.mark.gpu
def test_device_libm_matches_glibc():
    from mapad_b200 import api
    x = np.concatenate([sample_bits(4099 * 3), interesting()])
    for fn, name in [(0, "log2f"), (2, "log10f")]:
        got = api.debug_libm(fn, x)
        want = emu.libm(fn, x)  # host build == glibc (tests above)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), name
    pos = sample_bits(4099 * 3)
    xe = np.concatenate([pos, -pos, np.linspace(-160, 2, 20001, dtype=np.float32)])
    assert np.array_equal(emu.libm(1, xe).view(np.uint32), api.debug_libm(1, xe).view(np.uint32))
    xp = np.array([0.5, 0.475, 0.6, 0.55, 0.9], dtype=np.float32)
    for n in (1, 7, 51, 151):
        assert np.array_equal(emu.libm(3, xp, iarg=n).view(np.uint32), api.debug_libm(3, xp, iarg=n).view(np.uint32))

"""crt_powf5 (csrc/crt_powf5.h) must equal the host libm's powf(x, 5) -- the call the reference makes at RayTracer.cpp:407 --
bit for bit.  CPU: the header compiled with gcc against libm.  GPU: the same header evaluated on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = r'''
#include <math.h>
#include <stdint.h>
#include "%s/course-assignment-danielhalachev_b200/csrc/crt_powf5.h"
void host_powf5(const float *x, unsigned n, float *mine, float *libm) {
  for (unsigned i = 0; i < n; i++) { mine[i] = crt_powf5(x[i]); libm[i] = powf(x[i], 5); }
}
''' % ROOT


def _inputs():
    rng = np.random.default_rng(5)
    bits = np.concatenate([
        np.arange(0, 0x40000000, 997, dtype=np.uint32),                     # [0, 2): every 997th float
        rng.integers(0, 0x7f800000, 200000, dtype=np.uint32),               # any positive finite
        rng.integers(0, 0x00800000, 20000, dtype=np.uint32),                # subnormals
        rng.integers(0x80000000, 0xC0000000, 50000, dtype=np.uint32),       # negatives (odd power keeps the sign)
        np.array([0, 0x80000000, 0x7f800000, 0xff800000, 0x7fc00000, 0x3f800000, 1, 0x007fffff], dtype=np.uint32)])
    return bits.view(np.float32)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("powf5")
    src = d / "p.c"
    src.write_text(SRC)
    so = str(d / "p.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, str(src), "-lm"], check=True)
    return C.CDLL(so)


def _same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def test_powf5_host_matches_libm(host_lib):
    x = _inputs()
    mine, libm = np.zeros_like(x), np.zeros_like(x)
    host_lib.host_powf5(x.ctypes.data_as(C.c_void_p), C.c_uint(x.size), mine.ctypes.data_as(C.c_void_p), libm.ctypes.data_as(C.c_void_p))
    assert _same(mine, libm).all()


@pytest.mark.gpu
def test_powf5_device_matches_libm(host_lib, built):
    x = _inputs()
    mine, libm = np.zeros_like(x), np.zeros_like(x)
    host_lib.host_powf5(x.ctypes.data_as(C.c_void_p), C.c_uint(x.size), mine.ctypes.data_as(C.c_void_p), libm.ctypes.data_as(C.c_void_p))
    ctx = built.Context(0)
    dev = ctx.debug_powf5(x)
    ctx.close()
    assert _same(dev, libm).all()

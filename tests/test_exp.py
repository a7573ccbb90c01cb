"""exp_glibc (csrc/pmaf_math.cuh) must be bit-identical to the host libm's exp(), which is what
the reference's attractorForceScaling calls (cf_agent.cpp:220)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


@pytest.fixture(scope="module")
def hostexp():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libhostexp.so")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                    "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, os.path.join(HERE, "host_exp_check.cu")],
                   check=True)
    lib = C.CDLL(so)
    lib.hostexp_eval.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    return lib


def _eval(lib, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    lib.hostexp_eval(x.ctypes.data_as(dp), y.ctypes.data_as(dp), x.size)
    return y


def test_exp_glibc_matches_host_libm_bit_for_bit(hostexp):
    libm = C.CDLL("libm.so.6")
    libm.exp.restype = C.c_double
    libm.exp.argtypes = [C.c_double]
    rng = np.random.default_rng(0)
    # the planner's domain: x = -sqrt(d)/shell, d in [1e-5, shell); plus a wide sweep and edge values
    d = rng.uniform(1e-5, 0.8, 400000)
    x = np.concatenate([-np.sqrt(d) / rng.choice([0.35, 0.6, 0.8], d.size), rng.uniform(-700, 700, 200000),
                        -rng.uniform(0, 1, 200000) ** 8, [0.0, -0.0, 1e-300, -1e-300, 2.0 ** -54, -2.0 ** -54, 511.9999,
                                                          -511.9999, 512.0, -745.0, 709.0, np.inf, -np.inf]])
    want = np.array([libm.exp(float(v)) for v in x])
    got = _eval(hostexp, x)
    bad = np.flatnonzero(got.view(np.uint64) != want.view(np.uint64))
    assert bad.size == 0, (x[bad[:5]], got[bad[:5]], want[bad[:5]])
    assert np.isnan(_eval(hostexp, [np.nan])[0])

"""Shared pytest configuration.

Markers:  gpu -- needs a CUDA device (run on the B200 box: ``pytest -m gpu``); everything else runs
on CPU (``pytest -m "not gpu"``).  The GPU tests call the product through its C ABI / drop-in
modules and use ``oracle/`` only as the checker.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sif-xco2-cokriging_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (os.path.join(PKG, "src"), PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference modules (build container only)."""
    import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not present on this machine")
    return ref_loader.load()


@pytest.fixture(scope="session")
def hostmath():
    """Host instantiation of csrc/ck_math.cuh (validation build, g++)."""
    src = os.path.join(ROOT, "tests", "hostmath", "ck_hostmath.cpp")
    so = os.path.join(ROOT, "tests", "hostmath", "_ck_hostmath.so")
    deps = [src, os.path.join(PKG, "csrc", "ck_math.cuh"), os.path.join(PKG, "csrc", "ck_matern_setup.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)

    class H:
        @staticmethod
        def besselk(nu, x):
            x = np.ascontiguousarray(x, float)
            out = np.empty_like(x)
            assert lib.ckh_besselk(ctypes.c_double(nu), x.ctypes.data_as(dp), ctypes.c_long(x.size), out.ctypes.data_as(dp)) == 0
            return out

        @staticmethod
        def matern_cov(scale, nu, ell, nugget, h):
            h = np.ascontiguousarray(h, float)
            out = np.empty_like(h)
            rc = lib.ckh_matern_cov(ctypes.c_double(scale), ctypes.c_double(nu), ctypes.c_double(ell), ctypes.c_double(nugget),
                                    h.ctypes.data_as(dp), ctypes.c_long(h.size), out.ctypes.data_as(dp))
            assert rc == 0
            return out

        @staticmethod
        def distance(metric, X1, X2):
            X1, X2 = np.ascontiguousarray(X1, float), np.ascontiguousarray(X2, float)
            out = np.empty((len(X1), len(X2)))
            lib.ckh_distance(ctypes.c_int(metric), X1.ctypes.data_as(dp), ctypes.c_long(len(X1)), X2.ctypes.data_as(dp),
                             ctypes.c_long(len(X2)), out.ctypes.data_as(dp))
            return out

    def _dist_fast(metric, X1, X2):
        X1, X2 = np.ascontiguousarray(X1, float), np.ascontiguousarray(X2, float)
        out = np.empty((len(X1), len(X2)))
        lib.ckh_distance_fast(ctypes.c_int(metric), X1.ctypes.data_as(dp), ctypes.c_long(len(X1)), X2.ctypes.data_as(dp),
                              ctypes.c_long(len(X2)), out.ctypes.data_as(dp))
        return out

    def _matern_fast(scale, nu, ell, nugget, h):
        h = np.ascontiguousarray(h, float)
        out = np.empty_like(h)
        rc = lib.ckh_matern_cov_fast(ctypes.c_double(scale), ctypes.c_double(nu), ctypes.c_double(ell), ctypes.c_double(nugget),
                                     h.ctypes.data_as(dp), ctypes.c_long(h.size), out.ctypes.data_as(dp))
        assert rc == 0
        return out

    def _pieces(x):
        x = np.ascontiguousarray(x, float)
        o = [np.empty_like(x) for _ in range(3)]
        lib.ckh_fast_pieces(x.ctypes.data_as(dp), ctypes.c_long(x.size), *(v.ctypes.data_as(dp) for v in o))
        return o

    def _dist_pre(X1, X2):
        X1, X2 = np.ascontiguousarray(X1, float), np.ascontiguousarray(X2, float)
        out = np.empty((len(X1), len(X2)))
        lib.ckh_distance_pre(X1.ctypes.data_as(dp), ctypes.c_long(len(X1)), X2.ctypes.data_as(dp), ctypes.c_long(len(X2)),
                             out.ctypes.data_as(dp))
        return out

    def _matern_table(scale, nu, ell, nugget, h):
        h = np.ascontiguousarray(h, float)
        out = np.empty_like(h)
        rc = lib.ckh_matern_cov_table(ctypes.c_double(scale), ctypes.c_double(nu), ctypes.c_double(ell), ctypes.c_double(nugget),
                                      h.ctypes.data_as(dp), ctypes.c_long(h.size), out.ctypes.data_as(dp))
        assert rc == 0, rc
        return out

    H.matern_cov_table = staticmethod(_matern_table)
    H.distance_pre = staticmethod(_dist_pre)
    H.distance_fast = staticmethod(_dist_fast)
    H.matern_cov_fast = staticmethod(_matern_fast)
    H.fast_pieces = staticmethod(_pieces)
    return H


def relerr(a, b, floor=1e-300):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0

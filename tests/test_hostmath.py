"""Numerical validation of the kernels' scalar math (csrc/ck_math.cuh, host instantiation by g++)
against scipy / the reference-generated fixtures.  Tolerance: covariance entries 1e-12 relative
(north star); the achieved agreement is ~1e-13 (scipy.special.kv's own error dominates)."""
import numpy as np
import pytest
import scipy.special as sps

import cokrig_oracle as orc
from conftest import golden, relerr

TOL_COV = 1e-12


def test_besselk_vs_scipy(hostmath):
    rng = np.random.default_rng(0)
    x = np.concatenate([10 ** rng.uniform(-10, np.log10(690), 3000), np.linspace(1.5, 2.5, 201)])
    for nu in (0.2, 0.25, 0.39, 0.5, 0.75, 0.82, 1.0, 1.25, 1.49, 1.5, 2.0, 2.3, 3.0, 3.2, 3.5):
        assert relerr(hostmath.besselk(nu, x), sps.kv(nu, x)) < 2.5e-13, nu


def test_besselk_vs_mpmath(hostmath):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    x = np.concatenate([10 ** np.random.default_rng(1).uniform(-8, 2.8, 150), [1.9, 2.0, 2.1]])
    for nu in (0.2, 0.39, 0.82, 1.25, 2.3, 3.2):
        truth = np.array([float(mp.besselk(nu, mp.mpf(float(v)))) for v in x])
        assert relerr(hostmath.besselk(nu, x), truth) < 2e-14, nu


def test_matern_vs_reference_fixture(hostmath):
    g = golden("matern")
    for a, nu in enumerate(g["nus"]):
        for b, ell in enumerate(g["lens"]):
            got = hostmath.matern_cov(1.0, nu, ell, 0.0, g["h"])
            ref = g["corr"][a, b]
            nz = ref > 0
            assert relerr(got[nz], ref[nz]) < TOL_COV, (nu, ell)
            assert ((ref == 0) == (got == 0)).all()
    p = g["params"]
    assert relerr(hostmath.matern_cov(p[0] ** 2, p[2], p[5], p[8], g["h"]), g["cov0"]) < TOL_COV
    assert relerr(hostmath.matern_cov(p[10] * p[0] * p[1], p[3], p[6], 0.0, g["h"]), g["cross01"]) < TOL_COV


def test_far_field_and_underflow_cutoff(hostmath):
    g = golden("matern")
    for k, nu in enumerate((0.5, 0.82, 1.5, 3.5)):
        got, ref = hostmath.matern_cov(1.0, nu, 0.002, 0.0, g["far_h"]), g["far_corr"][k]
        assert ((ref == 0) == (got == 0)).all(), nu  # exact zero beyond the AMOS cut-off
        nz = ref > 0
        assert relerr(got[nz], ref[nz]) < TOL_COV


def test_special_values(hostmath):
    h = np.array([0.0, -3.0, np.nan, np.inf])
    got = hostmath.matern_cov(2.0, 1.5, 10.0, 0.25, h)
    ref = 2.0 * orc.matern_correlation(1.5, 10.0, h)
    ref[h == 0] += 0.25
    np.testing.assert_allclose(got, ref, rtol=1e-15)
    assert got[0] == 2.25 and got[2] == 2.0 and got[3] == 0.0  # h==0 -> nugget; NaN -> rho=1; inf -> 0


def test_distances(hostmath):
    g = golden("distances")
    assert (hostmath.distance(0, g["Y1"], g["Y2"]) == g["euc"]).all()  # Euclidean: bit-identical
    d = hostmath.distance(1, g["X1"], g["X2"])
    assert ((d == 0) == (g["hav"] == 0)).all()
    nz = g["hav"] > 0
    assert relerr(d[nz], g["hav"][nz]) < 1e-14

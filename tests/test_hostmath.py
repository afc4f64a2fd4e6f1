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


# ------------------------------------------------------------------------------------------------
# fast assembly-path math (K1, closed-form orders): polynomial sin / asin(sqrt) / exp, FMA distances
def test_fast_pieces_vs_mpmath(hostmath):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(2)
    y = np.concatenate([rng.uniform(-np.pi / 2, np.pi / 2, 1500), 10.0 ** rng.uniform(-12, 0, 500), [np.pi / 2, -np.pi / 2, 0.0]])
    s, _, _ = hostmath.fast_pieces(y)
    truth = np.array([float(mp.sin(mp.mpf(float(v)))) for v in y])
    assert relerr(s[y != 0], truth[y != 0]) < 2.5e-16 and s[-1] == 0.0
    a = np.concatenate([rng.uniform(0, 1, 1500), 10.0 ** rng.uniform(-30, 0, 500), [0.0, 0.25, 0.75, 1.0, 0.2499999999, 0.7500000001]])
    _, t, _ = hostmath.fast_pieces(a)
    truth = np.array([float(mp.asin(mp.sqrt(mp.mpf(float(v))))) for v in a])
    assert relerr(t[a > 0], truth[a > 0]) < 4.5e-16 and t[a == 0][0] == 0.0
    x = np.concatenate([rng.uniform(0, 700, 2000), 10.0 ** rng.uniform(-12, 0, 300), [0.0, 690.0, 700.0]])
    _, _, e = hostmath.fast_pieces(x)
    truth = np.array([float(mp.exp(-mp.mpf(float(v)))) for v in x])
    assert relerr(e, truth) < 4.5e-16


def test_fast_distances_vs_reference_fixture(hostmath):
    g = golden("distances")
    de = hostmath.distance_fast(0, g["Y1"], g["Y2"])
    assert ((de == 0) == (g["euc"] == 0)).all()
    assert relerr(de[g["euc"] > 0], g["euc"][g["euc"] > 0]) < 4.5e-16
    d = hostmath.distance_fast(1, g["X1"], g["X2"])
    assert ((d == 0) == (g["hav"] == 0)).all()  # h == 0 exactly for identical points: the nugget decision
    nz = g["hav"] > 0
    assert relerr(d[nz], g["hav"][nz]) < 2e-15
    # global pairs incl. the antimeridian fold, poles and antipodes, against the reference-order function
    rng = np.random.default_rng(5)
    X = np.c_[rng.uniform(-90, 90, 300), rng.uniform(-180, 180, 300)]
    X[:4] = [[90, 0], [-90, 10], [0, 180], [0, -180]]
    a, b = hostmath.distance_fast(1, X, X), hostmath.distance(1, X, X)
    assert (np.diagonal(a) == 0).all()
    off = (b > 1e-6) & (b < 19000.0)
    assert relerr(a[off], b[off]) < 4e-15
    # within ~1000 km of the antipode asin(sqrt(a)) is ill-conditioned in a (d theta / d a = 1 / (2 sqrt(a (1 - a)))): both
    # functions carry the 1-2 ulp rounding of a, so they may differ by ~1e-16 / sqrt(1 - a) -- micrometres at most
    assert np.abs(a - b).max() < 1e-8  # km


def test_precomputed_half_angle_haversine(hostmath):
    """K1's pair distance on the sphere (ck_dist_haversine_pre: per-point half-angle sines / cosines, no sine per pair):
    exact zeros for identical points, <= 1e-11 km absolute from the reference-order function (the cancellation bound),
    and covariance entries inside the 1e-12 relative tolerance on the 0.05 degree lattice (min distance 3 km)."""
    g = golden("distances")
    d = hostmath.distance_pre(g["X1"], g["X2"])
    assert ((d == 0) == (g["hav"] == 0)).all()
    assert np.abs(d - g["hav"]).max() < 1e-11
    nz = g["hav"] > 0
    assert relerr(d[nz], g["hav"][nz]) < 2e-12
    rng = np.random.default_rng(5)
    X = np.c_[rng.uniform(-90, 90, 300), rng.uniform(-180, 180, 300)]
    X[:4] = [[90, 0], [-90, 10], [0, 180], [0, -180]]
    a, b = hostmath.distance_pre(X, X), hostmath.distance(1, X, X)
    assert (np.diagonal(a) == 0).all() and a[2, 3] < 1e-9  # (0, 180) and (0, -180) are the same point
    off = (b > 1e-6) & (b < 19000.0)
    assert (np.abs(a - b)[off] < 1e-11 + 4e-15 * b[off]).all()  # absolute (cancellation) + relative (asin amplification) parts
    # covariance entries on a dense lattice: every closed-form order within 1e-12 relative of the reference formula
    lat, lon = np.arange(30.025, 31.5, 0.05), np.arange(-100.975, -99.5, 0.05)
    L = np.array([(u, v) for u in lat for v in lon])
    ref_h = orc.distance_matrix(L, L, fast_dist=True)
    got_h = hostmath.distance_pre(L, L)
    for nu, ell in ((0.5, 100.0), (1.5, 500.0), (2.5, 300.0), (3.5, 100.0)):
        ref = orc.matern_correlation(nu, ell, ref_h)
        got = hostmath.matern_cov_fast(1.0, nu, ell, 0.0, got_h.ravel()).reshape(ref.shape)
        assert relerr(got, ref) < 1e-12, (nu, ell)


def test_fast_matern_vs_reference_fixture(hostmath):
    g = golden("matern")
    for a, nu in enumerate(g["nus"]):
        if nu not in (0.5, 1.5, 2.5, 3.5):
            continue
        for b, ell in enumerate(g["lens"]):
            got = hostmath.matern_cov_fast(1.0, nu, ell, 0.0, g["h"])
            ref = g["corr"][a, b]
            nz = ref > 0
            assert relerr(got[nz], ref[nz]) < TOL_COV, (nu, ell)
            assert ((ref == 0) == (got == 0)).all()
    for k, nu in enumerate((0.5, 0.82, 1.5, 3.5)):
        if nu == 0.82:
            continue
        got, ref = hostmath.matern_cov_fast(1.0, nu, 0.002, 0.0, g["far_h"]), g["far_corr"][k]
        assert ((ref == 0) == (got == 0)).all(), nu
        assert relerr(got[ref > 0], ref[ref > 0]) < TOL_COV
    h = np.array([0.0, -3.0, np.nan, np.inf])
    got = hostmath.matern_cov_fast(2.0, 1.5, 10.0, 0.25, h)
    assert got[0] == 2.25 and got[2] == 2.0 and got[3] == 0.0
    assert got[1] == pytest.approx(hostmath.matern_cov(2.0, 1.5, 10.0, 0.25, h)[1], rel=1e-15)


def test_besselk_chebyshev_branch_vs_mpmath(hostmath):
    """x > 2, generic nu: per-block Chebyshev expansions of sqrt(x) e^x K(x) in 2 / x fitted from the continued fraction
    (csrc/ck_matern_setup.h); segment boundaries x = 2, 4, 8 and the far tail included."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(2.0001, 4, 60), rng.uniform(4, 8, 60), rng.uniform(8, 60, 60), 10 ** rng.uniform(1.8, 2.8, 40),
                        [2.0000001, 3.9999999, 4.0, 4.0000001, 7.9999999, 8.0, 8.0000001, 690.0]])
    for nu in (0.05, 0.2, 0.39, 0.75, 0.82, 1.0, 1.25, 1.49, 2.0, 2.3, 3.0, 3.2, 7.7):
        truth = np.array([float(mp.besselk(nu, mp.mpf(float(v)))) for v in x])
        assert relerr(hostmath.besselk(nu, x), truth) < 2e-15, nu


def test_generic_nu_table_vs_reference_fixture_and_mpmath(hostmath):
    """K1's generic-nu path: piecewise Chebyshev table of rho(x) e^x on quarter-octave segments (ck_matern_corr_tab,
    fitted per block by ck_matern_table_setup) -- reference fixture (scipy kv through the unmodified reference), the
    far-field flush to exactly 0, special values, segment edges, and mpmath truth."""
    g = golden("matern")
    for a, nu in enumerate(g["nus"]):
        if nu in (0.5, 1.5, 2.5, 3.5):
            continue
        for b, ell in enumerate(g["lens"]):
            got = hostmath.matern_cov_table(1.0, nu, ell, 0.0, g["h"])
            ref = g["corr"][a, b]
            nz = ref > 0
            assert relerr(got[nz], ref[nz]) < TOL_COV, (nu, ell)
            assert ((ref == 0) == (got == 0)).all()
    got, ref = hostmath.matern_cov_table(1.0, 0.82, 0.002, 0.0, g["far_h"]), g["far_corr"][1]
    assert ((ref == 0) == (got == 0)).all()
    assert relerr(got[ref > 0], ref[ref > 0]) < TOL_COV
    h = np.array([0.0, -3.0, np.nan, np.inf])
    got = hostmath.matern_cov_table(2.0, 0.82, 10.0, 0.25, h)
    assert got[0] == 2.25 and got[2] == 2.0 and got[3] == 0.0
    assert got[1] == pytest.approx(hostmath.matern_cov(2.0, 0.82, 10.0, 0.25, h)[1], rel=1e-13)
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(11)
    edges = np.array([2.0 ** e * (1 + q / 4) for e in range(-12, 10) for q in range(4)])
    x = np.concatenate([10 ** rng.uniform(-6, 2.8, 400), edges, np.nextafter(edges, 0), np.nextafter(edges, np.inf),
                        [2.0 ** -13, 1.0e-5, 680.0]])
    for nu in (0.05, 0.2, 0.39, 0.82, 1.0, 1.25, 2.0, 2.3, 3.2, 7.7):
        sq = np.sqrt(2 * nu)
        truth = np.array([float(2 ** (1 - mp.mpf(nu)) / mp.gamma(nu) * mp.mpf(float(v)) ** nu * mp.besselk(nu, mp.mpf(float(v))))
                          for v in x])
        got = hostmath.matern_cov_table(1.0, nu, 1.0, 0.0, x / sq)  # x = sqrt(2 nu) h / l up to one rounding of h
        ok = truth > 1e-290
        assert relerr(got[ok], truth[ok]) < 1e-13, nu

"""GPU parity tests proper: the sm_100a kernels, called through the C ABI (cokrig_b200.ops -> ctypes),
against (a) fixtures produced by the unmodified reference and (b) the pinned oracle on seeded inputs.

Tolerances (BASELINE.json north star): covariance entries 1e-12 relative, kriging predictions and
variances 1e-9 relative, bin counts bit-exact.  Euclidean distances are required to be bit-exact.
"""
import numpy as np
import pytest

import cokrig_oracle as orc
from conftest import golden, relerr

pytestmark = pytest.mark.gpu

TOL_COV = 1e-12
TOL_PRED = 1e-9


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cokrig_b200 import ops as _ops
    return _ops


def dev_coords(ops, *arrs):
    return [ops.coords_to_device(a) for a in arrs]


# ------------------------------------------------------------------------------------------ K1
def test_matern_eval_vs_reference_fixture(ops):
    g = golden("matern")
    for a, nu in enumerate(g["nus"]):
        for b, ell in enumerate(g["lens"]):
            got, ref = ops.matern_eval(g["h"], 1.0, nu, ell, 0.0), g["corr"][a, b]
            assert ((ref == 0) == (got == 0)).all()
            assert relerr(got[ref > 0], ref[ref > 0]) < TOL_COV, (nu, ell)
    p = g["params"]
    assert relerr(ops.matern_eval(g["h"], p[1] ** 2, p[4], p[7], p[9]), g["cov1"]) < TOL_COV
    assert relerr(ops.matern_eval(g["h"], p[10] * p[0] * p[1], p[3], p[6], 0.0), g["cross01"]) < TOL_COV
    for k, nu in enumerate((0.5, 0.82, 1.5, 3.5)):
        got, ref = ops.matern_eval(g["far_h"], 1.0, nu, 0.002, 0.0), g["far_corr"][k]
        assert ((ref == 0) == (got == 0)).all(), nu
        assert relerr(got[ref > 0], ref[ref > 0]) < TOL_COV


def test_matern_eval_shapes_and_special_values(ops):
    h = np.array([[0.0, -3.0], [np.nan, np.inf]])
    got = ops.matern_eval(h, 2.0, 1.5, 10.0, 0.25)
    assert got.shape == (2, 2)
    assert got[0, 0] == 2.25 and got[1, 0] == 2.0 and got[1, 1] == 0.0
    assert got[0, 1] == pytest.approx(2.0 * orc.matern_correlation(1.5, 10.0, 3.0)[0], rel=1e-14)
    assert ops.matern_eval(np.zeros(0), 1.0, 1.5, 1.0).shape == (0,)
    assert ops.matern_eval(0.0, 1.0, 1.5, 1.0, 0.5).shape == (1,)  # scalar h -> length-1 array (point_prediction.py:66)


def test_distance_blocks(ops):
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE
    g = golden("distances")
    e = ops.distance_block(*dev_coords(ops, g["Y1"], g["Y2"]), METRIC_EUCLID).cpu().numpy()
    assert (e == g["euc"]).all(), "Euclidean distances must be bit-identical to scipy.cdist"
    d = ops.distance_block(*dev_coords(ops, g["X1"], g["X2"]), METRIC_HAVERSINE).cpu().numpy()
    assert ((d == 0) == (g["hav"] == 0)).all()
    assert relerr(d[g["hav"] > 0], g["hav"][g["hav"] > 0]) < 1e-14


@pytest.mark.parametrize("name,metric,i", [("joint_euclid", "euclidean", 1), ("joint_haversine_generic", "haversine", 0),
                                           ("joint_haversine_half", "haversine", 0)])
def test_joint_and_cross_cov_vs_reference_fixture(ops, name, metric, i):
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE
    mid = METRIC_EUCLID if metric == "euclidean" else METRIC_HAVERSINE
    g = golden(name)
    cd = dev_coords(ops, g["coords0"], g["coords1"])
    sigma = ops.joint_cov(cd, g["params"], 2, mid).cpu().numpy()
    assert relerr(sigma, g["sigma"]) < TOL_COV
    assert (sigma == sigma.T).all(), "mirrored tiles must be bit-symmetric"
    cpd = ops.cross_cov(cd, ops.coords_to_device(g["pcoords"]), g["params"], 2, i, mid, spare_rows=0).cpu().numpy()
    assert relerr(cpd, g["c_dp"].T) < TOL_COV


def test_matern_block_ragged_and_empty(ops):
    from cokrig_b200 import METRIC_EUCLID
    rng = np.random.default_rng(5)
    P = orc.Params([1.3, 0.7, 0.5, 2.5, 3.5, .3, .4, .5, .01, .02, .4])
    for n1, n2 in [(1, 1), (1, 130), (63, 65), (64, 64), (129, 200), (257, 31)]:
        a, b = rng.uniform(0, 1, (n1, 2)), rng.uniform(0, 1, (n2, 2))
        b[0] = a[0]  # one coincident pair: nugget at h == 0 off the diagonal
        ref = orc.covariance(P, 1, orc.distance_matrix(a, b, units=None))
        got = ops.matern_block(*dev_coords(ops, a, b), METRIC_EUCLID, 0.7 ** 2, 3.5, .5, .02).cpu().numpy()
        assert relerr(got, ref) < TOL_COV, (n1, n2)
        assert got[0, 0] == pytest.approx(0.49 + 0.02, rel=1e-15)
    import torch
    empty = torch.empty((0, 2), dtype=torch.float64, device="cuda")
    assert ops.matern_block(empty, empty, METRIC_EUCLID, 1.0, 1.5, 1.0).shape == (0, 0)
    S = ops.joint_cov([ops.coords_to_device(rng.uniform(0, 1, (5, 2)))], [1.0, 1.5, 0.3, 0.1], 1, METRIC_EUCLID)
    assert S.shape == (5, 5) and S[0, 0].item() == pytest.approx(1.1)


# ------------------------------------------------------------------------------------------ K3
def _spd(n, seed, nugget=0.01):
    xy = np.random.default_rng(seed).uniform(0, 1, (n, 2))
    return np.exp(-orc.distance_matrix(xy, xy, units=None) / 0.2) + nugget * np.eye(n)


@pytest.mark.parametrize("n", [1, 7, 100, 128, 129, 257, 1000, 2049])
def test_potrf_and_trsm_vs_lapack(ops, n):
    import torch
    from scipy.linalg import cholesky, solve_triangular
    A = _spd(n, n)
    Lref = cholesky(A, lower=True)
    buf = torch.empty((n, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]
    buf.copy_(torch.from_numpy(A))
    f = ops.potrf(buf)
    assert f.info == 0
    L = f.lower().cpu().numpy()
    assert np.abs(L - Lref).max() / np.abs(Lref).max() < 1e-12
    assert np.abs(L @ L.T - A).max() < 1e-13 * n
    B = np.random.default_rng(n + 1).standard_normal((5, n))
    rb = torch.empty((5, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]
    rb.copy_(torch.from_numpy(B))
    V = f.solve_lower(rb).cpu().numpy()
    Vref = solve_triangular(Lref, B.T, lower=True).T
    assert np.abs(V - Vref).max() / np.abs(Vref).max() < 1e-11
    assert abs(f.logdet().item() - 2 * np.log(np.diagonal(Lref)).sum()) < 1e-10 * max(n, 1)


def test_potrf_unaligned_leading_dimension(ops):
    import torch
    n = 301  # odd ld: 8-byte cp.async path
    A = _spd(n, 9)
    f = ops.potrf(torch.from_numpy(A.copy()).cuda())
    L = f.lower().cpu().numpy()
    assert f.info == 0 and np.abs(L @ L.T - A).max() < 1e-12


def test_potrf_reports_first_bad_minor(ops):
    import torch
    from scipy.linalg import LinAlgError
    A = _spd(300, 3)
    A[200, 200] = -1.0
    f = ops.potrf(torch.from_numpy(A).cuda())
    assert f.info == 201  # LAPACK dpotrf: order of the first non-PD leading minor
    with pytest.raises(LinAlgError):
        f.raise_if_failed()
    f0 = ops.potrf(torch.empty((0, 0), dtype=torch.float64, device="cuda"))
    assert f0.info == 0


def test_potrf_leaves_upper_triangle_untouched(ops):
    import torch
    n = 200
    A = _spd(n, 2)
    marked = np.tril(A) + np.triu(np.full((n, n), 77.0), 1)
    out = ops.potrf(torch.from_numpy(marked).cuda()).L.cpu().numpy()
    assert (out[np.triu_indices(n, 1)] == 77.0).all()


@pytest.mark.parametrize("name,metric,i,n_procs", [("joint_euclid", "euclidean", 1, 2), ("joint_haversine_generic", "haversine", 0, 2),
                                                   ("joint_haversine_half", "haversine", 0, 2), ("joint_univariate", "euclidean", 0, 1)])
def test_joint_prediction_vs_reference_fixture(ops, name, metric, i, n_procs):
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE
    mid = METRIC_EUCLID if metric == "euclidean" else METRIC_HAVERSINE
    g = golden(name)
    coords = [g["coords0"]] + ([g["coords1"]] if n_procs == 2 else [])
    z = np.hstack([g["z0"]] + ([g["z1"]] if n_procs == 2 else []))
    cd = dev_coords(ops, *coords)
    f = ops.potrf(ops.joint_cov(cd, g["params"], n_procs, mid))
    cpd = ops.cross_cov(cd, ops.coords_to_device(g["pcoords"]), g["params"], n_procs, i, mid)
    P = orc.Params(g["params"], n_procs)
    c0 = P.sigma[i, i] ** 2 + P.nugget[i, i]
    pred, var = f.predict(cpd, ops.to_device(z), c0)
    assert f.info == 0
    assert relerr(pred.cpu().numpy(), g["pred"]) < TOL_PRED
    err = np.nan_to_num(np.sqrt(var.cpu().numpy()))
    # variances cancel from c0 ~ 1 down to ~1e-2..1e-6 at data locations: compare relative to c0
    assert np.abs(err ** 2 - g["pred_err"] ** 2).max() / c0 < TOL_PRED
    far = g["pred_err"] > 1e-3
    assert relerr(err[far], g["pred_err"][far]) < TOL_PRED


def test_gaussian_nll_vs_oracle(ops):
    from cokrig_b200 import METRIC_HAVERSINE
    g = golden("joint_haversine_generic")
    P = orc.Params(g["params"])
    ref = orc.gaussian_nll(P, [g["coords0"], g["coords1"]], [g["z0"], g["z1"]], "haversine")
    out, info = ops.gaussian_nll(dev_coords(ops, g["coords0"], g["coords1"]), ops.to_device(np.hstack([g["z0"], g["z1"]])),
                                 g["params"], 2, METRIC_HAVERSINE)
    assert int(info.item()) == 0 and abs(out[0].item() / ref - 1) < TOL_PRED


def test_cholesky_properties_at_scale(ops):
    """Size-independent checks at N = 8192 (oracle too slow / not needed): solve residual and
    reconstruction on sampled entries, symmetric assembly, determinism."""
    import torch
    from cokrig_b200 import METRIC_EUCLID
    n = 4096
    xy = ops.coords_to_device(np.random.default_rng(4).uniform(0, 1, (n, 2)))
    params = [1, .8, 1.5, 1.5, 1.5, .05, .05, .05, .02, .02, -.2]
    S = ops.joint_cov([xy, xy], params, 2, METRIC_EUCLID)
    S0 = S.clone()
    f = ops.potrf(S)
    assert f.info == 0
    L = f.lower()
    idx = torch.randint(0, 2 * n, (200,), device="cuda")
    recon = (L[idx] @ L.T)  # 200 sampled rows of L L^T
    assert (recon - S0[idx]).abs().max().item() < 1e-11
    z = torch.randn(2 * n, dtype=torch.float64, device="cuda")
    y = f.solve_lower(z.clone().reshape(1, -1))
    assert (L @ y.reshape(-1) - z).abs().max().item() < 1e-9
    S2 = S0.clone()
    f2 = ops.potrf(S2)
    assert torch.equal(torch.tril(f2.L), L), "factorisation must be bit-reproducible"


# ------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("name,metric,cov", [("variogram_haversine_semivariogram", "haversine", False),
                                             ("variogram_haversine_covariogram", "haversine", True),
                                             ("variogram_euclid", "euclidean", False)])
def test_variogram_vs_reference_fixture(ops, name, metric, cov):
    import fields
    from cokrig_b200 import METRIC_EUCLID, METRIC_HAVERSINE
    mid = METRIC_EUCLID if metric == "euclidean" else METRIC_HAVERSINE
    g = golden(name)
    coords, values = [g["coords0"], g["coords1"]], [g["v0"], g["v1"]]
    for (i, j) in ((0, 0), (0, 1), (1, 1)):
        centers, edges, counts, sums = fields._device_variogram(coords[i], values[i], coords[j], values[j], i == j, mid,
                                                                cov, float(g["max_dist"]), int(g["n_bins"]))
        np.testing.assert_array_equal(centers, g[f"center{i}{j}"])           # bit-exact bin centres
        np.testing.assert_array_equal(counts, g[f"count{i}{j}"])             # bit-exact counts
        ref = g[f"mean{i}{j}"]
        ok = counts > 0
        assert relerr((sums[ok] / counts[ok]), ref[ok]) < 1e-12
        assert np.isnan(ref[~ok]).all()


def test_variogram_vs_oracle_larger_and_reproducible(ops):
    import fields
    from cokrig_b200 import METRIC_HAVERSINE
    lat, lon = np.arange(22.025, 58, 0.05), np.arange(-124.975, -65, 0.05)

    def draw(seed, n):
        idx = np.random.default_rng(seed).choice(len(lat) * len(lon), n, replace=False)
        return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]
    ca, cb = draw(2, 2500), draw(3, 2300)
    va, vb = np.random.default_rng(5).standard_normal(2500), np.random.default_rng(6).standard_normal(2300)
    ref = orc.empirical_variograms([ca, cb], [va, vb], 1500.0, 50, "haversine")
    for (i, j) in ((0, 0), (0, 1)):
        A, B = (ca, ca) if i == j else (ca, cb)
        a, b = (va, va) if i == j else (va, vb)
        r1 = fields._device_variogram(A, a, B, b, i == j, METRIC_HAVERSINE, False, 1500.0, 50)
        r2 = fields._device_variogram(A, a, B, b, i == j, METRIC_HAVERSINE, False, 1500.0, 50)
        d = ref.loc[(i, j)]
        np.testing.assert_array_equal(r1[0], d["bin_center"].values)
        np.testing.assert_array_equal(r1[2], d["bin_count"].values)
        assert relerr(r1[3] / r1[2], d["bin_mean"].values) < 1e-12
        assert (r1[3] == r2[3]).all(), "FP64 bin sums must be bit-reproducible"
        assert r1[2].sum() == (orc.variogram_cloud([ca, cb], [va, vb], i, j, "haversine")[0] <= 1500.0).sum()


def test_variogram_guard_band_pairs_are_resolved_on_host(ops):
    """Force pairs exactly onto max_dist: they must be returned as flagged pairs, not decided on device."""
    from cokrig_b200 import METRIC_HAVERSINE
    ca = np.array([[30.0, -100.0], [30.0, -99.0], [31.0, -100.0], [35.0, -90.0]])
    d = orc.distance_matrix(ca, ca, fast_dist=True)
    X = ops.coords_to_device(ca)
    v = ops.to_device(np.arange(4.0))
    edges = np.array([0.0, d[0, 1] / 2, d[0, 1], 5000.0])  # an edge exactly on a pair distance
    counts, sums, flagged = ops.vario_bin(X, v, 1.5, X, v, 1.5, METRIC_HAVERSINE, True, False, 1e6, edges)
    assert [tuple(p) for p in flagged] == [(0, 1)]
    assert counts.sum() == 5  # the sixth pair was left to the host


# ------------------------------------------------------------------------------------------ K4
def test_local_prediction_vs_reference_fixture(ops):
    from cokrig_b200 import METRIC_EUCLID
    g = golden("point_euclid")
    cd = dev_coords(ops, g["coords0"], g["coords1"])
    zz = [ops.to_device(g["z0"]), ops.to_device(g["z1"])]
    pred, sd, k, info = ops.local_predict(cd, zz, ops.coords_to_device(g["pcoords"]), g["params"], 2, 1, METRIC_EUCLID,
                                          float(g["max_dist"]))
    np.testing.assert_array_equal(k, g["k"])                      # neighbour counts bit-exact
    assert k[-1] == 0 and np.isnan(pred[-1]) and np.isnan(sd[-1])  # empty neighbourhood -> (nan, nan)
    assert relerr(pred[:-1], g["pred"][:-1]) < TOL_PRED
    assert np.abs(sd[:-1] ** 2 - g["sd"][:-1] ** 2).max() < TOL_PRED
    pred, sd, k, info = ops.local_predict(cd, zz, ops.coords_to_device(g["cv_pcoords"]), g["params"], 2, 1, METRIC_EUCLID,
                                          float(g["cv_max_dist"]), cv=True)
    assert relerr(pred, g["cv_pred"]) < TOL_PRED and relerr(sd, g["cv_sd"]) < TOL_PRED


def test_local_prediction_vs_oracle_haversine_generic(ops):
    from cokrig_b200 import METRIC_HAVERSINE
    g = golden("joint_haversine_generic")
    P = orc.Params(g["params"])
    cm, zz = [g["coords0"], g["coords1"]], [g["z0"], g["z1"]]
    pr, sd, k, _ = orc.point_predict(P, 0, cm, zz, g["pcoords"], 800.0, "haversine")
    gp, gs, gk, info = ops.local_predict(dev_coords(ops, *cm), [ops.to_device(z) for z in zz], ops.coords_to_device(g["pcoords"]),
                                         g["params"], 2, 0, METRIC_HAVERSINE, 800.0)
    np.testing.assert_array_equal(gk, k)
    ok = ~np.isnan(pr)
    assert (np.isnan(gp) == ~ok).all()
    assert relerr(gp[ok], pr[ok]) < TOL_PRED and np.abs(gs[ok] ** 2 - sd[ok] ** 2).max() < TOL_PRED


def test_local_prediction_non_pd_returns_nan(ops):
    from cokrig_b200 import METRIC_EUCLID
    xy = np.array([[0.0, 0.0], [0.0, 0.0], [0.1, 0.0]])  # duplicated point, zero nugget -> singular local matrix
    cd = dev_coords(ops, xy, xy[:1] + 5)
    zz = [ops.to_device(np.array([1.0, 2.0, 3.0])), ops.to_device(np.array([0.0]))]
    pred, sd, k, info = ops.local_predict(cd, zz, ops.coords_to_device(np.array([[0.05, 0.0]])),
                                          [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, 0, 0, .5], 2, 0, METRIC_EUCLID, 1.0)
    assert k[0] == 3 and info[0] > 0 and np.isnan(pred[0]) and np.isnan(sd[0])

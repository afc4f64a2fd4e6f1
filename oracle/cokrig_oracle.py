"""CPU oracle: numpy/scipy restatement of the reference's cokriging hot path.

TEST INFRASTRUCTURE ONLY.  Nothing on the product path may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the timed CPU baseline.

The reference (91Mrwu/sif-xco2-cokriging) is pure Python; its arithmetic lives in
un-pinned third-party libraries (environment.yaml:5-27).  This restatement keeps
the reference's own glue, operation order and quirks and calls the same
third-party routines (versions in this image): ``scipy.special.kv/gammaln``
1.18.1, ``scipy.linalg.cho_factor/cho_solve/cholesky`` (OpenBLAS LAPACK),
``sklearn.metrics.pairwise.haversine_distances`` 1.9.0,
``scipy.spatial.distance.cdist``, ``pandas.cut`` 3.0.2, ``scipy.optimize``.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against
fixtures produced by the UNMODIFIED reference (``tests/golden/make_golden.py``,
run in the build container where /root/reference exists) and against the two
notebook print-outs the reference holds (research/simulation_experiment.ipynb
cells [11] and [16]).  ``gaussian_nll`` and ``loocv_closed_form`` have no
reference counterpart in the current ``src/``: parity unpinned
for those two.

All ``file:line`` citations are relative to /root/reference/.
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
import scipy.special as sps
from scipy.linalg import LinAlgError, cho_factor, cho_solve, cholesky
from scipy.spatial.distance import cdist

EARTH_RADIUS = 6371  # src/fields.py:17


# --------------------------------------------------------------------------- params
class Params:
    """Flat parameter vector -> the reference's n_procs x n_procs matrices.

    Order (src/model.py:130-152): sigma_ii, nu_ij (i<=j, row-major), len_scale_ij,
    nugget_ii, rho_ij (i<j).  11 values for n_procs=2, 4 for n_procs=1.
    """

    def __init__(self, values, n_procs: int = 2):
        v = np.asarray(values, dtype=float)
        n = n_procs
        n_tri = n * (n + 1) // 2
        n_rho = n * (n - 1) // 2
        if v.size != 2 * n + 2 * n_tri + n_rho:
            raise ValueError("Incorrect number of parameters in input array.")  # src/model.py:146-147
        self.n_procs = n
        iu = np.triu_indices(n)
        iu1 = np.triu_indices(n, k=1)
        self.sigma = np.full((n, n), np.nan)
        self.nu = np.full((n, n), np.nan)
        self.len_scale = np.full((n, n), np.nan)
        self.nugget = np.full((n, n), np.nan)
        self.rho = np.full((n, n), np.nan)
        o = 0
        np.fill_diagonal(self.sigma, v[o:o + n]); o += n
        self.nu[iu] = v[o:o + n_tri]; o += n_tri
        self.len_scale[iu] = v[o:o + n_tri]; o += n_tri
        np.fill_diagonal(self.nugget, v[o:o + n]); o += n
        # RhoParam: diagonal keeps the default (0.0, or nan for n_procs=1), src/model.py:91-98
        np.fill_diagonal(self.rho, 0.0 if n > 1 else np.nan)
        self.rho[iu1] = v[o:o + n_rho]
        self.values = v.copy()


# --------------------------------------------------------------------------- Matern
def matern_correlation(nu: float, len_scale: float, h) -> np.ndarray:
    """src/model.py:354-385 (log-form prefactor times scipy.special.kv)."""
    h = np.atleast_1d(np.abs(h))
    hs = h[h > 0.0] / len_scale
    corr = np.ones_like(h, dtype=float)
    with np.errstate(all="ignore"):
        corr[h > 0.0] = np.exp(
            (1.0 - nu) * np.log(2) - sps.gammaln(nu) + nu * np.log(np.sqrt(2.0 * nu) * hs)
        ) * sps.kv(nu, np.sqrt(2.0 * nu) * hs)
    corr[np.logical_not(np.isfinite(corr))] = 0.0
    return np.maximum(corr, 0.0)


def covariance(p: Params, i: int, h, use_nugget: bool = True) -> np.ndarray:
    """src/model.py:193-197 (nugget wherever h == 0 exactly)."""
    cov = p.sigma[i, i] ** 2 * matern_correlation(p.nu[i, i], p.len_scale[i, i], h)
    if use_nugget:
        cov[np.atleast_1d(h) == 0] += p.nugget[i, i]
    return cov


def cross_covariance(p: Params, i: int, j: int, h) -> np.ndarray:
    """src/model.py:199-207 (nanprod of ALL sigmas; indices swapped if i > j)."""
    if i > j:
        i, j = j, i
    return p.rho[i, j] * np.nanprod(p.sigma) * matern_correlation(p.nu[i, j], p.len_scale[i, j], h)


def semivariance(p: Params, i: int, h) -> np.ndarray:
    """src/model.py:209-213."""
    return p.sigma[i, i] ** 2 * (1.0 - matern_correlation(p.nu[i, i], p.len_scale[i, i], h)) + p.nugget[i, i]


def cross_semivariance(p: Params, i: int, j: int, h) -> np.ndarray:
    """src/model.py:215-222."""
    if i > j:
        i, j = j, i
    sill = 0.5 * np.nansum(p.sigma ** 2 + p.nugget)
    return sill - cross_covariance(p, i, j, h)


# --------------------------------------------------------------------------- distances
def distance_matrix(X1, X2, units="km", fast_dist=False) -> np.ndarray:
    """src/fields.py:318-342.  Rows are [lat, lon] degrees (haversine) or [x, y]."""
    X1 = np.atleast_2d(X1)
    X2 = np.atleast_2d(X2)
    if fast_dist:
        from sklearn.metrics.pairwise import haversine_distances
        return haversine_distances(np.radians(X1), np.radians(X2)) * EARTH_RADIUS
    elif units is not None:
        raise NotImplementedError("geodesic distances (geopy callback) are outside the hot path")
    return cdist(X1, X2)


def _metric_args(metric: str):
    if metric == "haversine":
        return dict(units="km", fast_dist=True)
    if metric == "euclidean":
        return dict(units=None, fast_dist=False)
    raise ValueError(metric)


# --------------------------------------------------------------------------- variogram
def cloud_calc(values_a, values_b, covariogram: bool) -> np.ndarray:
    """src/fields.py:378-386 (residuals about the mean of ALL values of each field)."""
    ra = values_a - values_a.mean()
    rb = values_b - values_b.mean()
    if covariogram:
        return np.multiply.outer(ra, rb)
    return 0.5 * (np.subtract.outer(ra, rb)) ** 2


def variogram_cloud(coords, values, i, j, metric, covariogram=False):
    """src/fields.py:192-206 -> (distance, cloud) 1-d arrays in the reference's pair order."""
    dist = distance_matrix(coords[i], coords[j], **_metric_args(metric))
    if i == j:
        idx = np.triu_indices(dist.shape[0], k=1, m=dist.shape[1])
        return dist[idx], cloud_calc(values[i], values[i], covariogram)[idx]
    return dist.flatten(), cloud_calc(values[i], values[j], covariogram).flatten()


def construct_variogram_bins(distance: np.ndarray, n_bins: int):
    """src/fields.py:389-403."""
    min_dist = distance[distance > 0].min()
    max_dist = distance.max()
    centers = np.linspace(min_dist, max_dist, n_bins)
    width = centers[1] - centers[0]
    edges = np.arange(min_dist - 0.5 * width, max_dist + width, width)
    if not np.allclose((edges[1:] + edges[:-1]) / 2, centers):
        warnings.warn("WARNING: variogram bins are not centered.")
    edges[0] = 0
    return centers, edges


def get_variogram(coords, values, i, j, max_dist, n_bins, metric="haversine", covariogram=False) -> pd.DataFrame:
    """src/fields.py:208-232, emitting all n_bins rows (reference-era pandas kept
    empty categories: bin_mean NaN, bin_count 0)."""
    d, c = variogram_cloud(coords, values, i, j, metric, covariogram)
    keep = d <= max_dist
    d, c = d[keep], c[keep]
    centers, edges = construct_variogram_bins(d, n_bins)
    cat = pd.cut(pd.Series(d), edges, labels=centers, include_lowest=True)
    g = pd.DataFrame({"bin_center": cat, "variogram": c}).groupby("bin_center", observed=False)["variogram"]
    df = g.agg(["mean", "count"]).rename(columns={"mean": "bin_mean", "count": "bin_count"}).reset_index()
    df["bin_center"] = df["bin_center"].astype("string").astype("float")
    if (df["bin_count"] < 30).any():
        warnings.warn("WARNING: Fewer than 30 pairs used for at least one bin in variogram calculation.")
    df["i"], df["j"] = i, j
    return df.set_index(["i", "j", df.index])


def empirical_variograms(coords, values, max_dist, n_bins, metric="haversine", covariogram=False) -> pd.DataFrame:
    """src/fields.py:234-252."""
    n = len(coords)
    return pd.concat([get_variogram(coords, values, i, j, max_dist, n_bins, metric, covariogram)
                      for i in range(n) for j in range(n) if i <= j])


# --------------------------------------------------------------------------- joint cokriging
def joint_cov(p: Params, coords_main, metric, cv=None) -> np.ndarray:
    """src/joint_prediction.py:124-153.  cv = (i, ix) deletes row/col ix of process i."""
    n = p.n_procs
    blocks = {}
    for i in range(n):
        for j in range(n):
            if i <= j:
                d = distance_matrix(coords_main[i], coords_main[j], **_metric_args(metric))
                blocks[i, j] = covariance(p, i, d) if i == j else cross_covariance(p, i, j, d)
            else:
                blocks[i, j] = blocks[j, i].T.copy()
    if cv is not None:
        ci, ix = cv
        for i in range(n):
            for j in range(n):
                if i == ci:
                    blocks[i, j] = np.delete(blocks[i, j], ix, axis=0)
                if j == ci:
                    blocks[i, j] = np.delete(blocks[i, j], ix, axis=1)
    return np.block([[blocks[i, j] for j in range(n)] for i in range(n)])


def pred_cov(p: Params, i: int, pcoords, metric) -> np.ndarray:
    """src/joint_prediction.py:94-102."""
    d = distance_matrix(pcoords, pcoords, **_metric_args(metric))
    return covariance(p, i, d, use_nugget=True)


def pred_cross_cov(p: Params, i: int, coords_main, pcoords, metric, cv_ix=None) -> np.ndarray:
    """src/joint_prediction.py:104-122 -> (N x m), rows stacked process 0 then 1."""
    dists = [distance_matrix(c, pcoords, **_metric_args(metric)) for c in coords_main]
    if cv_ix is not None:
        dists[i] = np.delete(dists[i], cv_ix, axis=0)
    vecs = []
    for j in range(p.n_procs):
        if i == j:
            vecs.append(covariance(p, i, dists[i], use_nugget=True))
        else:
            vecs.append(cross_covariance(p, i, j, dists[j]))
    return np.vstack(vecs)


def joint_predict(p: Params, i: int, coords_main, values_main, pcoords, metric, cv_ix=None):
    """src/joint_prediction.py:50-78 -> (pred, pred_err, model_valid)."""
    pcoords = np.atleast_2d(pcoords)
    c_pp = pred_cov(p, i, pcoords, metric)
    c_dp = pred_cross_cov(p, i, coords_main, pcoords, metric, cv_ix=cv_ix)
    sigma = joint_cov(p, coords_main, metric, cv=None if cv_ix is None else (i, cv_ix))
    data = [np.asarray(v, float).copy() for v in values_main]
    valid = True
    if cv_ix is not None:
        data[i] = np.delete(data[i], cv_ix, axis=0)
    else:
        try:  # src/joint_prediction.py:260-274
            cho_factor(np.vstack([np.hstack([c_pp, c_dp.T]), np.hstack([c_dp, sigma])]))
        except LinAlgError:
            valid = False
    z = np.hstack(data)
    w = cho_solve(cho_factor(sigma, lower=True, check_finite=False), c_dp.copy(), check_finite=False).T
    var = np.diagonal(c_pp - np.matmul(w, c_dp))
    with np.errstate(invalid="ignore"):
        return np.matmul(w, z), np.nan_to_num(np.sqrt(var)), valid


# --------------------------------------------------------------------------- point cokriging
def cov_blocks(p: Params, coords_main, metric) -> dict:
    """src/point_prediction.py:98-113 (upper-triangular block dict keyed '00','01','11')."""
    blocks = {}
    for i in range(p.n_procs):
        for j in range(i, p.n_procs):
            d = distance_matrix(coords_main[i], coords_main[j], **_metric_args(metric))
            blocks[f"{i}{j}"] = covariance(p, i, d) if i == j else cross_covariance(p, i, j, d)
    return blocks


def local_prediction(p: Params, i: int, blocks, coords_main, values_main, s0, max_dist, metric, cv=False):
    """src/point_prediction.py:127-241 for one target -> (pred, sd, k, valid)."""
    n = p.n_procs
    c0 = covariance(p, i, 0, use_nugget=True)[0]  # src/point_prediction.py:66
    dists = [distance_matrix(s0, c, **_metric_args(metric)) for c in coords_main]
    conds = [(d <= max_dist) for d in dists]
    if cv:
        conds[i] = (dists[i] > 0) & (dists[i] <= max_dist)
    ix = [c.squeeze(axis=0) for c in conds]
    ldist = [d[c] for d, c in zip(dists, conds)]
    z = np.hstack([np.asarray(values_main[k])[ix[k]] for k in range(n)])
    if z.size == 0:
        return np.nan, np.nan, 0, True
    cvec = np.hstack([covariance(p, i, ldist[k], use_nugget=True) if k == i
                      else cross_covariance(p, i, k, ldist[k]) for k in range(n)])
    lb = {}
    for a in range(n):
        for b in range(n):
            lb[a, b] = blocks[f"{a}{b}"][np.ix_(ix[a], ix[b])] if a <= b else blocks[f"{b}{a}"][np.ix_(ix[b], ix[a])].T
    lcov = np.block([[lb[a, b] for b in range(n)] for a in range(n)])
    valid = True
    try:  # src/point_prediction.py:183-198
        cho_factor(np.vstack([np.hstack([c0, cvec]), np.column_stack([cvec, lcov])]))
    except LinAlgError:
        valid = False
    try:  # src/point_prediction.py:200-222
        w = cho_solve(cho_factor(lcov, lower=True, check_finite=False), cvec.copy(), check_finite=False).T
        pred = np.matmul(w, z)
        with np.errstate(invalid="ignore"):
            sd = np.sqrt(c0 - np.matmul(w, cvec))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return pred, np.nanmax([sd, 0.0]), z.size, valid
    except LinAlgError:
        return np.nan, np.nan, z.size, valid


def point_predict(p: Params, i: int, coords_main, values_main, pcoords, max_dist, metric, cv=False):
    """src/point_prediction.py:243-249 -> (pred, sd, k, valid) arrays over targets."""
    blocks = cov_blocks(p, coords_main, metric)
    out = [local_prediction(p, i, blocks, coords_main, values_main, s0, max_dist, metric, cv)
           for s0 in np.atleast_2d(pcoords)]
    pred, sd, k, valid = map(np.array, zip(*out))
    return pred, sd, k, valid


# --------------------------------------------------------------------------- simulation
def expand_grid(xbounds=(0, 1), ybounds=(0, 1), xcount=51, ycount=51) -> np.ndarray:
    """src/sim.py:14-31 -> (xcount*ycount, 2) coordinates [x, y]."""
    xs = np.linspace(*xbounds, num=xcount)
    ys = np.linspace(*ybounds, num=ycount)
    return np.array(np.meshgrid(xs, ys)).T.reshape(-1, 2)


def sim_joint_cov(p: Params, coords) -> np.ndarray:
    """src/sim.py:45-50."""
    d = cdist(coords, coords)
    c11 = covariance(p, 0, d)
    c22 = covariance(p, 1, d)
    c12 = cross_covariance(p, 0, 1, d)
    return np.block([[c11, c12], [c12.T, c22]])


def sim_fields(p: Params, coords, seed):
    """src/sim.py:33-65 -> (cmat, L, [field0, field1])."""
    rng = np.random.default_rng(seed)
    cmat = sim_joint_cov(p, coords)
    low = cholesky(cmat, lower=True)
    noise = rng.standard_normal(2 * len(coords))
    s = low @ noise
    n = len(coords)
    return cmat, low, [s[:n], s[n:]]


# --------------------------------------------------------------------------- WLS fit
def wls(ydata, yfit, counts) -> float:
    """src/model.py:388-391."""
    return float(np.sum(counts * ((ydata - yfit) / yfit) ** 2))


def composite_wls(values, df_vario: pd.DataFrame, n_procs: int = 2) -> float:
    """src/model.py:277-283 (rows whose model value is exactly 0 are dropped)."""
    p = Params(values, n_procs)
    fit = np.empty(len(df_vario))
    ii = df_vario.index.get_level_values(0).values
    jj = df_vario.index.get_level_values(1).values
    h = df_vario["bin_center"].values
    for i in range(n_procs):
        for j in range(i, n_procs):
            m = (ii == i) & (jj == j)
            if m.any():
                fit[m] = semivariance(p, i, h[m]) if i == j else cross_semivariance(p, i, j, h[m])
    y = df_vario["bin_mean"].values
    c = df_vario["bin_count"].values
    nz = fit != 0.0
    return wls(y[nz], fit[nz], c[nz])


# --------------------------------------------------------------------------- new capabilities (parity unpinned)
def gaussian_nll(p: Params, coords_main, values_main, metric) -> float:
    """0.5*(z' S^-1 z + logdet S + N log 2pi) -- no counterpart in current src/ (SURVEY.md 0.2)."""
    s = joint_cov(p, coords_main, metric)
    z = np.hstack(values_main)
    low = cholesky(s, lower=True)
    from scipy.linalg import solve_triangular
    y = solve_triangular(low, z, lower=True)
    return float(0.5 * (y @ y) + np.log(np.diagonal(low)).sum() + 0.5 * z.size * np.log(2 * np.pi))


def loocv_reference(p: Params, i: int, coords_main, values_main, metric):
    """src/joint_prediction.py:207-257 core: one full re-assembly + refactor per held-out datum."""
    preds, sds = [], []
    for ix in range(len(values_main[i])):
        pr, sd, _ = joint_predict(p, i, coords_main, values_main, coords_main[i][ix], metric, cv_ix=ix)
        preds.append(pr[0]); sds.append(sd[0])
    return np.array(preds), np.array(sds)


# --------------------------------------------------------------------------- CPU baseline helper
def joint_predict_phases(p: Params, i: int, coords_main, values_main, pcoords, metric) -> dict:
    """src/joint_prediction.py:50-78 executed phase by phase with wall-clock times (bench.py's
    cpu_baseline / --impl reference legs).  Same calls as ``joint_predict``."""
    import time
    t = {}
    t0 = time.perf_counter()
    c_pp = pred_cov(p, i, pcoords, metric)
    c_dp = pred_cross_cov(p, i, coords_main, pcoords, metric)
    sigma = joint_cov(p, coords_main, metric)
    t["assemble_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    try:
        cho_factor(np.vstack([np.hstack([c_pp, c_dp.T]), np.hstack([c_dp, sigma])]), overwrite_a=True)
    except LinAlgError:
        pass
    t["verify_s"] = time.perf_counter() - t0
    z = np.hstack(values_main)
    t0 = time.perf_counter()
    fac = cho_factor(sigma, lower=True, overwrite_a=True, check_finite=False)
    t["factor_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    w = cho_solve(fac, c_dp.copy(), overwrite_b=True, check_finite=False).T
    var = np.diagonal(c_pp - np.matmul(w, c_dp))
    pred = np.matmul(w, z)
    with np.errstate(invalid="ignore"):
        err = np.nan_to_num(np.sqrt(var))
    t["solve_s"] = time.perf_counter() - t0
    t["pred"], t["pred_err"] = pred, err
    return t

"""Round-2 fixtures: the UNMODIFIED reference run at the BASELINE configurations' real sizes.

    python tests/golden/make_golden_r2.py        (build container only: needs /root/reference; ~5 minutes, ~8 GB)

  at_size_c1.npz   C1: sim.BivariateRandomField on the 40 x 40 grid (seed 1, n = 1 600 per variable), 500 targets
                   default_rng(7).uniform(0, 1, (500, 2)), nugget 0.01 (parity run) and 0 (notebook-faithful):
                   point_prediction.Predictor._predict_chunk at max_dist = 0.2 (k = 142..390 neighbours per target, mean 329) and
                   the joint predictor's core (src/joint_prediction.py:50-78), plus an LU solve of the same joint system
                   = the parity floor between two valid FP64 solvers.
  at_size_c2.npz   C2: 10 000 cells per variable from the 0.05 degree CONUS lattice (default_rng(2) / default_rng(3)),
                   standard-normal values, VarioConfig(max_dist=1500, n_bins=50): MultiField.get_variogram for
                   (0, 0), (0, 1), (1, 1) -- 5e7 / 1e8 / 5e7 pairs through the reference's pandas path.
  variogram_cloud.npz   MultiField._variogram_cloud on a small case (both kinds).
  preprocess.npz   the preprocessing / back-transform chain (src/fields.py:283-375, src/joint_prediction.py:155-205) on a
                   tiny (time, lat, lon) cube.  xarray is not installed here, so the Dataset plumbing of those functions
                   cannot run; the fixture executes their NUMERIC steps in order with the reference's own callee
                   (stat_tools.simple_linear_regression, unmodified) and the same library calls (sklearn LinearRegression
                   on pandas-standardised covariates, numpy nanmean / nanstd): parity of the xarray glue itself is unpinned.
Only outputs and the small inputs are stored; the large inputs are regenerated from the seeds by the tests
(`inputs_c1` / `inputs_c2` below are imported by them).
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

C1_PARAMS = [1.0, 1.0, 1.5, 1.5, 1.5, 0.2, 0.2, 0.2, 0.01, 0.01, -0.6]


def conus_cells(seed, n):
    lat = np.arange(22.025, 58, 0.05)
    lon = np.arange(-124.975, -65, 0.05)
    idx = np.random.default_rng(seed).choice(len(lat) * len(lon), n, replace=False)
    return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]


def inputs_c1():
    """Grid coordinates in sim.CartesianGrid order and the 500 targets."""
    xs = np.linspace(0, 1, 40)
    grid = np.array(np.meshgrid(xs, xs)).T.reshape(-1, 2)
    return grid, np.random.default_rng(7).uniform(0, 1, (500, 2))


def inputs_c2(n=10000):
    ca, cb = conus_cells(2, n), conus_cells(3, n)
    va, vb = np.random.default_rng(5).standard_normal(n), np.random.default_rng(6).standard_normal(n)
    return [ca, cb], [va, vb]


def inputs_preprocess():
    rng = np.random.default_rng(21)
    T, nla, nlo = 9, 6, 7
    lat, lon = np.linspace(30.0, 35.0, nla), np.linspace(-100.0, -94.0, nlo)
    cube = rng.standard_normal((T, nla, nlo)) + 0.3 * np.arange(T)[:, None, None] + 0.2 * lat[None, :, None] \
        - 0.1 * lon[None, None, :]
    cube[rng.uniform(size=cube.shape) < 0.15] = np.nan
    cube[5] = np.nan  # one month without any data
    return cube, lat, lon, 3


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    from scipy.linalg import cho_factor, cho_solve
    ref = ref_loader.load()
    warnings.simplefilter("ignore")

    def model_from(values, n_procs=2):
        p = ref.model.MaternParams(n_procs=n_procs).set_values(np.asarray(values, float))
        return ref.model.MultivariateMatern(n_procs=n_procs, params=p)

    # ------------------------------------------------------------------ C1
    grid_xy, pc = inputs_c1()
    out = {}
    grid = ref.sim.CartesianGrid(xcount=40, ycount=40)
    assert (grid.coords.values == grid_xy).all()
    for tag, tau in (("t01", 0.01), ("t0", 0.0)):
        pv = np.array(C1_PARAMS)
        pv[8:10] = tau
        mod = model_from(pv)
        rf = ref.sim.BivariateRandomField(mod, grid, seed=1)
        z = [rf.fields[0]["value"].values, rf.fields[1]["value"].values]
        mf = ref_loader.make_multifield(ref, [grid_xy, grid_xy], z)
        PP = ref.point_prediction.Predictor(mod, mf, fast_dist=False, dist_units=None)
        PP.i = 1
        c0 = PP.mod.covariance(1, 0, use_nugget=True)[0]
        dfp = PP._predict_chunk(pd.DataFrame(pc.copy(), columns=["x", "y"]), c0, 0.2)
        kk = np.array([sum(int(c.sum()) for c in PP._local_dist_ix(s0, 0.2)[0]) for s0 in pc])
        JP = ref.joint_prediction.Predictor(mod, mf, fast_dist=False, dist_units=None)
        JP.i = 1
        c_pp, c_dp, sigma = JP._pred_cov(pc), JP._pred_cross_cov(pc), JP._joint_cov()
        zz = np.hstack(z)
        w = cho_solve(cho_factor(sigma.copy(), lower=True), c_dp.copy()).T
        pred, var = w @ zz, np.diagonal(c_pp - w @ c_dp)
        w_lu = np.linalg.solve(sigma, c_dp).T  # a second valid FP64 solver: the parity floor
        pred_lu, var_lu = w_lu @ zz, np.diagonal(c_pp - w_lu @ c_dp)
        out.update({f"z0_{tag}": z[0], f"z1_{tag}": z[1], f"point_pred_{tag}": dfp["pred"].values,
                    f"point_sd_{tag}": dfp["pred_err"].values, f"point_k_{tag}": kk, f"joint_pred_{tag}": pred,
                    f"joint_var_{tag}": var, f"joint_pred_lu_{tag}": pred_lu, f"joint_var_lu_{tag}": var_lu,
                    f"cond_{tag}": np.linalg.cond(sigma)})
        print(tag, "k", kk.min(), kk.max(), "cond", out[f"cond_{tag}"],
              "floor pred", np.abs(pred_lu / pred - 1).max(), "floor var/c0", np.abs(var_lu - var).max() / c0)
    np.savez_compressed(os.path.join(HERE, "at_size_c1.npz"), params=np.array(C1_PARAMS), **out)

    # ------------------------------------------------------------------ variogram cloud (small)
    ca, cb = conus_cells(12, 60), conus_cells(13, 50)
    cb[:8] = ca[:8]
    va, vb = np.random.default_rng(14).standard_normal(60), np.random.default_rng(15).standard_normal(50)
    mfv = ref_loader.make_multifield(ref, [ca, cb], [va, vb])
    outc = {"coords0": ca, "coords1": cb, "v0": va, "v1": vb}
    for kind in ("Semivariogram", "Covariogram"):
        cfg = ref.fields.VarioConfig(1500, 10, kind=kind)
        for (i, j) in ((0, 0), (0, 1), (1, 1)):
            cl = ref.fields.MultiField._variogram_cloud(mfv, i, j, cfg)
            outc[f"{kind.lower()}_dist{i}{j}"] = cl["distance"].values
            outc[f"{kind.lower()}_cloud{i}{j}"] = cl["variogram"].values
    np.savez_compressed(os.path.join(HERE, "variogram_cloud.npz"), **outc)

    # ------------------------------------------------------------------ preprocessing chain (numeric steps of the reference)
    from sklearn.linear_model import LinearRegression
    cube, lat, lon, ti = inputs_preprocess()
    # fit_linear_trend (fields.py:283-287): mean over (lat, lon) per time step, then the reference's own regression helper
    trend = ref.stat_tools.simple_linear_regression(np.nanmean(cube.reshape(len(cube), -1), axis=1))
    field = (cube - trend[:, None, None])[ti]                                   # _preprocess_ds :352-356
    LON, LAT = np.meshgrid(lon, lat, indexing="ij")
    df = pd.DataFrame({"lon": LON.ravel(), "lat": LAT.ravel(), "v": field.T.ravel()}).dropna(subset=["v"]).reset_index(drop=True)
    means = df[["lon", "lat"]].mean(axis=0, skipna=True).values                 # fit_ols :301-305
    scales = df[["lon", "lat"]].std(axis=0, skipna=True).values
    cov = df[["lon", "lat"]].copy()
    for i, c in enumerate(["lon", "lat"]):
        cov[c] = (cov[c] - means[i]) / scales[i]
    model = LinearRegression().fit(cov, df["v"])
    ols = model.predict(cov)
    resid = df["v"].values - ols                                                # :366
    mean, scale = np.nanmean(resid), np.nanstd(resid)                           # :369-370
    std_vals = (resid - mean) / scale
    # _postprocess_predictions (joint_prediction.py:155-205) on made-up standardised predictions at 5 locations
    ploc = pd.DataFrame({"lat": [30.5, 31.5, 33.0, 34.5, 32.2], "lon": [-99.5, -97.0, -95.5, -94.2, -98.8],
                         "pred": [0.3, -1.2, 0.8, 0.05, -0.4], "pred_err": [0.5, 0.7, 0.2, 0.9, 0.6]})
    pc = ploc[["lon", "lat"]].copy()
    for i, c in enumerate(["lon", "lat"]):
        pc[c] = (pc[c] - means[i]) / scales[i]
    back_pred = ploc["pred"].values * scale + mean + model.predict(pc) + trend[ti]
    back_err = ploc["pred_err"].values * scale
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), trend=trend, lon=df["lon"].values, lat=df["lat"].values,
                        standardised=std_vals, spatial_trend=ols, covariate_means=means, covariate_scales=scales,
                        spatial_mean=mean, scale_fact=scale, coef=model.coef_, intercept=model.intercept_,
                        p_lat=ploc["lat"].values, p_lon=ploc["lon"].values, p_pred=ploc["pred"].values,
                        p_err=ploc["pred_err"].values, back_pred=back_pred, back_err=back_err)
    if os.environ.get("CK_GOLDEN_ONLY") == "preprocess":
        return

    # ------------------------------------------------------------------ C2 at size
    coords, values = inputs_c2()
    mf2 = ref_loader.make_multifield(ref, coords, values)
    cfg = ref.fields.VarioConfig(1500, 50)
    out2 = {}
    for (i, j) in ((0, 0), (0, 1), (1, 1)):
        cloud = ref.fields.MultiField._variogram_cloud(mf2, i, j, cfg)
        cloud = cloud[cloud.distance <= cfg.max_dist]
        centers, edges = ref.fields._construct_variogram_bins(cloud, cfg.n_bins)
        g = ref.fields.MultiField.get_variogram(mf2, i, j, cfg).loc[(i, j)]
        full = pd.DataFrame({"bin_center": centers}).merge(g, on="bin_center", how="left")
        out2[f"center{i}{j}"] = centers
        out2[f"mean{i}{j}"] = full["bin_mean"].values
        out2[f"count{i}{j}"] = full["bin_count"].fillna(0).values.astype(np.int64)
        out2[f"pairs{i}{j}"] = np.int64(len(cloud))
        print("C2", i, j, len(cloud), "pairs kept")
        del cloud
    np.savez_compressed(os.path.join(HERE, "at_size_c2.npz"), max_dist=1500.0, n_bins=50, n=10000, **out2)


if __name__ == "__main__":
    main()

"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference modules are imported as they lie on disk through oracle/ref_loader.py (import stubs for
xarray / numba_scipy / geopy / regionmask only).  Everything stored here is an OUTPUT OF
THE REFERENCE'S OWN FUNCTIONS on seeded inputs (inputs are stored alongside), plus the two
known-answer vectors the reference holds in research/simulation_experiment.ipynb cells [11], [16].
The GPU box has no /root/reference: the tests there read only these .npz files.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd
from scipy.linalg import cho_factor, cho_solve

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402

ref = ref_loader.load()
warnings.simplefilter("ignore")


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, {k: np.asarray(v).shape for k, v in arrays.items()})


def model_from(values, n_procs=2):
    p = ref.model.MaternParams(n_procs=n_procs).set_values(np.asarray(values, float))
    return ref.model.MultivariateMatern(n_procs=n_procs, params=p)


def conus_cells(seed, n):
    lat = np.arange(22.025, 58, 0.05)
    lon = np.arange(-124.975, -65, 0.05)
    idx = np.random.default_rng(seed).choice(len(lat) * len(lon), n, replace=False)
    return np.c_[lat[idx % len(lat)], lon[idx // len(lat)]]


def joint_reference(pred, i, pcoords):
    """src/joint_prediction.py:50-78 on the reference's own block builders."""
    pred.i = i
    c_pp = pred._pred_cov(pcoords)
    c_dp = pred._pred_cross_cov(pcoords)
    sigma = pred._joint_cov()
    z = np.hstack([f.values_main for f in pred.mf.fields])
    w = cho_solve(cho_factor(sigma.copy(), lower=True), c_dp.copy()).T
    var = np.diagonal(c_pp - w @ c_dp)
    return c_pp, c_dp, sigma, w @ z, np.nan_to_num(np.sqrt(var))


# ---------------------------------------------------------------- Matern / model
rng = np.random.default_rng(11)
h = np.concatenate([[0.0, 1e-9], 10 ** rng.uniform(-4, 4.3, 400), np.linspace(0, 3000, 60)])
nus = np.array([0.2, 0.39, 0.5, 0.75, 0.82, 1.0, 1.25, 1.5, 2.0, 2.5, 3.2, 3.5])
lens = np.array([100.0, 500.0, 2000.0])
corr = np.array([[ref.model._matern_correlation(nu, ell, h) for ell in lens] for nu in nus])
pv = np.array([1.1, 0.9, 0.75, 1.0, 1.25, 300.0, 400.0, 800.0, 0.01, 0.03, 0.3])
m = model_from(pv)
save("matern", h=h, nus=nus, lens=lens, corr=corr, params=pv,
     cov0=m.covariance(0, h.copy()), cov1=m.covariance(1, h.copy()), cov1_nonug=m.covariance(1, h.copy(), use_nugget=False),
     cross01=m.cross_covariance(0, 1, h.copy()), cross10=m.cross_covariance(1, 0, h.copy()),
     semi0=m.semivariance(0, h.copy()), semi1=m.semivariance(1, h.copy()), xsemi=m.cross_semivariance(0, 1, h.copy()),
     far_h=np.linspace(0.05, 1.6, 300),
     far_corr=np.array([ref.model._matern_correlation(nu, 0.002, np.linspace(0.05, 1.6, 300)) for nu in (0.5, 0.82, 1.5, 3.5)]))

# parameter container API
p2, p1 = ref.model.MaternParams(2), ref.model.MaternParams(1)
save("params_api", names2=np.array(list(p2.get_names()), dtype="U20"), values2=p2.get_values().astype(float),
     bounds2=np.array([list(b) for b in p2.get_bounds()], float),
     names1=np.array(list(p1.get_names()), dtype="U20"), values1=p1.get_values().astype(float),
     sigma2=p2.sigma.values, nu2=p2.nu.values, rho2=p2.rho.values, rho1=p1.rho.values,
     set_vals=pv, nu_after=ref.model.MaternParams(2).set_values(pv).nu.values,
     rho_after=ref.model.MaternParams(2).set_values(pv).rho.values)

# ---------------------------------------------------------------- distances
X1, X2 = conus_cells(21, 120), conus_cells(22, 90)
X2[:10] = X1[:10]
Y1, Y2 = rng.uniform(0, 1, (80, 2)), rng.uniform(0, 1, (70, 2))
save("distances", X1=X1, X2=X2, hav=ref.fields.distance_matrix(X1, X2, fast_dist=True), Y1=Y1, Y2=Y2,
     euc=ref.fields.distance_matrix(Y1, Y2, units=None, fast_dist=False))

# ---------------------------------------------------------------- simulation (src/sim.py)
pv_sim = np.array([1.0, 1.0, 1.5, 1.5, 1.5, 0.2, 0.2, 0.2, 0.01, 0.01, -0.6])
grid = ref.sim.CartesianGrid(xcount=12, ycount=12)
rf = ref.sim.BivariateRandomField(model_from(pv_sim), grid, seed=1)
samples = rf.sample(size=40, epsilon=0.1)
save("sim", params=pv_sim, coords=grid.coords.values, cmat=rf.cmat, chol=rf.chol_fact_lower,
     field0=rf.fields[0].values, field1=rf.fields[1].values, samp0=samples[0].values, samp1=samples[1].values)

# ---------------------------------------------------------------- joint cokriging, Euclidean (C1-mini) and haversine
mf = ref_loader.make_multifield(ref, [grid.coords.values] * 2, [rf.fields[0]["value"].values, rf.fields[1]["value"].values])
pc = np.random.default_rng(7).uniform(0, 1, (40, 2))
pc[:3] = grid.coords.values[[5, 50, 100]]  # targets on data locations (nugget-at-zero quirk)
P = ref.joint_prediction.Predictor(model_from(pv_sim), mf, fast_dist=False, dist_units=None)
c_pp, c_dp, sigma, pred, err = joint_reference(P, 1, pc)
save("joint_euclid", params=pv_sim, coords0=mf.fields[0].coords_main, coords1=mf.fields[1].coords_main,
     z0=mf.fields[0].values_main, z1=mf.fields[1].values_main, pcoords=pc, c_pp=c_pp, c_dp=c_dp, sigma=sigma, pred=pred,
     pred_err=err)

for tag, pvh in (("generic", [1.0, 0.8, 0.75, 1.0, 1.25, 500, 500, 500, .02, .02, -.2]),
                 ("half", [1.0, 0.8, 1.5, 1.5, 1.5, 500, 500, 500, .02, .02, -.2])):
    ca, cb = conus_cells(4, 150), conus_cells(5, 130)
    cb[:20] = ca[:20]
    za, zb = np.random.default_rng(41).standard_normal(150), np.random.default_rng(42).standard_normal(130)
    mfh = ref_loader.make_multifield(ref, [ca, cb], [za, zb])
    pch = conus_cells(6, 60)
    pch[:5] = ca[:5]
    Ph = ref.joint_prediction.Predictor(model_from(pvh), mfh, fast_dist=True)
    c_pp, c_dp, sigma, pred, err = joint_reference(Ph, 0, pch)
    save("joint_haversine_" + tag, params=np.array(pvh, float), coords0=ca, coords1=cb, z0=za, z1=zb, pcoords=pch,
         c_pp=c_pp, c_dp=c_dp, sigma=sigma, pred=pred, pred_err=err)

# univariate kriging
mf1 = ref_loader.make_multifield(ref, [mf.fields[1].coords_main], [mf.fields[1].values_main])
P1 = ref.joint_prediction.Predictor(model_from(pv_sim[[1, 4, 7, 9]], 1), mf1, fast_dist=False, dist_units=None)
c_pp, c_dp, sigma, pred, err = joint_reference(P1, 0, pc)
save("joint_univariate", params=pv_sim[[1, 4, 7, 9]], coords0=mf1.fields[0].coords_main, z0=mf1.fields[0].values_main,
     pcoords=pc, sigma=sigma, c_dp=c_dp, pred=pred, pred_err=err)

# ---------------------------------------------------------------- point cokriging (src/point_prediction.py)
PP = ref.point_prediction.Predictor(model_from(pv_sim), mf, fast_dist=False, dist_units=None)
PP.i = 1
pcp = np.vstack([pc[:25], [[5.0, 5.0]]])  # last target has no neighbour -> (nan, nan)
c0 = PP.mod.covariance(1, 0, use_nugget=True)[0]
dfp = PP._predict_chunk(pd.DataFrame(pcp, columns=["x", "y"]), c0, 0.3)
kk = np.array([sum(int(c.sum()) for c in PP._local_dist_ix(s0, 0.3)[0]) for s0 in pcp])
PP.cv = True
cvp = grid.coords.values[:20]
dfcv = PP._predict_chunk(pd.DataFrame(cvp.copy(), columns=["x", "y"]), c0, 0.25)
save("point_euclid", params=pv_sim, coords0=mf.fields[0].coords_main, coords1=mf.fields[1].coords_main,
     z0=mf.fields[0].values_main, z1=mf.fields[1].values_main, pcoords=pcp, max_dist=0.3, pred=dfp["pred"].values,
     sd=dfp["pred_err"].values, k=kk, blocks00=PP.Sigma["00"], blocks01=PP.Sigma["01"], blocks11=PP.Sigma["11"],
     cv_pcoords=cvp, cv_max_dist=0.25, cv_pred=dfcv["pred"].values, cv_sd=dfcv["pred_err"].values)

# ---------------------------------------------------------------- empirical variograms (src/fields.py)
ca, cb = conus_cells(2, 320), conus_cells(3, 300)
cb[:60] = ca[:60]  # partial co-location: h = 0 pairs in the cross variogram
va, vb = np.random.default_rng(51).standard_normal(320), 0.5 * np.random.default_rng(52).standard_normal(300) + 2.0
mfv = ref_loader.make_multifield(ref, [ca, cb], [va, vb])
for kind in ("Semivariogram", "Covariogram"):
    cfg = ref.fields.VarioConfig(1500, 25, kind=kind)
    df = ref.fields.MultiField.empirical_variograms(mfv, cfg).df
    out = {}
    for (i, j) in ((0, 0), (0, 1), (1, 1)):
        g = df.loc[(i, j)]
        cloud = ref.fields.MultiField._variogram_cloud(mfv, i, j, cfg)
        cloud = cloud[cloud.distance <= cfg.max_dist]
        centers, edges = ref.fields._construct_variogram_bins(cloud, cfg.n_bins)
        full = pd.DataFrame({"bin_center": centers}).merge(g, on="bin_center", how="left")  # pandas>=2 drops empty bins
        out[f"center{i}{j}"] = centers
        out[f"edges{i}{j}"] = edges
        out[f"mean{i}{j}"] = full["bin_mean"].values
        out[f"count{i}{j}"] = full["bin_count"].fillna(0).values.astype(np.int64)
    save("variogram_haversine_" + kind.lower(), coords0=ca, coords1=cb, v0=va, v1=vb, max_dist=1500.0, n_bins=25, **out)

ce = [rng.uniform(0, 1, (200, 2)), rng.uniform(0, 1, (180, 2))]
ve = [rng.standard_normal(200), rng.standard_normal(180)]
mfe = ref_loader.make_multifield(ref, ce, ve)
cfg = ref.fields.VarioConfig(0.6, 15, dist_units=None, fast_dist=False)
df = ref.fields.MultiField.empirical_variograms(mfe, cfg).df
out = {}
for (i, j) in ((0, 0), (0, 1), (1, 1)):
    g = df.loc[(i, j)]
    out[f"center{i}{j}"] = g["bin_center"].values
    out[f"mean{i}{j}"] = g["bin_mean"].values
    out[f"count{i}{j}"] = g["bin_count"].values.astype(np.int64)
save("variogram_euclid", coords0=ce[0], coords1=ce[1], v0=ve[0], v1=ve[1], max_dist=0.6, n_bins=15, **out)

# ---------------------------------------------------------------- composite WLS cost + fit (src/model.py:277-317)
est = ref.fields.MultiField.empirical_variograms(mfv, ref.fields.VarioConfig(1500, 25))
mod = model_from([1.0, 0.5, 1.5, 1.5, 1.5, 500, 500, 500, 0.05, 0.05, 0.1])
cost0 = mod._composite_wls(mod.params.get_values(), est.df)
mod_fit = ref.model.MultivariateMatern().fit(est)
dfv = est.df.reset_index()
save("wls_fit", i=dfv["i"].values, j=dfv["j"].values, bin_center=dfv["bin_center"].values, bin_mean=dfv["bin_mean"].values,
     bin_count=dfv["bin_count"].values, params0=np.array([1.0, 0.5, 1.5, 1.5, 1.5, 500, 500, 500, 0.05, 0.05, 0.1]),
     cost0=cost0, fit_params=mod_fit.params.get_values().astype(float), fit_cost=mod_fit.fit_result.cost)

# ---------------------------------------------------------------- stat_tools
x = rng.standard_normal(40).cumsum()
x[[3, 17]] = np.nan
y = rng.standard_normal(40).cumsum()
y[[5]] = np.nan
Z1, Z2 = rng.standard_normal((4, 3, 30)), rng.standard_normal((4, 3, 30))
Z1[0, 0, :5] = np.nan
dz, slope = ref.stat_tools.detrend(x.copy())
save("stat_tools", x=x, y=y, slr=ref.stat_tools.simple_linear_regression(x.copy()), detrended=dz, slope=np.asarray(slope, float),
     xcor0=ref.stat_tools.compute_xcor_1d(x, y, lag=0), xcor2=ref.stat_tools.compute_xcor_1d(x, y, lag=2),
     xcor_tau=ref.stat_tools.compute_xcor_1d(x, y, lag=2, tau=100), Z1=Z1, Z2=Z2,
     xcor_nd=ref.stat_tools.compute_xcor_nd(Z1, Z2, lag=1, tau=10), count=ref.stat_tools.get_count(Z1))

# ---------------------------------------------------------------- known answers: research/simulation_experiment.ipynb [3]-[5], [11], [16]
param_vals = [1.0, 1.0, 1.5, 1.5, 1.5, 0.2, 0.2, 0.2, 0.0, 0.0, -0.6]
true_mod = model_from(param_vals)
grid51 = ref.sim.CartesianGrid(xcount=51, ycount=51)
rf51 = ref.sim.BivariateRandomField(true_mod, grid51, seed=1)
samples = rf51.sample(size=100, epsilon=np.sqrt(0.01))
srt = [s.sort_values(["x", "y"]) for s in samples]  # order produced by to_fields() (outer merge -> xarray -> dataframe)
mfk = ref_loader.make_multifield(ref, [s[["x", "y"]].values for s in srt], [srt[0]["Z0"].values, srt[1]["Z1"].values])
cok = ref.joint_prediction.Predictor(true_mod, mfk, fast_dist=False, dist_units=None)
_, _, _, pred_co, err_co = joint_reference(cok, 1, grid51.coords.values)
mfu = ref_loader.make_multifield(ref, [srt[1][["x", "y"]].values], [srt[1]["Z1"].values])
uni = model_from(np.array(param_vals)[[1, 4, 7, 9]], 1)
kr = ref.joint_prediction.Predictor(uni, mfu, fast_dist=False, dist_units=None)
_, _, _, pred_kr, err_kr = joint_reference(kr, 0, grid51.coords.values)
print("cokriging head/tail", pred_co[:4], pred_co[-3:], err_co[:4], err_co[-3:])
print("kriging   head/tail", pred_kr[:4], pred_kr[-3:], err_kr[:3], err_kr[-3:])
save("known_answer_simulation_experiment", params=np.array(param_vals), pcoords=grid51.coords.values,
     coords0=mfk.fields[0].coords_main, coords1=mfk.fields[1].coords_main, z0=mfk.fields[0].values_main,
     z1=mfk.fields[1].values_main, cokrig_pred=pred_co, cokrig_err=err_co, krig_pred=pred_kr, krig_err=err_kr,
     # digits printed in the notebook (cells [11] and [16])
     nb_cokrig_pred_head=np.array([1.025, 1.129, 1.177, 1.106]), nb_cokrig_pred_tail=np.array([-0.3236, -0.2804, -0.2439]),
     nb_cokrig_err_head=np.array([0.2072, 0.1824, 0.1494, 0.0871]), nb_cokrig_err_tail=np.array([0.6993, 0.7249, 0.754]),
     nb_krig_pred_head=np.array([1.014, 1.091, 1.125, 1.07]), nb_krig_pred_tail=np.array([-0.6809, -0.6303, -0.5793]),
     nb_krig_err_head=np.array([0.2073, 0.1838, 0.1526]), nb_krig_err_tail=np.array([0.7533, 0.7751, 0.7987]),
     samp0=samples[0].values, samp1=samples[1].values)

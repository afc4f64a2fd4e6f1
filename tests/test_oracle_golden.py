"""Pin the oracle: every function of oracle/cokrig_oracle.py against fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py) and against the notebook known-answer vectors
(research/simulation_experiment.ipynb cells [11], [16]).  CPU only."""
import numpy as np
import pytest

import cokrig_oracle as orc
from conftest import GOLDEN, golden, relerr


def test_matern_correlation_and_scaling():
    g = golden("matern")
    for a, nu in enumerate(g["nus"]):
        for b, ell in enumerate(g["lens"]):
            np.testing.assert_array_equal(orc.matern_correlation(nu, ell, g["h"]), g["corr"][a, b])
    P = orc.Params(g["params"])
    np.testing.assert_array_equal(orc.covariance(P, 0, g["h"]), g["cov0"])
    np.testing.assert_array_equal(orc.covariance(P, 1, g["h"], use_nugget=False), g["cov1_nonug"])
    np.testing.assert_array_equal(orc.cross_covariance(P, 0, 1, g["h"]), g["cross01"])
    np.testing.assert_array_equal(orc.cross_covariance(P, 1, 0, g["h"]), g["cross10"])
    np.testing.assert_array_equal(orc.semivariance(P, 1, g["h"]), g["semi1"])
    np.testing.assert_array_equal(orc.cross_semivariance(P, 0, 1, g["h"]), g["xsemi"])
    for k, nu in enumerate((0.5, 0.82, 1.5, 3.5)):  # far field incl. the kv underflow cut-off
        np.testing.assert_array_equal(orc.matern_correlation(nu, 0.002, g["far_h"]), g["far_corr"][k])


def test_params_layout():
    g = golden("params_api")
    P = orc.Params(g["set_vals"])
    np.testing.assert_array_equal(P.nu, g["nu_after"])
    np.testing.assert_array_equal(P.rho, g["rho_after"])
    with pytest.raises(ValueError):
        orc.Params(np.ones(5))


def test_distances():
    g = golden("distances")
    np.testing.assert_array_equal(orc.distance_matrix(g["X1"], g["X2"], fast_dist=True), g["hav"])
    np.testing.assert_array_equal(orc.distance_matrix(g["Y1"], g["Y2"], units=None), g["euc"])


def test_sim_fields():
    g = golden("sim")
    P = orc.Params(g["params"])
    np.testing.assert_array_equal(orc.expand_grid(xcount=12, ycount=12), g["coords"])
    cmat, low, fields = orc.sim_fields(P, g["coords"], seed=1)
    np.testing.assert_array_equal(cmat, g["cmat"])
    np.testing.assert_allclose(low, g["chol"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(fields[0], g["field0"][:, 2], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fields[1], g["field1"][:, 2], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,metric,i", [("joint_euclid", "euclidean", 1), ("joint_haversine_generic", "haversine", 0),
                                           ("joint_haversine_half", "haversine", 0)])
def test_joint_prediction(name, metric, i):
    g = golden(name)
    P = orc.Params(g["params"])
    cm = [g["coords0"], g["coords1"]]
    np.testing.assert_array_equal(orc.joint_cov(P, cm, metric), g["sigma"])
    np.testing.assert_array_equal(orc.pred_cross_cov(P, i, cm, g["pcoords"], metric), g["c_dp"])
    np.testing.assert_array_equal(orc.pred_cov(P, i, g["pcoords"], metric), g["c_pp"])
    pred, err, _ = orc.joint_predict(P, i, cm, [g["z0"], g["z1"]], g["pcoords"], metric)
    assert relerr(pred, g["pred"]) < 1e-11
    np.testing.assert_allclose(err, g["pred_err"], rtol=1e-9, atol=1e-7)


def test_joint_univariate():
    g = golden("joint_univariate")
    P = orc.Params(g["params"], 1)
    np.testing.assert_array_equal(orc.joint_cov(P, [g["coords0"]], "euclidean"), g["sigma"])
    pred, err, _ = orc.joint_predict(P, 0, [g["coords0"]], [g["z0"]], g["pcoords"], "euclidean")
    assert relerr(pred, g["pred"]) < 1e-11
    np.testing.assert_allclose(err, g["pred_err"], rtol=1e-9, atol=1e-7)


def test_point_prediction():
    g = golden("point_euclid")
    P = orc.Params(g["params"])
    cm, zz = [g["coords0"], g["coords1"]], [g["z0"], g["z1"]]
    blocks = orc.cov_blocks(P, cm, "euclidean")
    for key in ("00", "01", "11"):
        np.testing.assert_array_equal(blocks[key], g["blocks" + key])
    pred, sd, k, _ = orc.point_predict(P, 1, cm, zz, g["pcoords"], float(g["max_dist"]), "euclidean")
    np.testing.assert_array_equal(k, g["k"])
    assert np.isnan(pred[-1]) and np.isnan(g["pred"][-1]) and np.isnan(sd[-1])
    assert relerr(pred[:-1], g["pred"][:-1]) < 1e-11
    np.testing.assert_allclose(sd[:-1], g["sd"][:-1], rtol=1e-9, atol=1e-9)
    pred, sd, _, _ = orc.point_predict(P, 1, cm, zz, g["cv_pcoords"], float(g["cv_max_dist"]), "euclidean", cv=True)
    assert relerr(pred, g["cv_pred"]) < 1e-11
    np.testing.assert_allclose(sd, g["cv_sd"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name,metric,cov", [("variogram_haversine_semivariogram", "haversine", False),
                                             ("variogram_haversine_covariogram", "haversine", True),
                                             ("variogram_euclid", "euclidean", False)])
def test_variograms(name, metric, cov):
    g = golden(name)
    coords, values = [g["coords0"], g["coords1"]], [g["v0"], g["v1"]]
    df = orc.empirical_variograms(coords, values, float(g["max_dist"]), int(g["n_bins"]), metric, cov)
    for (i, j) in ((0, 0), (0, 1), (1, 1)):
        d = df.loc[(i, j)]
        np.testing.assert_array_equal(d["bin_center"].values, g[f"center{i}{j}"])
        np.testing.assert_array_equal(d["bin_count"].values, g[f"count{i}{j}"])
        np.testing.assert_allclose(d["bin_mean"].values, g[f"mean{i}{j}"], rtol=1e-13, atol=0, equal_nan=True)


def test_composite_wls():
    import pandas as pd
    g = golden("wls_fit")
    df = pd.DataFrame({"i": g["i"], "j": g["j"], "bin_center": g["bin_center"], "bin_mean": g["bin_mean"],
                       "bin_count": g["bin_count"]})
    df = df.set_index(["i", "j", df.index])
    assert abs(orc.composite_wls(g["params0"], df) / float(g["cost0"]) - 1) < 1e-12
    assert abs(orc.composite_wls(g["fit_params"], df) / float(g["fit_cost"]) - 1) < 1e-9


def test_known_answers_notebook():
    """research/simulation_experiment.ipynb [11] (cokriging) and [16] (kriging): printed digits."""
    g = golden("known_answer_simulation_experiment")
    P = orc.Params(g["params"])
    pred, err, _ = orc.joint_predict(P, 1, [g["coords0"], g["coords1"]], [g["z0"], g["z1"]], g["pcoords"], "euclidean")
    np.testing.assert_allclose(pred[:4], g["nb_cokrig_pred_head"], rtol=5e-4)
    np.testing.assert_allclose(pred[-3:], g["nb_cokrig_pred_tail"], rtol=5e-4)
    np.testing.assert_allclose(err[:4], g["nb_cokrig_err_head"], rtol=5e-4)
    np.testing.assert_allclose(err[-3:], g["nb_cokrig_err_tail"], rtol=5e-4)
    np.testing.assert_allclose(pred, g["cokrig_pred"], rtol=1e-7, atol=1e-9)
    P1 = orc.Params(g["params"][[1, 4, 7, 9]], 1)
    pred, err, _ = orc.joint_predict(P1, 0, [g["coords1"]], [g["z1"]], g["pcoords"], "euclidean")
    np.testing.assert_allclose(pred[:4], g["nb_krig_pred_head"], rtol=5e-4)
    np.testing.assert_allclose(pred[-3:], g["nb_krig_pred_tail"], rtol=5e-4)
    np.testing.assert_allclose(err[:3], g["nb_krig_err_head"], rtol=5e-4)
    np.testing.assert_allclose(err[-3:], g["nb_krig_err_tail"], rtol=5e-4)


def test_closed_form_loocv_matches_reference_loop():
    """one-factorisation LOOCV equals the reference's re-assembly loop."""
    rng = np.random.default_rng(3)
    cm = [rng.uniform(0, 1, (40, 2)), rng.uniform(0, 1, (35, 2))]
    zz = [rng.standard_normal(40), rng.standard_normal(35)]
    P = orc.Params([1, 1.2, 1.5, 1.5, 1.5, .2, .2, .2, .02, .03, -.5])
    p_loop, s_loop = orc.loocv_reference(P, 1, cm, zz, "euclidean")
    sinv = np.linalg.inv(orc.joint_cov(P, cm, "euclidean"))
    z = np.hstack(zz)
    k = np.arange(40, 75)
    p_cf = z[k] - (sinv @ z)[k] / np.diagonal(sinv)[k]
    s_cf = 1 / np.sqrt(np.diagonal(sinv)[k])
    assert relerr(p_cf, p_loop) < 1e-9 and relerr(s_cf, s_loop) < 1e-10


def test_oracle_at_baseline_c1_size_vs_reference_fixture():
    """The oracle at BASELINE config C1's real size (N = 3 200, k = 142..390 neighbours per target) against outputs of the
    unmodified reference (tests/golden/make_golden_r2.py): point cokriging on a 40-target subset and the joint solve."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden_r2 import C1_PARAMS, inputs_c1
    g = golden("at_size_c1")
    grid, pc = inputs_c1()
    P = orc.Params(C1_PARAMS)
    z = [g["z0_t01"], g["z1_t01"]]
    pr, sd, k, _ = orc.point_predict(P, 1, [grid, grid], z, pc[:40], 0.2, "euclidean")
    np.testing.assert_array_equal(k, g["point_k_t01"][:40])
    assert relerr(pr, g["point_pred_t01"][:40]) < 1e-13 and np.abs(sd - g["point_sd_t01"][:40]).max() < 1e-13
    cm, low, zs = orc.sim_fields(P, grid, seed=1)  # the simulated fields themselves (src/sim.py:33-65)
    np.testing.assert_allclose(zs[0], z[0], rtol=0, atol=1e-11)
    pj, ej, _ = orc.joint_predict(P, 1, [grid, grid], z, pc[:50], "euclidean")
    assert relerr(pj, g["joint_pred_t01"][:50]) < 1e-10
    assert np.abs(ej ** 2 - np.maximum(g["joint_var_t01"][:50], 0)).max() < 1e-12

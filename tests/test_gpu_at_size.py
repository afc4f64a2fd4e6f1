"""GPU parity at the BASELINE configurations' REAL sizes, against fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden_r2.py; the large inputs are regenerated from the seeds):

  C1  40 x 40 simulated grid (N = 3 200), 500 targets: point cokriging with k = 142..390 neighbours per target (mean 329) through
      ck_local_predict AND point_prediction.Predictor.predict_frame (k bit-exact, pred / sd 1e-9), and the joint
      predictor with the trailing updates on the INT8 tensor cores and on FP64 DMMA, at nugget 0.01 (parity run) and
      0 (notebook-faithful, kappa ~ 2e6) with the LU-vs-Cholesky parity floor printed next to the error;
  C2  10 000 cells per variable, 50 bins: 5e7 / 1e8 / 5e7 pairs, counts and centres bit-exact, means 1e-12.
Also: MultiField._variogram_cloud and the module-level joint_prediction._verify_model against the reference.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, golden, relerr

sys.path.insert(0, GOLDEN)
from make_golden_r2 import C1_PARAMS, inputs_c1, inputs_c2  # noqa: E402  (input generators only; no reference needed)

pytestmark = pytest.mark.gpu
TOL = 1e-9


def make_model(values, n_procs=2):
    import model
    return model.MultivariateMatern(n_procs=n_procs, params=model.MaternParams(n_procs=n_procs).set_values(np.asarray(values, float)))


def c1_case(tag):
    g = golden("at_size_c1")
    grid, pc = inputs_c1()
    pv = np.array(C1_PARAMS)
    pv[8:10] = 0.01 if tag == "t01" else 0.0
    return g, grid, pc, pv, [g[f"z0_{tag}"], g[f"z1_{tag}"]]


@pytest.mark.parametrize("tag", ["t01", "t0"])
def test_c1_point_cokriging_at_size(tag):
    """src/point_prediction.py:127-249 at C1: 142..390 neighbours per target (multi-panel, multi-tile path)."""
    import fields, point_prediction
    from cokrig_b200 import METRIC_EUCLID, ops
    g, grid, pc, pv, z = c1_case(tag)
    ref_pred, ref_sd, ref_k = g[f"point_pred_{tag}"], g[f"point_sd_{tag}"], g[f"point_k_{tag}"]
    assert ref_k.max() > 380 and ref_k.mean() > 300 and ref_k.min() > 128
    cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
    c0 = pv[1] ** 2 + pv[9]
    # nugget 0: the local systems have kappa ~ 1e6 and the reference itself is at its rounding floor (~1e-9)
    tol = TOL if tag == "t01" else 2e-8
    # both ways of obtaining the local matrices: re-computed from the coordinates, and gathered from the stored joint
    # covariance like the reference's np.ix_ (src/point_prediction.py:159-179)
    for mode, sigma in (("recompute", None), ("gather", ops.joint_cov(cd, pv, 2, METRIC_EUCLID))):
        pred, sd, k, info = ops.local_predict(cd, [ops.to_device(v) for v in z], ops.coords_to_device(pc), pv, 2, 1,
                                              METRIC_EUCLID, 0.2, sigma=sigma)
        np.testing.assert_array_equal(k, ref_k)  # neighbour sets: bit-exact
        e_pred, e_var = relerr(pred, ref_pred), float(np.abs(sd ** 2 - ref_sd ** 2).max() / c0)
        print(f"\nC1 point {tag} [{mode}]: pred rel {e_pred:.2e}, var/c0 {e_var:.2e}, k {k.min()}..{k.max()}")
        assert (info >= 0).all() or tag == "t0"
        assert e_pred < tol and e_var < tol
        big = ref_sd > 1e-3
        assert relerr(sd[big], ref_sd[big]) < tol
    mf = fields.MultiField.from_arrays([grid, grid], z)
    P = point_prediction.Predictor(make_model(pv), mf, fast_dist=False, dist_units=None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = P.predict_frame(1, pd.DataFrame(pc, columns=["x", "y"]), max_dist=0.2)
    np.testing.assert_array_equal(P.last_neighbour_counts, ref_k)
    assert relerr(df["pred"].values, ref_pred) < tol and np.abs(df["pred_err"].values ** 2 - ref_sd ** 2).max() / c0 < tol


@pytest.mark.parametrize("tag", ["t01", "t0"])
def test_c1_joint_int8_vs_dmma_vs_reference(tag):
    """src/joint_prediction.py:50-78 at C1 (N = 3 200, m = 500): INT8 tensor-core updates vs FP64 DMMA vs the reference,
    with the parity floor (LU vs Cholesky on the reference side) printed."""
    import fields, joint_prediction
    from cokrig_b200 import METRIC_EUCLID, _lib, ops
    g, grid, pc, pv, z = c1_case(tag)
    ref_pred, ref_var = g[f"joint_pred_{tag}"], g[f"joint_var_{tag}"]
    c0 = pv[1] ** 2 + pv[9]
    floor_pred = relerr(g[f"joint_pred_lu_{tag}"], ref_pred)
    floor_var = float(np.abs(g[f"joint_var_lu_{tag}"] - ref_var).max() / c0)
    cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
    zd = ops.to_device(np.hstack(z))
    res = {}
    try:
        for name, (enabled, min_rows) in (("int8", (1, 1024)), ("dmma", (0, -1))):
            _lib.lib.ck_oz_configure(enabled, min_rows)
            assert _lib.lib.ck_oz_active(len(grid) * 2) == enabled
            f = ops.potrf(ops.joint_cov(cd, pv, 2, METRIC_EUCLID))
            pred, var = f.predict(ops.cross_cov(cd, ops.coords_to_device(pc), pv, 2, 1, METRIC_EUCLID), zd, c0)
            assert f.info == 0
            res[name] = (pred.cpu().numpy(), var.cpu().numpy())
    finally:
        _lib.lib.ck_oz_configure(1, 1024)
    print(f"\nC1 joint {tag}: kappa {float(g[f'cond_{tag}']):.2e}; parity floor (reference LU vs Cholesky): pred {floor_pred:.2e}, "
          f"var/c0 {floor_var:.2e}")
    for name, (pred, var) in res.items():
        e_pred, e_var = relerr(pred, ref_pred), float(np.abs(var - ref_var).max() / c0)
        print(f"  {name}: pred rel {e_pred:.2e}, var/c0 {e_var:.2e}")
        assert e_pred < max(TOL, 4 * floor_pred) and e_var < max(TOL, 4 * floor_var), name
    d_pred = relerr(res["int8"][0], res["dmma"][0])
    print(f"  int8 vs dmma: pred rel {d_pred:.2e}, var/c0 {np.abs(res['int8'][1] - res['dmma'][1]).max() / c0:.2e}")
    assert d_pred < max(TOL, 4 * floor_pred)
    # the same through the drop-in predictor (default switches)
    mf = fields.MultiField.from_arrays([grid, grid], z)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = joint_prediction.Predictor(make_model(pv), mf, fast_dist=False, dist_units=None).predict_frame(
            1, pd.DataFrame(pc, columns=["x", "y"]))
    assert relerr(df["pred"].values, ref_pred) < max(TOL, 4 * floor_pred)
    with np.errstate(invalid="ignore"):
        ref_err = np.nan_to_num(np.sqrt(ref_var))
    assert np.abs(df["pred_err"].values ** 2 - ref_err ** 2).max() / c0 < max(TOL, 4 * floor_var)


def test_c2_variograms_at_size():
    """src/fields.py:208-232 at C2: 10 000 cells per variable, 50 bins -- counts and centres bit-exact, means 1e-12."""
    import fields
    from cokrig_b200 import METRIC_HAVERSINE
    g = golden("at_size_c2")
    coords, values = inputs_c2(int(g["n"]))
    for (i, j) in ((0, 0), (0, 1), (1, 1)):
        centers, _, counts, sums = fields._device_variogram(coords[i], values[i], coords[j], values[j], i == j,
                                                            METRIC_HAVERSINE, False, float(g["max_dist"]), int(g["n_bins"]))
        np.testing.assert_array_equal(centers, g[f"center{i}{j}"])
        np.testing.assert_array_equal(counts, g[f"count{i}{j}"])
        assert int(counts.sum()) == int(g[f"pairs{i}{j}"])
        assert relerr(sums / counts, g[f"mean{i}{j}"]) < 1e-12
    mf = fields.MultiField.from_arrays(coords, values)
    est = mf.empirical_variograms(fields.VarioConfig(float(g["max_dist"]), int(g["n_bins"])))
    assert len(est.df) == 150
    np.testing.assert_array_equal(est.df.loc[(0, 1)]["bin_count"].values, g["count01"])


@pytest.mark.parametrize("kind", ["Semivariogram", "Covariogram"])
def test_variogram_cloud_dropin(kind):
    """src/fields.py:192-206: pair order (row-major strict upper triangle / all pairs), values, shapes."""
    import fields
    g = golden("variogram_cloud")
    mf = fields.MultiField.from_arrays([g["coords0"], g["coords1"]], [g["v0"], g["v1"]])
    cfg = fields.VarioConfig(1500, 10, kind=kind)
    for (i, j) in ((0, 0), (0, 1), (1, 1)):
        df = mf._variogram_cloud(i, j, cfg)
        assert list(df.columns) == ["distance", "variogram"]
        d_ref, c_ref = g[f"{kind.lower()}_dist{i}{j}"], g[f"{kind.lower()}_cloud{i}{j}"]
        assert len(df) == len(d_ref)
        assert ((df["distance"].values == 0) == (d_ref == 0)).all()
        assert relerr(df["distance"].values[d_ref > 0], d_ref[d_ref > 0]) < 1e-14
        np.testing.assert_allclose(df["variogram"].values, c_ref, rtol=1e-15, atol=0)


def test_joint_verify_model_module_function():
    """src/joint_prediction.py:260-274: raises LinAlgError iff the augmented matrix is not positive definite; and the
    predictor's in-call test is the exact one (Schur complement), e.g. duplicate targets."""
    import fields, joint_prediction
    from scipy.linalg import LinAlgError
    g = golden("joint_euclid")
    c_pp, c_dp, sigma = g["c_pp"][3:, 3:], g["c_dp"][:, 3:], g["sigma"]  # drop the three targets that sit on data
    joint_prediction._verify_model(c_pp, c_dp, sigma)  # PD: no exception
    with pytest.raises(LinAlgError):
        joint_prediction._verify_model(g["c_pp"], g["c_dp"], sigma - 0.02 * np.eye(len(sigma)))
    bad = c_pp.copy()
    bad[0, 0] = -1.0
    with pytest.raises(LinAlgError):
        joint_prediction._verify_model(bad, c_dp, sigma)
    # the predictor's in-call test is the exact one: Cholesky of the m x m Schur complement C_pp - V V^T.  A matrix with a
    # positive diagonal that is not PD must be caught (the diagonal-only test of round 1 would miss it).
    import torch
    from cokrig_b200 import METRIC_EUCLID, ops
    cd = [ops.coords_to_device(g["coords0"]), ops.coords_to_device(g["coords1"])]
    f = ops.potrf(ops.joint_cov(cd, g["params"], 2, METRIC_EUCLID))
    V = torch.zeros((3, ops.padded_ld(f.n)), dtype=torch.float64, device="cuda")[:, :f.n]
    cpp_bad = torch.tensor([[1.0, 0.9, 0.9], [0.9, 1.0, -0.9], [0.9, -0.9, 1.0]], dtype=torch.float64, device="cuda")
    assert f.schur_info(V, cpp_bad) == 3
    cpp_ok = torch.tensor([[1.0, 0.5, 0.2], [0.5, 1.0, 0.5], [0.2, 0.5, 1.0]], dtype=torch.float64, device="cuda")
    assert f.schur_info(V, cpp_ok) == 0
    mf = fields.MultiField.from_arrays([g["coords0"], g["coords1"]], [g["z0"], g["z1"]])
    P = joint_prediction.Predictor(make_model(g["params"]), mf, fast_dist=False, dist_units=None)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        P.predict_frame(1, g["pcoords"][3:])
    assert not any("not positive definte" in str(x.message) for x in w)
    P0 = joint_prediction.Predictor(make_model(np.r_[g["params"][:8], 0.0, 0.0, g["params"][10]]), mf, fast_dist=False,
                                    dist_units=None)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        P0.predict_frame(1, g["pcoords"])  # zero nugget, three targets on data locations: singular augmented matrix
    assert any("not positive definte" in str(x.message) for x in w)

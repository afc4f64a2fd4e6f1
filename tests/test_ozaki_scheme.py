"""CPU checks of the INT8 update scheme (numpy model of csrc/ck_ozaki.cu in tests/ozaki_model.py): the digit recoding
is exact, the sliced product is at least as accurate as a plain FP64 product, and a blocked Cholesky + cokriging
solve whose every update goes through the scheme reproduces the oracle's predictions like LAPACK does."""
import numpy as np

import cokrig_oracle as orc
import ozaki_model as oz


def test_digits_recode_the_fixed_point_value_exactly():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 64)) * np.exp(rng.uniform(-30, 30, (50, 1)))
    x[3] = 0.0
    digits, scales = oz.split(x)
    assert all(((d >= -128) & (d <= 127)).all() for d in digits)
    rec = sum(d * 2.0 ** (-8 * p) for p, d in enumerate(digits)) * scales[:, None]
    rowmax = np.maximum(np.abs(x).max(axis=1, keepdims=True), 1e-300)
    assert (np.abs(rec - x) / rowmax).max() <= 2.0 ** -56 / 0.49
    assert (rec[3] == 0).all() and scales[3] == 0


def test_sliced_product_is_fp64_accurate():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((96, 1024)), rng.standard_normal((80, 1024))
    ld = np.longdouble
    exact = a.astype(ld) @ b.T.astype(ld)
    rowmax = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :]
    err_oz = np.abs(oz.product(a, b).astype(ld) - exact).astype(float) / rowmax
    err_fp = np.abs((a @ b.T).astype(ld) - exact).astype(float) / rowmax
    assert err_oz.max() < 32 * np.sqrt(1024) * 2.0 ** -56
    assert err_oz.max() < err_fp.max()  # fixed point with exact accumulation beats sequential FP64 rounding


def test_cokriging_through_the_scheme_matches_the_oracle():
    params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .01, .01, -.6]
    P = orc.Params(params)
    grid = orc.expand_grid(xcount=24, ycount=24)
    _, _, z = orc.sim_fields(P, grid, seed=1)
    pc = np.random.default_rng(7).uniform(0, 1, (200, 2))
    sigma = orc.joint_cov(P, [grid, grid], "euclidean")
    cdp = orc.pred_cross_cov(P, 1, [grid, grid], pc, "euclidean")
    ref_pred, ref_err, _ = orc.joint_predict(P, 1, [grid, grid], z, pc, "euclidean")
    L = oz.cholesky_blocked(sigma, 256, oz.product)
    V = oz.solve_blocked(L, np.c_[cdp, np.hstack(z)].T, 256, oz.product)
    pred, var = V[:-1] @ V[-1], 1.01 - (V[:-1] ** 2).sum(axis=1)
    assert np.abs(pred - ref_pred).max() / np.abs(ref_pred).max() < 1e-11
    assert np.abs(var - ref_err ** 2).max() < 1e-11
    from scipy.linalg import cholesky
    assert np.abs(L - cholesky(sigma, lower=True)).max() < 1e-12

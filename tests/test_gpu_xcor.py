"""GPU parity of the batched temporal cross-correlation kernel (ck_xcor_lags) against the reference's
compute_xcor_nd fixture (tests/golden/stat_tools.npz, produced by the unmodified reference) and, for whole lag ranges,
detrending, tau thresholds and the arg-max of optim_lag_nd, against the numpy port of src/stat_tools.py that the CPU tier
pins to the same fixture."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _same(a, b, tol=1e-12):
    assert a.shape == b.shape
    assert (np.isnan(a) == np.isnan(b)).all()
    ok = ~np.isnan(b)
    assert np.abs(a[ok] - b[ok]).max() < tol if ok.any() else True


def test_xcor_kernel_vs_reference_fixture():
    import stat_tools
    g = golden("stat_tools")
    got = stat_tools.xcor_lags_device(g["Z1"], g["Z2"], [1], tau=10)[0]
    _same(got, g["xcor_nd"])


def test_xcor_kernel_all_lags_tau_and_detrend_vs_numpy_port():
    import stat_tools
    rng = np.random.default_rng(8)
    Z1 = rng.standard_normal((9, 7, 71)).cumsum(axis=-1)
    Z2 = 0.6 * np.roll(Z1, 2, axis=-1) + rng.standard_normal((9, 7, 71))
    Z1[rng.uniform(size=Z1.shape) < 0.2] = np.nan
    Z2[rng.uniform(size=Z2.shape) < 0.3] = np.nan
    Z1[0, 0] = np.nan            # a cell without data
    Z2[1, 1, 5:] = np.nan        # almost empty series
    Z1[2, 2] = 3.0               # constant series: zero denominator -> NaN
    lags = list(range(-3, 5))
    for tau in (None, 25):
        got = stat_tools.xcor_lags_device(Z1, Z2, lags, tau=tau)
        for k, lag in enumerate(lags):
            _same(got[k], stat_tools.compute_xcor_nd(Z1, Z2, lag=lag, tau=tau))
    # detrend + every lag + arg-max == optim_lag_nd's arithmetic (src/stat_tools.py:181-233) on plain arrays
    D1 = np.apply_along_axis(lambda v: stat_tools.detrend(v)[0], -1, Z1)
    D2 = np.apply_along_axis(lambda v: stat_tools.detrend(v)[0], -1, Z2)
    bnds = (-2, 4)
    stack = np.ma.masked_invalid(np.stack([stat_tools.compute_xcor_nd(D1, D2, lag=lag, tau=20) for lag in np.arange(*bnds)], axis=2))
    ref_idx = np.ma.argmax(np.abs(stack), axis=2)
    ref_val = np.ma.filled(np.squeeze(np.take_along_axis(stack, np.expand_dims(ref_idx, axis=2), 2), axis=2).astype(float), np.nan)
    idx, val = stat_tools.optim_lag_arrays(Z1, Z2, bnds, tau=20)
    _same(val, ref_val)
    decided = ~np.isnan(ref_val)
    np.testing.assert_array_equal(idx[decided], ref_idx[decided])
    assert (idx[~decided] == 0).all()  # all-NaN cells: lag index 0, like np.ma.argmax


def test_xcor_kernel_shapes_and_bad_arguments():
    from cokrig_b200 import ops
    x = np.random.default_rng(1).standard_normal((5, 30))
    xc, bi, bx = ops.xcor_lags(x, x, [0])
    assert xc.shape == (1, 5) and bi.shape == (5,) and np.allclose(xc, 1.0) and np.allclose(bx, 1.0)
    with pytest.raises(ValueError):
        ops.xcor_lags(x, x[:, :-1], [0])
    with pytest.raises(Exception):
        ops.xcor_lags(x, x, list(range(100)))  # more than 64 lags per call

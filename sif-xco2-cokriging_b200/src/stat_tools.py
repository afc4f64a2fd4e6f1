"""Drop-in replacement for the reference's ``src/stat_tools.py`` (temporal statistics).

Same names and signatures as /root/reference/src/stat_tools.py:9-271.  These are per-cell 1-D
reductions over <= ~300 time steps (counts, linear detrending, lagged cross-correlation): O(n) host
work outside the cokriging hot path, kept as vectorised numpy.  The
``apply_*`` wrappers and ``optim_lag_nd`` / ``get_stats`` need xarray (imported lazily).

Device counterparts on plain arrays (no xarray): ``xcor_lags_device`` evaluates ``compute_xcor_nd`` for a whole range of
lags -- and ``optim_lag_arrays`` the detrend + per-lag cross-correlation + arg-max of ``optim_lag_nd`` -- in ONE streaming
kernel launch over the (lon, lat, time) cubes (``ck_xcor_lags``) instead of one masked-array numpy pass per lag.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _xr():
    import xarray
    return xarray


# -- counts / replications -----------------------------------------------------------------------
def get_count(da):
    """Number of non-missing elements along the time (last) dimension: (lon x lat x time) -> (lon x lat)."""
    return np.count_nonzero(~np.isnan(da), axis=-1)


def apply_count(da):
    return _xr().apply_ufunc(get_count, da, input_core_dims=[["time"]], output_dtypes=[float], dask="parallelized")


# -- trend fitting ---------------------------------------------------------------------------------
def _index_trend(x: np.ndarray):
    """Least-squares line through (index, value) over the non-missing entries: (mask, slope, fitted)."""
    ok = ~np.isnan(x)
    t = np.arange(x.size, dtype=float)[ok]
    y = x[ok]
    tc = t - t.mean()
    denom = float(np.dot(tc, tc))
    slope = float(np.dot(tc, y - y.mean()) / denom) if denom > 0 else 0.0
    return ok, slope, y.mean() + slope * tc


def simple_linear_regression(x):
    """Fit a linear trend to a vector using its indices as the covariate; returns the trend vector
    (missing entries stay missing)."""
    if np.isnan(x).all():
        return x
    ok, _, fitted = _index_trend(x)
    pred = np.copy(x)
    pred[ok] = fitted
    return pred


def detrend(x):
    """Remove the linear index trend from a vector; returns (detrended vector, slope).
    An all-missing input is returned unchanged together with NaN."""
    if np.isnan(x).all():
        return x, np.nan
    ok, slope, fitted = _index_trend(x)
    z = np.copy(x)
    z[ok] = x[ok] - fitted
    return z, np.array([slope])


def apply_detrend(da):
    return _xr().apply_ufunc(detrend, da, input_core_dims=[["time"]], output_core_dims=[["time"], []],
                             output_dtypes=[float, float], dask="parallelized", vectorize=True)


# -- cross-correlation -----------------------------------------------------------------------------
def compute_xcor_1d(v1, v2, lag=0, tau=None):
    """Empirical cross-correlation of two 1-d series at an integer lag; NaN if fewer than `tau`
    jointly non-missing products are available."""
    x = np.ma.masked_invalid(np.asarray(v1, dtype=float))
    y = np.ma.masked_invalid(np.asarray(v2, dtype=float))
    x = x - x.mean()
    y = y - y.mean()
    if lag != 0:
        x, y = x[lag:], y[:-lag]
    prod = x * y
    if tau is not None and np.count_nonzero(~np.isnan(prod)) < tau:
        return np.nan
    xcor = np.sum(prod) / (np.sqrt(np.sum(x * x)) * np.sqrt(np.sum(y * y)))
    return np.ma.filled(xcor.astype(float), np.nan)


def compute_xcor_nd(Z1, Z2, lag=0, tau=None):
    """Empirical cross-correlation along the last (time) axis of two (lon x lat x time) arrays."""
    X = np.ma.masked_invalid(np.asarray(Z1, dtype=float))
    Y = np.ma.masked_invalid(np.asarray(Z2, dtype=float))
    X = X - X.mean(axis=-1, keepdims=True)
    Y = Y - Y.mean(axis=-1, keepdims=True)
    if lag != 0:
        X, Y = X[:, :, lag:], Y[:, :, :-lag]
    prod = X * Y
    xcor = np.sum(prod, axis=-1) / (np.sqrt(np.sum(X * X, axis=-1)) * np.sqrt(np.sum(Y * Y, axis=-1)))
    if tau:
        xcor = np.ma.masked_where(np.count_nonzero(~np.isnan(prod), axis=-1) < tau, xcor)
    return np.ma.filled(xcor.astype(float), np.nan)


def apply_xcor(da1, da2, lag=0, tau=None):
    Z1, _ = apply_detrend(da1)
    Z2, _ = apply_detrend(da2)
    return _xr().apply_ufunc(compute_xcor_nd, Z1, Z2, kwargs={"lag": lag, "tau": tau},
                             input_core_dims=[["time"], ["time"]], output_dtypes=[float], dask="parallelized")


def optim_lag_nd(da1, da2, lag_bnds, tau=None):
    """Lag (within integer bounds) maximising |cross-correlation| along time, per (lon, lat) cell."""
    xarray = _xr()
    Z1, _ = apply_detrend(da1)
    Z2, _ = apply_detrend(da2)
    fields = [xarray.apply_ufunc(compute_xcor_nd, Z1, Z2, kwargs={"lag": lag, "tau": tau},
                                 input_core_dims=[["time"], ["time"]], output_dtypes=[float],
                                 dask="parallelized").values for lag in np.arange(*lag_bnds)]
    stack = np.ma.masked_invalid(np.stack(fields, axis=2))
    optim_lag = np.ma.argmax(np.abs(stack), axis=2)  # all-NaN slices give lag index 0
    xcor = np.squeeze(np.take_along_axis(stack, np.expand_dims(optim_lag, axis=2), 2), axis=2)
    return xarray.Dataset({"optim_lag": (["lon", "lat"], optim_lag), "xcor": (["lon", "lat"], xcor)},
                          coords={"lon": da1.lon, "lat": da1.lat})


def xcor_lags_device(Z1, Z2, lags, tau=None, detrend: bool = False):
    """``compute_xcor_nd(Z1, Z2, lag, tau)`` for every lag of `lags` at once, on the device: array of shape
    (len(lags), lon, lat).  detrend=True first removes each cell's linear index trend (``apply_detrend``)."""
    from _backend import ops
    return ops.xcor_lags(Z1, Z2, lags, tau=tau, detrend=detrend)[0]


def optim_lag_arrays(Z1, Z2, lag_bnds, tau=None):
    """``optim_lag_nd`` (:181-233) on plain (lon, lat, time) arrays: per cell the INDEX into ``np.arange(*lag_bnds)`` of
    the lag with the largest |cross-correlation| of the detrended series (0 where every lag is NaN) and that
    cross-correlation.  Returns (optim_lag, xcor) -- the two variables of the reference's Dataset."""
    from _backend import ops
    _, idx, val = ops.xcor_lags(Z1, Z2, np.arange(*lag_bnds), tau=tau, detrend=True)
    return idx.astype(np.int64), val


# -- wrappers --------------------------------------------------------------------------------------
def get_stats(DS):
    """Counts, slopes and residual standard deviations for the SIF and XCO2 arrays of a dataset."""
    DS["sif_count"] = apply_count(DS.sif)
    DS["xco2_count"] = apply_count(DS.xco2)
    sif_resid, DS["sif_slope"] = apply_detrend(DS.sif)
    xco2_resid, DS["xco2_slope"] = apply_detrend(DS.xco2)
    DS["sif_std"] = sif_resid.std(dim="time")
    DS["xco2_std"] = xco2_resid.std(dim="time")
    return DS


def get_stats_df(df_group, lags=[0], tau=None):
    """Count, slope, std. dev. and lagged cross-correlations for `sif` / `xco2` data frame columns."""
    sif_resid, sif_slope = detrend(df_group["sif"].values)
    xco2_resid, xco2_slope = detrend(df_group["xco2"].values)
    df = pd.DataFrame({
        "sif_count": df_group["sif"].dropna().count(), "xco2_count": df_group["xco2"].dropna().count(),
        "sif_slope": sif_slope, "xco2_slope": xco2_slope,
        "sif_std": np.nanstd(sif_resid), "xco2_std": np.nanstd(xco2_resid)})
    for lag in lags:
        df[f"xcor_lag{lag}"] = compute_xcor_1d(xco2_resid, sif_resid, lag=lag, tau=tau)
    return df

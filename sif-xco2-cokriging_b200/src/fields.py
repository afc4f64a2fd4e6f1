"""Drop-in replacement for the reference's ``src/fields.py``: fields, distances, empirical variograms.

Public names and signatures follow /root/reference/src/fields.py:20-403.  The hot functions --
``distance_matrix``, ``MultiField.calc_dist_matrix``, ``MultiField.get_variogram`` /
``empirical_variograms`` -- run on the device (``ck_distance_block``, ``ck_vario_minmax``,
``ck_vario_bin``); the variogram never materialises the pair cloud (the reference builds 16 B per
pair plus pandas categoricals, fields.py:192-222).  xarray-based preprocessing (``Field`` for real
data, ``_preprocess_ds``, ``fit_ols``) is host-only O(n) work outside the hot path and is kept
behind a lazy ``import xarray``; ``Field.from_arrays`` / ``MultiField.from_arrays`` build the same
objects straight from numpy arrays, and ``preprocess_arrays`` / ``Field.from_cube`` /
``MultiField.from_cubes`` run the reference's whole preprocessing chain (temporal linear trend of the
spatial mean, OLS spatial trend on standardised covariates, standardisation; fields.py:283-375) on a
plain ``(time, lat, lon)`` numpy cube -- no xarray at run time.

Preserved semantics: rows are [lat, lon] degrees (haversine x 6371 km) or
[x, y] (Euclidean, ``units=None``); residuals use the mean of ALL values; marginal variograms use
pairs a < b, cross variograms all pairs including co-located ones; bins are built from the minimum
non-zero and the maximum retained distance; ``pd.cut(include_lowest=True)`` decisions; all
``n_bins`` rows are emitted (empty bins: NaN mean, 0 count -- reference-era pandas behaviour).
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass

import numpy as np
import pandas as pd

from _backend import ops

EARTH_RADIUS = 6371  # radius in kilometers


class VarioConfig:
    """Configuration of an empirical variogram.

    dist_units: distance units (km by default); fast_dist: great-circle (haversine) distances.
    """

    def __init__(self, max_dist: float, n_bins: int, n_procs: int = 2, kind: str = "Semivariogram",
                 dist_units: str = "km", fast_dist: bool = True) -> None:
        self.max_dist = max_dist
        self.n_bins = n_bins
        self.n_procs = n_procs
        self.kind = kind
        self.dist_units = dist_units
        self.fast_dist = fast_dist
        self.covariogram = self.kind == "Covariogram"


@dataclass
class EmpiricalVariogram:
    """Empirical variogram"""

    df: pd.DataFrame
    config: VarioConfig
    timestamp: str
    timedeltas: list


class Field:
    """Data values and coordinates of a single process at a fixed time."""

    def __init__(self, ds, covariates: list, timestamp, type: str = "real") -> None:
        from data_utils import get_main_coords
        self.timestamp = timestamp
        self.data_name, self.var_name = _get_field_names(ds)
        if type == "real":
            self.ds = _preprocess_ds(ds, timestamp, covariates)
            self.ds_main = get_main_coords(self.ds).sel(time=timestamp)
            df = self.to_dataframe()
            df_main = self.to_dataframe(main=True)
            self.coords = df[["lat", "lon"]].values
            self.coords_main = df_main[["lat", "lon"]].values
            self.values = df[self.data_name].values
            self.values_main = df_main[self.data_name].values
            self.temporal_trend = self.ds.attrs["temporal_trend"]
            self.spatial_trend = df["spatial_trend"].values
            self.spatial_mean = self.ds.attrs["spatial_mean"]
            self.scale_fact = self.ds.attrs["scale_fact"]
            self.covariate_means = self.ds.attrs["covariate_means"]
            self.covariate_scales = self.ds.attrs["covariate_scales"]
            self.variance_estimate = df[self.var_name].values
            self.covariates = df[covariates]
        else:
            self.ds_main = ds.assign_coords(coords={"time": np.nan})
            df_main = self.to_dataframe(main=True)
            self.coords = self.coords_main = df_main[["x", "y"]].values
            self.values = self.values_main = df_main[self.data_name].values
        self.size = len(self.values)

    @classmethod
    def from_arrays(cls, coords, values, coords_main=None, values_main=None, timestamp=np.nan,
                    data_name: str = "value") -> "Field":
        """Build a field directly from arrays (no xarray): coords (n, 2), values (n,)."""
        self = object.__new__(cls)
        self.timestamp = timestamp
        self.data_name, self.var_name = data_name, data_name + "_var"
        self.coords = np.ascontiguousarray(np.asarray(coords, dtype=float))
        self.values = np.ascontiguousarray(np.asarray(values, dtype=float))
        self.coords_main = self.coords if coords_main is None else np.ascontiguousarray(np.asarray(coords_main, dtype=float))
        self.values_main = self.values if values_main is None else np.ascontiguousarray(np.asarray(values_main, dtype=float))
        self.size = len(self.values)
        self.ds = self.ds_main = None
        return self

    @classmethod
    def from_cube(cls, cube, lat, lon, t_index: int, covariates=None, timestamp=np.nan, main_mask=None,
                  variance=None, data_name: str = "value") -> "Field":
        """The reference's ``Field(ds, covariates, timestamp, type="real")`` (fields.py:64-95) from plain arrays:
        ``cube`` is (time, lat, lon) with NaN for missing cells, ``t_index`` selects the time slice, ``covariates`` as in
        ``preprocess_arrays``; ``main_mask`` (lat, lon) marks the base-grid cells (``get_main_coords``); rows are ordered
        (lon, lat) like the data frame of a (lon, lat, time) dataset."""
        pre = preprocess_arrays(cube, lat, lon, t_index, covariates)
        self = object.__new__(cls)
        self.timestamp = timestamp
        self.data_name, self.var_name = data_name, data_name + "_var"
        self.ds = self.ds_main = None
        self.attrs = pre["attrs"]
        df = pre["frame"]
        keep = df["value"].notna().values
        self.coords = df.loc[keep, ["lat", "lon"]].values
        self.values = df.loc[keep, "value"].values
        if main_mask is None:
            self.coords_main, self.values_main = self.coords, self.values
        else:
            mm = np.asarray(main_mask, dtype=bool).T.ravel()[keep]  # frame rows are (lon, lat) ordered
            self.coords_main, self.values_main = self.coords[mm], self.values[mm]
        self.temporal_trend = self.attrs["temporal_trend"]
        self.spatial_trend = df.loc[keep, "spatial_trend"].values
        self.spatial_mean = self.attrs["spatial_mean"]
        self.scale_fact = self.attrs["scale_fact"]
        self.covariate_means = self.attrs["covariate_means"]
        self.covariate_scales = self.attrs["covariate_scales"]
        self.variance_estimate = None if variance is None else np.asarray(variance, dtype=float)[t_index].T.ravel()[keep]
        self.covariates = df.loc[keep, pre["covariate_names"]]
        self.size = len(self.values)
        return self

    def to_dataframe(self, main: bool = False):
        """Converts the field to a data frame."""
        ds = self.ds_main if main else self.ds
        if ds is None:  # array-built field
            c, v = (self.coords_main, self.values_main) if main else (self.coords, self.values)
            return pd.DataFrame({"lat": c[:, 0], "lon": c[:, 1], self.data_name: v})
        return ds.to_dataframe().reset_index().dropna(subset=[self.data_name])

    def to_xarray(self):
        """Converts the field to an xarray dataset."""
        return (pd.DataFrame({"lat": self.coords[:, 0], "lon": self.coords[:, 1], self.data_name: self.values})
                .set_index(["lon", "lat"]).to_xarray()
                .assign_coords({"time": np.array(self.timestamp, dtype=np.datetime64)}))


class MultiField:
    """A multivariate process (one ``Field`` per process) with modelling attributes.

    datasets: list of xarray datasets; covariates: per-dataset covariate names for the spatial trend;
    timestamp: main timestamp; timedeltas: month offsets per dataset.
    """

    def __init__(self, datasets: list, covariates: list, timestamp, timedeltas: list, type: str = "real") -> None:
        self.type = type
        self.datasets = datasets
        if type == "real":
            _check_length_match(datasets, covariates, timedeltas)
            self.timestamp = np.datetime_as_string(timestamp, unit="D")
            self.timedeltas = timedeltas
            self.covariates = covariates
            self.fields = np.array([Field(ds, cov, self._apply_timedelta(td), type=type)
                                    for ds, cov, td in zip(datasets, covariates, timedeltas)])
        else:
            self.timestamp = np.nan
            self.timedeltas = [np.nan, np.nan]
            self.fields = np.array([Field(ds, None, np.nan, type=type) for ds in datasets])
        self.n_procs = len(self.fields)
        self.n_data = self._count_data()

    @classmethod
    def from_arrays(cls, coords: list, values: list, coords_main: list = None, values_main: list = None,
                    timestamp=np.nan, timedeltas=None, type: str = "sim") -> "MultiField":
        """Array-based constructor: one (n_i, 2) coordinate array and one value vector per process."""
        self = object.__new__(cls)
        self.type = type
        self.datasets = None
        self.timestamp = timestamp
        self.timedeltas = [np.nan] * len(coords) if timedeltas is None else timedeltas
        cm = [None] * len(coords) if coords_main is None else coords_main
        vm = [None] * len(coords) if values_main is None else values_main
        fields = [Field.from_arrays(c, v, a, b, timestamp=timestamp) for c, v, a, b in zip(coords, values, cm, vm)]
        self.fields = np.empty(len(fields), dtype=object)
        for k, f in enumerate(fields):
            self.fields[k] = f
        self.n_procs = len(self.fields)
        self.n_data = self._count_data()
        return self

    @classmethod
    def from_cubes(cls, cubes: list, lats: list, lons: list, t_indices: list, covariates: list = None, timestamp=np.nan,
                   timedeltas=None, main_masks: list = None) -> "MultiField":
        """``MultiField(datasets, covariates, timestamp, timedeltas)`` for real data (fields.py:135-171) from plain
        (time, lat, lon) cubes: one ``Field.from_cube`` per process."""
        n = len(cubes)
        self = object.__new__(cls)
        self.type = "real"
        self.datasets = None
        self.timestamp = timestamp
        self.timedeltas = [np.nan] * n if timedeltas is None else timedeltas
        self.covariates = covariates
        cov = [None] * n if covariates is None else covariates
        mm = [None] * n if main_masks is None else main_masks
        self.fields = np.empty(n, dtype=object)
        for k in range(n):
            self.fields[k] = Field.from_cube(cubes[k], lats[k], lons[k], t_indices[k], cov[k], timestamp, mm[k])
        self.n_procs = n
        self.n_data = self._count_data()
        return self

    def distribute(self, shard) -> "MultiField":
        """Multi-GPU (not in the reference): partition the variogram pair tiles of this object over the
        ranks of `shard` (cokrig_b200.parallel.VarioShard).  Every rank must then call get_variogram /
        empirical_variograms collectively; all ranks obtain the same bits as a single-GPU run."""
        self._shard = shard
        return self

    def _apply_timedelta(self, timedelta: int) -> str:
        """Timestamp with the month offset applied, as a string."""
        from datetime import datetime
        from dateutil.relativedelta import relativedelta
        t0 = datetime.strptime(self.timestamp, "%Y-%m-%d")
        return (t0 + relativedelta(months=timedelta)).strftime("%Y-%m-%d")

    def _count_data(self) -> int:
        """Total number of data values across all fields."""
        return np.sum([f.size for f in self.fields])

    def calc_dist_matrix(self, ids: tuple, units: str, fast_dist: bool, main: bool = False) -> np.ndarray:
        assert len(ids) == 2
        coord_list = [self.fields[i].coords_main if main else self.fields[i].coords for i in ids]
        return distance_matrix(*coord_list, units=units, fast_dist=fast_dist)

    def _variogram_cloud(self, i: int, j: int, config: VarioConfig) -> pd.DataFrame:
        """The (cross-)variogram cloud as a data frame (materialises every pair; kept for API
        compatibility -- ``get_variogram`` does not use it)."""
        dist = self.calc_dist_matrix((i, j), config.dist_units, config.fast_dist)
        cloud = _cloud_calc(self.fields[[i, j]] if i != j else self.fields[[i, i]], config.covariogram)
        if i == j:
            idx = np.triu_indices(dist.shape[0], k=1, m=dist.shape[1])
            dist, cloud = dist[idx], cloud[idx]
        else:
            dist, cloud = dist.flatten(), cloud.flatten()
        assert cloud.shape == dist.shape
        return pd.DataFrame({"distance": dist, "variogram": cloud})

    def get_variogram(self, i: int, j: int, config: VarioConfig) -> pd.DataFrame:
        """(Cross-)variogram of the configured kind for fields (i, j): bin centres, bin means and bin
        counts, indexed (i, j, row).  Computed on the device without materialising the cloud."""
        metric = ops.metric_id(config.dist_units, config.fast_dist)
        centers, edges, counts, sums = _device_variogram(
            self.fields[i].coords, self.fields[i].values, self.fields[j].coords, self.fields[j].values,
            i == j, metric, config.covariogram, config.max_dist, config.n_bins, shard=getattr(self, "_shard", None))
        with np.errstate(invalid="ignore", divide="ignore"):
            means = np.where(counts > 0, sums / np.maximum(counts, 1), np.nan)
        df = pd.DataFrame({"bin_center": centers, "bin_mean": means, "bin_count": counts.astype(np.int64)})
        if (df["bin_count"] < 30).any():
            warnings.warn("WARNING: Fewer than 30 pairs used for at least one bin in variogram calculation.")
        df["i"], df["j"] = i, j
        return df.set_index(["i", "j", df.index])

    def empirical_variograms(self, config: VarioConfig) -> EmpiricalVariogram:
        """Empirical variogram of every field and cross-variogram of every pair of fields.

        Returns an ``EmpiricalVariogram`` whose ``df`` is indexed by the field ids (i, j)."""
        variograms = [self.get_variogram(i, j, config)
                      for i in range(self.n_procs) for j in range(self.n_procs) if i <= j]
        return EmpiricalVariogram(pd.concat(variograms), config, self.timestamp, self.timedeltas)


# ------------------------------------------------------------------------------------------------
def _host_distance(metric: int, p, q) -> float:
    """One pair distance with the reference's formula evaluated through libm (math.*), i.e. the same
    routines sklearn's Cython haversine and scipy's cdist call: used to re-decide the few pairs whose
    device distance lies within a few ulp of a decision boundary."""
    if metric == ops.METRIC_HAVERSINE:
        la1, lo1, la2, lo2 = (float(np.radians(v)) for v in (p[0], p[1], q[0], q[1]))
        s0 = math.sin(0.5 * (la1 - la2))
        s1 = math.sin(0.5 * (lo1 - lo2))
        return 2.0 * math.asin(math.sqrt(s0 * s0 + math.cos(la1) * math.cos(la2) * s1 * s1)) * EARTH_RADIUS
    dx, dy = float(p[0]) - float(q[0]), float(p[1]) - float(q[1])
    return math.sqrt(dx * dx + dy * dy)


def _bins_from_extrema(min_dist: float, max_dist: float, n_bins: int):
    """Same numpy expressions as the reference (fields.py:394-402): first edge moved to zero."""
    centers = np.linspace(min_dist, max_dist, n_bins)
    width = centers[1] - centers[0]
    edges = np.arange(min_dist - 0.5 * width, max_dist + width, width)
    if not np.allclose((edges[1:] + edges[:-1]) / 2, centers):
        warnings.warn("WARNING: variogram bins are not centered.")
    edges[0] = 0
    return centers, edges


def _device_variogram(coords_a, values_a, coords_b, values_b, same_field, metric, covariogram, max_dist, n_bins,
                      shard=None):
    va = np.ascontiguousarray(np.asarray(values_a, dtype=float))
    vb = np.ascontiguousarray(np.asarray(values_b, dtype=float))
    ca = np.ascontiguousarray(np.asarray(coords_a, dtype=float))
    cb = np.ascontiguousarray(np.asarray(coords_b, dtype=float))
    Xa, Xb = ops.coords_to_device(ca), ops.coords_to_device(cb)
    part = {}
    if shard is not None and shard.world > 1:  # row-block partition of the pair tiles
        part = {"tile_rows": shard.tile_rows(*ops.vario_tiling(len(ca), len(cb)), same_field)}
    res = ops.vario_extrema(Xa, Xb, metric, same_field, max_dist,
                            combine=(lambda *t: shard.combine_extrema(*t, device=Xa.device)) if part else None, **part)
    if part:
        res["candidates"] = shard.gather_pairs(res["candidates"])
    mn, mx = res["min"], res["max"]
    # re-decide the extrema with libm on the candidate pairs (bit-identical to the reference's values)
    if metric == ops.METRIC_HAVERSINE and res["candidates"] is not None:
        dd = np.array([_host_distance(metric, ca[a], cb[b]) for a, b in res["candidates"]])
        ok = dd <= max_dist
        if (ok & (dd > 0)).any():
            mn = dd[ok & (dd > 0)].min()
        if ok.any():
            mx = dd[ok].max()
    if not np.isfinite(mn) or not np.isfinite(mx):
        raise ValueError("no pair of points with a positive distance within max_dist")
    centers, edges = _bins_from_extrema(mn, mx, n_bins)
    counts, sums, flagged = ops.vario_bin(Xa, ops.to_device(va), va.mean(), Xb, ops.to_device(vb), vb.mean(), metric,
                                          same_field, covariogram, max_dist, edges,
                                          combine=shard.combine_partials if part else None, **part)
    if part:
        flagged = shard.gather_pairs(flagged)
    if flagged is not None and len(flagged):
        ra, rb = va - va.mean(), vb - vb.mean()
        for a, b in flagged:  # pairs within a few ulp of an edge / max_dist: decide like the reference
            d = _host_distance(metric, ca[a], cb[b])
            if not d <= max_dist:
                continue
            k = int(np.searchsorted(edges, d, side="left"))
            if d == edges[0]:
                k = 1
            if k < 1 or k > n_bins:
                continue
            counts[k - 1] += 1
            sums[k - 1] += (ra[a] * rb[b]) if covariogram else 0.5 * (ra[a] - rb[b]) ** 2
    return centers, edges, counts, sums


def _check_length_match(*args):
    """Check that each input list has the same length."""
    if len({len(i) for i in args}) != 1:
        raise ValueError("Not all lists have the same length")


def _get_field_names(ds):
    """Data and estimated-variance variable names of a dataset."""
    var_name = [name for name in list(ds.keys()) if "_var" in name][0]
    return var_name.replace("_var", ""), var_name


def get_group_ids(group: pd.DataFrame):
    """The group ids as a tuple (i, j)."""
    return group.index.get_level_values("i")[0], group.index.get_level_values("j")[0]


def _median_abs_dev(x: np.ndarray) -> float:
    """Median absolute deviation scaled for a normal distribution."""
    return 1.4826 * np.nanmedian(np.abs(x - np.nanmedian(x)))


def fit_linear_trend(da):
    """Linear trend (over time) of the spatial monthly averages."""
    from xarray import DataArray
    from stat_tools import simple_linear_regression
    x = da.mean(dim=["lat", "lon"])
    return DataArray(simple_linear_regression(x.values), dims=["time"], coords={"time": da.time})


def fit_ols(ds, data_name: str, covar_names: list):
    """Mean surface by ordinary least squares on standardised covariates; also returns the fitted
    model and the standardisation statistics."""
    from sklearn.linear_model import LinearRegression
    df = (ds.to_dataframe().drop(columns=["time", f"{data_name}_var"]).dropna(subset=[data_name]).reset_index())
    if df.shape[0] == 0:  # no data
        return ds[data_name] * np.nan
    means = df[covar_names].mean(axis=0, skipna=True).values
    scales = df[covar_names].std(axis=0, skipna=True).values
    covariates = (df[covar_names] - means) / scales
    model = LinearRegression().fit(covariates, df[data_name])
    out = df[["lon", "lat"]].copy()
    out["ols_mean"] = model.predict(covariates)
    ds_pred = out.set_index(["lon", "lat"]).to_xarray().assign_coords(coords={"time": ds[data_name].time})
    return ds_pred["ols_mean"], model, means, scales


# ------------------------------------------------------------------------------------------------
# The preprocessing chain on plain arrays (no xarray): same steps, same library calls as fields.py:283-375
def fit_linear_trend_array(cube: np.ndarray) -> np.ndarray:
    """``fit_linear_trend`` (fields.py:283-287): linear trend over time of the spatial mean of every time slice
    (NaN cells skipped, as xarray's ``mean`` does); returns the trend value per time step."""
    from stat_tools import simple_linear_regression
    cube = np.asarray(cube, dtype=float)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)  # all-NaN slices -> NaN, like xarray
        x = np.nanmean(cube.reshape(cube.shape[0], -1), axis=1)
    return simple_linear_regression(x)


def fit_ols_frame(df: pd.DataFrame, data_name: str, covar_names: list):
    """``fit_ols`` (fields.py:290-315) on a data frame with columns lon, lat, `data_name` and the covariates: OLS of
    the data on the STANDARDISED covariates (sample std, ddof = 1, like pandas).  Returns (frame [lon, lat, ols_mean]
    of the rows with data, fitted sklearn model, covariate means, covariate scales)."""
    from sklearn.linear_model import LinearRegression
    d = df.dropna(subset=[data_name]).reset_index(drop=True)
    if d.shape[0] == 0:
        return d[["lon", "lat"]].assign(ols_mean=np.nan), None, None, None
    means = d[covar_names].mean(axis=0, skipna=True).values
    scales = d[covar_names].std(axis=0, skipna=True).values
    covariates = d[covar_names].copy()
    for i, name in enumerate(covar_names):
        covariates[name] = (covariates[name] - means[i]) / scales[i]
    model = LinearRegression().fit(covariates, d[data_name])
    out = d[["lon", "lat"]].copy()
    out["ols_mean"] = model.predict(covariates)
    return out, model, means, scales


def preprocess_arrays(cube, lat, lon, t_index: int, covariates=None) -> dict:
    """``_preprocess_ds`` (fields.py:345-375) on a (time, lat, lon) cube: (1) subtract the linear temporal trend of the
    spatial means, (2) take slice `t_index`, (3) subtract the OLS spatial trend fitted on standardised covariates,
    (4) standardise the residuals (nanmean / nanstd).  ``covariates``: None or a list of names among "lon", "lat"
    (default ["lon", "lat"], what the reference uses when no covariate dataset is given), or a dict name -> (lat, lon)
    array.  Returns {"frame": rows ordered (lon, lat) with columns lon, lat, value (standardised residual, NaN where
    missing), spatial_trend and the covariates; "attrs": temporal_trend, spatial_model, covariate_means,
    covariate_scales, spatial_mean, scale_fact; "covariate_names"}."""
    cube = np.asarray(cube, dtype=float)
    lat, lon = np.asarray(lat, dtype=float), np.asarray(lon, dtype=float)
    trend = fit_linear_trend_array(cube)
    field = cube[t_index] - trend[t_index]                      # (lat, lon)
    LON, LAT = np.meshgrid(lon, lat, indexing="ij")              # (lon, lat): row order of a (lon, lat, time) dataset
    df = pd.DataFrame({"lon": LON.ravel(), "lat": LAT.ravel(), "value": field.T.ravel()})
    if covariates is None:
        names = ["lon", "lat"]
    elif isinstance(covariates, dict):
        names = list(covariates)
        for name in names:
            df[name] = np.asarray(covariates[name], dtype=float).T.ravel()
    else:
        names = list(covariates)
    ols, model, means, scales = fit_ols_frame(df, "value", names)
    df = df.merge(ols.rename(columns={"ols_mean": "spatial_trend"}), on=["lon", "lat"], how="left")
    resid = df["value"].values - df["spatial_trend"].values
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        mean, scale = np.nanmean(resid), np.nanstd(resid)
    df["value"] = (resid - mean) / scale
    attrs = {"temporal_trend": trend[t_index], "spatial_model": model, "covariate_means": means, "covariate_scales": scales,
             "spatial_mean": mean, "scale_fact": scale}
    return {"frame": df, "attrs": attrs, "covariate_names": names}


def distance_matrix(X1: np.ndarray, X2: np.ndarray, units: str = "km", fast_dist: bool = False) -> np.ndarray:
    """Pairwise distances between two coordinate sets (rows [lat, lon]; [x, y] when ``units is None``).

    fast_dist=True: great-circle (haversine) kilometres; units=None: Euclidean; both on the device.
    Otherwise geodesic distances through geopy (host callback, as in the reference; never used at scale)."""
    X1 = np.atleast_2d(X1)
    X2 = np.atleast_2d(X2)
    if fast_dist or units is None:
        metric = ops.metric_id(units, fast_dist)
        return ops.distance_block(ops.coords_to_device(X1), ops.coords_to_device(X2), metric).cpu().numpy()
    from geopy.distance import geodesic
    from scipy.spatial.distance import cdist
    return cdist(X1, X2, lambda s_i, s_j: getattr(geodesic(s_i, s_j), units))


def _preprocess_ds(ds, timestamp: str, covariates: list):
    """Temporal detrending, OLS spatial trend removal and standardisation of one dataset (host, xarray)."""
    data_name, _ = _get_field_names(ds)
    ds_copy = ds.copy()
    ds_copy["temporal_trend"] = fit_linear_trend(ds_copy[data_name])
    ds_copy[data_name] = ds_copy[data_name] - ds_copy["temporal_trend"]
    ds_field = ds_copy.sel(time=timestamp)
    ds_field.attrs["temporal_trend"] = ds_field["temporal_trend"].values
    (ds_field["spatial_trend"], ds_field.attrs["spatial_model"], ds_field.attrs["covariate_means"],
     ds_field.attrs["covariate_scales"]) = fit_ols(ds_field, data_name, covariates)
    ds_field[data_name] = ds_field[data_name] - ds_field["spatial_trend"]
    ds_field.attrs["spatial_mean"] = np.nanmean(ds_field[data_name].values)
    ds_field.attrs["scale_fact"] = np.nanstd(ds_field[data_name].values)
    ds_field[data_name] = (ds_field[data_name] - ds_field.attrs["spatial_mean"]) / ds_field.attrs["scale_fact"]
    return ds_field


def _cloud_calc(fields: list, covariogram: bool) -> np.ndarray:
    """Semivariogram or covariogram cloud values for all point pairs (host, O(n^2) memory)."""
    ra, rb = (f.values - f.values.mean() for f in fields)
    if covariogram:
        return np.multiply.outer(ra, rb)
    return 0.5 * (np.subtract.outer(ra, rb)) ** 2


def _construct_variogram_bins(df_cloud: pd.DataFrame, n_bins: int):
    """Partition the lag domain of a variogram cloud into `n_bins` bins; first bin extended to zero."""
    d = df_cloud["distance"].values
    return _bins_from_extrema(d[d > 0].min(), d.max(), n_bins)

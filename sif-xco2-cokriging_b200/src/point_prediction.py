"""Drop-in replacement for the reference's ``src/point_prediction.py`` (local-neighbourhood simple cokriging).

Signatures follow /root/reference/src/point_prediction.py:21-355.  The per-target Python loop of the
reference (``DataFrame.apply`` -> ``_local_prediction``: neighbour masks, ``np.ix_`` gathers, two
k x k Cholesky factorisations; :127-249) is replaced by ONE batched device call
(``ck_local_count`` + ``ck_local_predict``): one CTA per target selects the neighbours, re-computes
the local covariance from coordinates, factors and solves.  ``partitions`` is accepted and ignored
(the reference forks a multiprocessing.Pool, :69-81; a single device batch needs no processes).

Same outputs and corner cases: ``(nan, nan)`` with a warning when no datum lies within ``max_dist``
or when the local matrix is not positive definite; ``pred_err = nanmax(sqrt(c0 - w.c), 0)``;
"Invalid model" warning when the augmented matrix is not PD (kriging variance <= 0, the Schur
complement of the reference's (1+k)^2 check -- or the local matrix itself not PD).

Which parameters are used where follows the reference exactly: the joint covariance ``Sigma`` is
frozen when the Predictor is constructed (:42) -- the parameter values are snapshotted in
``__init__`` and both the lazily assembled host blocks ``Sigma`` ("00", "01", "11") and the local
matrices inside the kernel are built from that snapshot -- while ``c0`` (:66) and the
target-to-neighbour covariance vector (:115-125) use the model's parameters at call time.
"""
from __future__ import annotations

import warnings
from multiprocessing import cpu_count

import numpy as np
import pandas as pd

from _backend import ops
from fields import MultiField, distance_matrix  # noqa: F401
from model import MultivariateMatern

# number of partitions for parallel computation (kept for API compatibility; unused on the device)
NCORES = cpu_count()


class Predictor:
    """Multivariate prediction framework."""

    def __init__(self, mod: MultivariateMatern, mf: MultiField, covariates=None, dist_units: str = "km",
                 fast_dist: bool = True) -> None:
        if mod.n_procs != mf.n_procs:
            raise ValueError("Number of theoretical processes different from empirical processes.")
        self.n_procs = mod.n_procs
        self.mod = mod
        self.mf = mf
        self.covariates = covariates
        self.dist_units = dist_units
        self.fast_dist = fast_dist
        self._Sigma = None
        self._sigma_params = np.array(mod.params.get_values(), dtype=float)  # Sigma is frozen here (reference :42)
        self._sigma_dev = None  # the joint covariance on the device (built on first use), gathered by the kernel
        self.cv = False  # placeholder for cross-validation

    def _metric(self) -> int:
        return ops.metric_id(self.dist_units, self.fast_dist)

    @property
    def Sigma(self) -> dict:
        """Upper-triangular block dict of the joint covariance on ``coords_main`` (host arrays)."""
        if self._Sigma is None:
            self._Sigma = self._cov_blocks()
        return self._Sigma

    # largest stacked size for which the N x N joint covariance is kept on the device (8 N^2 bytes: 8.6 GB at the limit);
    # above it the kernel re-computes the local matrices from the coordinates instead of gathering them
    sigma_device_max_n = 32768

    def _device_sigma(self, coords_d):
        """The device counterpart of the reference's stored ``Sigma`` (:98-113): assembled once by K1 from the
        parameters frozen at construction, then gathered by ck_local_predict for every target."""
        n = sum(int(c.shape[0]) for c in coords_d)
        if n == 0 or n > self.sigma_device_max_n:
            return None
        if self._sigma_dev is None:
            self._sigma_dev = ops.joint_cov(coords_d, self._sigma_params, self.n_procs, self._metric())
        return self._sigma_dev

    def _cov_blocks(self) -> dict:
        """Each block of the block-covariance matrix (within a process or between processes)."""
        import copy
        p = copy.deepcopy(self.mod.params).set_values(self._sigma_params)  # the snapshot taken at construction
        metric = self._metric()
        X = [ops.coords_to_device(np.asarray(f.coords_main, dtype=float)) for f in self.mf.fields]
        blocks = dict()
        for i in range(self.n_procs):
            for j in range(i, self.n_procs):
                if i == j:
                    b = ops.matern_block(X[i], X[i], metric, p.sigma.values[i, i] ** 2, p.nu.values[i, i],
                                         p.len_scale.values[i, i], p.nugget.values[i, i], symmetric=True)
                else:
                    scale = p.rho.values[i, j] * float(np.nanprod(p.sigma.values))
                    b = ops.matern_block(X[i], X[j], metric, scale, p.nu.values[i, j], p.len_scale.values[i, j], 0.0)
                blocks[f"{i}{j}"] = b.cpu().numpy()
        return blocks

    # ---- the reference's per-target helpers (src/point_prediction.py:115-222), kept with their signatures and return
    # values for callers that use them directly.  They are NOT on the prediction path (`_predict_chunk` hands all targets
    # to one batched device call); every number they return still comes from the device: distances from
    # ck_distance_block, covariances from ck_matern_eval / ck_matern_block, the factorisation from ck_potrf, the local
    # solve from ck_local_predict with a single target.
    def _pred_cov(self, dists: list) -> list:
        """Covariance (own process, with nugget) and cross-covariance vectors for the local distances of one target."""
        return [self.mod.covariance(self.i, dists[j], use_nugget=True) if j == self.i
                else self.mod.cross_covariance(self.i, j, dists[j]) for j in range(self.n_procs)]

    def _local_dist_ix(self, s0: np.ndarray, max_dist: float) -> tuple[list, list]:
        """Per dataset: boolean mask of the data within `max_dist` of `s0` (the datum at `s0` itself excluded from
        process `i` during cross-validation) and the retained distances."""
        dists = [distance_matrix(s0, f.coords_main, units=self.dist_units, fast_dist=self.fast_dist) for f in self.mf.fields]
        keep = [d <= max_dist for d in dists]
        if self.cv:
            keep[self.i] = (dists[self.i] > 0) & (dists[self.i] <= max_dist)
        return [k.squeeze() for k in keep], [d[k] for d, k in zip(dists, keep)]

    def _local_values(self, s0: np.ndarray, max_dist: float) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(local covariance vector, local covariance matrix, local data vector) for the target `s0`."""
        ix, local_dists = self._local_dist_ix(s0, max_dist)
        local_data = np.hstack([np.asarray(self.mf.fields[i].values_main)[ix[i]] for i in range(self.n_procs)])
        rows = []
        for i in range(self.n_procs):
            row = []
            for j in range(self.n_procs):
                blk = self.Sigma[f"{i}{j}"][np.ix_(ix[i], ix[j])] if i <= j else self.Sigma[f"{j}{i}"][np.ix_(ix[j], ix[i])].T
                row.append(blk)
            rows.append(row)
        local_cov = np.block(rows)
        assert local_data.shape[0] == local_cov.shape[0]
        return np.hstack(self._pred_cov(local_dists)), local_cov, local_data

    def _verify_model(self, c0: float, local_pred_cov: np.ndarray, local_cov: np.ndarray):
        """Raises LinAlgError unless [[c0, c^T], [c, Sigma_loc]] is positive definite (device Cholesky)."""
        from scipy.linalg import LinAlgError
        aug = np.vstack([np.hstack([c0, local_pred_cov]), np.column_stack([local_pred_cov, local_cov])])
        factor = ops.potrf(ops.to_device(np.ascontiguousarray(aug, dtype=float)))
        if factor.info != 0:
            raise LinAlgError(f"{factor.info}-th leading minor of the array is not positive definite")

    @staticmethod
    def _pred_calc(c0: float, local_pred_cov: np.ndarray, local_cov: np.ndarray, local_data: np.ndarray) -> tuple[float, float]:
        """Prediction and standard deviation from explicit local quantities: device Cholesky of `local_cov`, one
        triangular solve with [c ; z] (v = L^-1 c, y = L^-1 z), pred = v.y, sd = nanmax(sqrt(c0 - v.v), 0); (nan, nan)
        with a warning if the matrix is not positive definite."""
        k = int(np.asarray(local_data).size)
        factor = ops.potrf(ops.to_device(np.ascontiguousarray(local_cov, dtype=float)))
        if factor.info != 0:
            warnings.warn("Local covariance matrix not positive definte; returning NaN.")
            return np.nan, np.nan
        import torch
        rhs = torch.empty((2, ops.padded_ld(k)), dtype=torch.float64, device="cuda")[:, :k]
        rhs.copy_(torch.from_numpy(np.vstack([np.asarray(local_pred_cov, dtype=float).reshape(1, k),
                                               np.asarray(local_data, dtype=float).reshape(1, k)])))
        v = factor.solve_lower(rhs).cpu().numpy()
        with np.errstate(invalid="ignore"):
            pred_std = np.sqrt(c0 - float(v[0] @ v[0]))
        return float(v[0] @ v[1]), float(np.nanmax([pred_std, 0.0]))

    def _local_prediction(self, s0: np.ndarray, c0: float, max_dist: float) -> tuple[float, float]:
        """(prediction, prediction standard deviation) at the single location `s0`."""
        row = pd.DataFrame(np.atleast_2d(np.asarray(s0, dtype=float))[:, :2], columns=["d1", "d2"])
        out = self._predict_chunk(row, c0, max_dist)
        return float(out["pred"].iloc[0]), float(out["pred_err"].iloc[0])

    def _predict_chunk(self, df_chunk: pd.DataFrame, c0: float, max_dist: float):
        """Batched local prediction for all rows of `df_chunk` (adds ``pred`` and ``pred_err``)."""
        pc = df_chunk.iloc[:, :2].values.astype(float)
        coords = [ops.coords_to_device(np.asarray(f.coords_main, dtype=float)) for f in self.mf.fields]
        values = [ops.to_device(np.asarray(f.values_main, dtype=float)) for f in self.mf.fields]
        pred, sd, k, info = ops.local_predict(coords, values, ops.coords_to_device(pc), self._sigma_params,
                                              self.n_procs, self.i, self._metric(), max_dist, cv=self.cv,
                                              params_pred=self.mod.params.get_values(), c0=float(c0),
                                              sigma=self._device_sigma(coords))
        for row in np.flatnonzero(k == 0):
            warnings.warn(f"No data within maximum distance {max_dist} at location {pc[row]}.")
        for row in np.flatnonzero(info != 0):  # augmented matrix not PD: Schur complement <= 0 or Sigma_loc itself not PD
            warnings.warn(f"Invalid model at prediction location {pc[row]}. This can happen at data locations.")
        if (info > 0).any():
            warnings.warn("Local covariance matrix not positive definte; returning NaN.")
        self.last_neighbour_counts = k
        df_chunk[["pred", "pred_err"]] = np.column_stack([pred, sd])
        return df_chunk

    def predict_frame(self, i: int, pcoords, max_dist: float = 1e3) -> pd.DataFrame:
        """DataFrame with the columns of `pcoords` plus ``pred`` and ``pred_err`` (no xarray needed)."""
        self.i = i
        c0 = self.mod.covariance(self.i, 0, use_nugget=True)[0]
        if not isinstance(pcoords, pd.DataFrame):
            pcoords = pd.DataFrame(np.atleast_2d(np.asarray(pcoords, dtype=float)), columns=["d1", "d2"])
        return self._predict_chunk(pcoords.copy(), c0, max_dist)

    def __call__(self, i: int, pcoords: pd.DataFrame, max_dist: float = 1e3, partitions: int = None,
                 postprocess: bool = True):
        """Multivariate local prediction at every location of `pcoords` ([[lat, lon]] or [[x, y]]).

        max_dist: data farther than this from a target are ignored; partitions: ignored (kept for
        compatibility); postprocess: back-transform to the original data scale.
        Returns an xarray Dataset with ``pred`` and ``pred_err``."""
        df_pred = self.predict_frame(i, pcoords, max_dist=max_dist)
        if postprocess:
            return self._postprocess_predictions(df_pred)
        ds = df_pred.set_index(pcoords.columns.values.tolist()).to_xarray()
        try:
            np.isnan(self.mf.fields[self.i].timestamp)
            return ds
        except TypeError:
            return ds.assign_coords(coords={"time": np.datetime64(self.mf.fields[self.i].timestamp)})

    def _postprocess_predictions(self, df: pd.DataFrame):
        """Back-transform to the original data scale (shared with the joint predictor; host, xarray)."""
        import joint_prediction
        return joint_prediction.Predictor._postprocess_predictions(self, df)

    def cross_validation_frame(self, i: int, max_dist: float = 1e3) -> pd.DataFrame:
        """LOOCV on the standardised scale: columns d1, d2, data, pred, residual, pred_err."""
        self.cv = True
        f = self.mf.fields[i]
        data = pd.DataFrame(np.hstack((f.coords_main, np.atleast_2d(f.values_main).T)), columns=["d1", "d2", "data"])
        df = self.predict_frame(i, data[["d1", "d2"]], max_dist=max_dist)
        df["data"] = data["data"].values
        df["residual"] = df["data"] - df["pred"]
        return df[["d1", "d2", "data", "pred", "residual", "pred_err"]]

    def cross_validation(self, i: int, max_dist: float = 1e3, partitions: int = None, postprocess: bool = True):
        """Leave-one-out cross-validation at each data location of process i (value withheld via the
        ``0 < d`` neighbour rule, :140-142).  Returns a data frame with the prediction residuals."""
        df = self.cross_validation_frame(i, max_dist=max_dist)
        if not postprocess:
            return df
        df = df.rename(columns={"d1": "lat", "d2": "lon"})
        ds = self._postprocess_predictions(df[["lat", "lon", "pred", "pred_err"]].copy())
        out = (ds.to_dataframe().reset_index().dropna(subset=["pred"])
               .merge(df[["lat", "lon", "data"]], on=["lat", "lon"], how="outer"))
        out["residual"] = out["data"] - out["pred"]
        return out[["lat", "lon", "data", "pred", "residual", "pred_err"]]


def prediction_coords(extents: tuple = (-125, -65, 22, 58), lon_res: float = 0.5, lat_res: float = 0.5) -> np.ndarray:
    """Prediction coordinates (land only)."""
    from data_utils import GridConfig, land_grid
    grid = GridConfig(extents=extents, lon_res=lon_res, lat_res=lat_res)
    return land_grid(grid).reset_index()[["lat", "lon"]]

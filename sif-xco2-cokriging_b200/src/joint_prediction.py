"""Drop-in replacement for the reference's ``src/joint_prediction.py`` (global-neighbourhood simple cokriging).

Signatures follow /root/reference/src/joint_prediction.py:13-283.  The hot path of
``Predictor.__call__`` (assemble C_dp and Sigma, Cholesky, solve, predict, variance; :50-78) runs
entirely on the device:

    ck_joint_cov -> ck_potrf -> ck_cross_cov -> ck_potrs_predict

The N x N matrix is never brought to the host.  Differences in *how* (not what) is computed:
  * one triangular solve instead of cho_solve + two matmuls: pred = (L^-1 c).(L^-1 z),
    var = c0 - |L^-1 c|^2, c0 = sigma_i^2 + nugget_i (the diagonal of the reference's ``pred_cov``);
  * the reference's ``_verify_model`` factors the augmented (m+N)^2 matrix only to emit a warning
    (:60-66, :260-274).  Block Cholesky: that matrix is PD  <=>  Sigma is PD and the m x m Schur
    complement C_pp - V V^T (V = L^-1 C_dp, already computed by the solve) is PD.  For
    m <= ``Predictor.verify_max_targets`` (4096) the Schur complement is formed and factored on the
    device (m^2 N + m^3/3 flops instead of (m+N)^3/3) -- the exact test.  Above that size only the
    necessary condition "every predictive variance > 0" (the diagonal of the Schur complement) is
    checked: a non-PD Schur complement with a positive diagonal goes undetected there.  When Sigma
    itself is not PD the warning is emitted and LinAlgError raised, like the reference;
  * ``cross_validation`` uses the closed-form leave-one-out identities from ONE factorisation
    (pred_k = z_k - (S^-1 z)_k / (S^-1)_kk, sd_k = (S^-1)_kk^-1/2) instead of n_i re-assemblies
    and re-factorisations (:207-257); equal to the reference loop up to rounding.
``predict_frame`` / ``cross_validation_frame`` return plain DataFrames (no xarray needed).
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
from scipy.linalg import LinAlgError

from _backend import ops
from fields import MultiField, distance_matrix  # noqa: F401
from model import MultivariateMatern


INVALID_MODEL = "Prediction joint covariance matrix is not positive definte; model technically invalid."


class Predictor:
    """Multivariate prediction framework."""

    # largest number of targets for which the exact positive-definiteness test of the augmented matrix is run
    # (Cholesky of the m x m Schur complement); above it only the diagonal (predictive variances) is tested
    verify_max_targets = 4096

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

    # -- device pipeline -----------------------------------------------------------------------
    def _metric(self) -> int:
        return ops.metric_id(self.dist_units, self.fast_dist)

    def _device_data(self, cv_ix: int = None):
        coords = [np.asarray(f.coords_main, dtype=float) for f in self.mf.fields]
        values = [np.asarray(f.values_main, dtype=float).copy() for f in self.mf.fields]
        if cv_ix is not None:
            coords[self.i] = np.delete(coords[self.i], cv_ix, axis=0)
            values[self.i] = np.delete(values[self.i], cv_ix, axis=0)
        return [ops.coords_to_device(c) for c in coords], ops.to_device(np.hstack(values))

    def _solve(self, pcoords: np.ndarray, cv_ix: int = None):
        """(pred, var, valid) at the rows of `pcoords`: numpy arrays and whether the reference's augmented matrix
        [[C_pp, C_dp^T], [C_dp, Sigma]] is positive definite (None when not tested: cross-validation calls).
        Raises LinAlgError if Sigma is not PD."""
        p = self.mod.params
        params = p.get_values()
        metric = self._metric()
        coords_d, z_d = self._device_data(cv_ix)
        sigma = ops.joint_cov(coords_d, params, self.n_procs, metric)
        factor = ops.potrf(sigma)
        pc_d = ops.coords_to_device(pcoords)
        cpd = ops.cross_cov(coords_d, pc_d, params, self.n_procs, self.i, metric)
        c0 = p.sigma.values[self.i, self.i] ** 2 + p.nugget.values[self.i, self.i]
        pred, var = factor.predict(cpd, z_d, c0)
        if factor.info != 0:
            if cv_ix is None:  # the reference's _verify_model fails first (warning), then cho_factor raises (:60-73)
                warnings.warn(INVALID_MODEL)
            factor.raise_if_failed()
        pred, var = pred.cpu().numpy(), var.cpu().numpy()
        valid = None
        if cv_ix is None:
            valid = not (var <= 0.0).any()  # diagonal of the Schur complement: necessary
            m = pc_d.shape[0]
            if valid and m <= self.verify_max_targets:  # exact: Cholesky of C_pp - V V^T
                cpp = ops.matern_block(pc_d, pc_d, metric, p.sigma.values[self.i, self.i] ** 2, p.nu.values[self.i, self.i],
                                       p.len_scale.values[self.i, self.i], p.nugget.values[self.i, self.i], symmetric=True)
                valid = factor.schur_info(cpd[:m], cpp) == 0
        return pred, var, valid

    def predict_frame(self, i: int, pcoords, cv_ix: int = None) -> pd.DataFrame:
        """Predictions and prediction standard errors as a DataFrame: the columns of `pcoords`
        followed by ``pred`` and ``pred_err`` (what the reference builds at :76-78)."""
        self.i = i
        if cv_ix is not None:
            pcoords = pd.DataFrame({"d1": np.atleast_1d(pcoords)[0], "d2": np.atleast_1d(pcoords)[1]}, index=[0])
        elif not isinstance(pcoords, pd.DataFrame):
            pcoords = pd.DataFrame(np.atleast_2d(np.asarray(pcoords, dtype=float)), columns=["d1", "d2"])
        pred, var, valid = self._solve(pcoords.values.astype(float), cv_ix=cv_ix)
        if valid is False:
            warnings.warn(INVALID_MODEL)
        df_pred = pcoords.copy()
        df_pred["pred"] = pred
        with np.errstate(invalid="ignore"):
            df_pred["pred_err"] = np.nan_to_num(np.sqrt(var))
        return df_pred

    def __call__(self, i: int, pcoords: pd.DataFrame, postprocess: bool = True, cv_ix: int = None):
        """Multivariate prediction at every location of `pcoords` ([[lat, lon]] or [[x, y]]).

        i: process to predict; postprocess: transform back to the scale of the original data using
        the MultiField / covariates; cv_ix: index of the datum withheld for cross-validation.
        Returns an xarray Dataset with ``pred`` and ``pred_err`` (as the reference does)."""
        df_pred = self.predict_frame(i, pcoords, cv_ix=cv_ix)
        if postprocess:
            df_pred.rename(columns={"d1": "lat", "d2": "lon"}, inplace=True)
            return self._postprocess_predictions(df_pred)
        coord_cols = [c for c in df_pred.columns if c not in ("pred", "pred_err")]
        ds = df_pred.set_index(coord_cols).to_xarray()
        try:
            np.isnan(self.mf.fields[self.i].timestamp)
            return ds
        except TypeError:
            return ds.assign_coords(coords={"time": np.datetime64(self.mf.fields[self.i].timestamp)})

    # -- host-side views of the device blocks (reference API) ------------------------------------
    def _pred_cov(self, pcoords: np.ndarray) -> np.ndarray:
        """Variance-covariance matrix of the prediction locations for the current process (m x m)."""
        p = self.mod.params
        X = ops.coords_to_device(np.asarray(pcoords, dtype=float))
        return ops.matern_block(X, X, self._metric(), p.sigma.values[self.i, self.i] ** 2, p.nu.values[self.i, self.i],
                                p.len_scale.values[self.i, self.i], p.nugget.values[self.i, self.i],
                                symmetric=True).cpu().numpy()

    def _pred_cross_cov(self, pcoords: np.ndarray, cv_ix: int = None) -> np.ndarray:
        """(N x m) covariance and cross-covariance vectors between data and prediction locations."""
        coords_d, _ = self._device_data(cv_ix)
        cpd = ops.cross_cov(coords_d, ops.coords_to_device(np.asarray(pcoords, dtype=float)),
                            self.mod.params.get_values(), self.n_procs, self.i, self._metric(), spare_rows=0)
        return np.ascontiguousarray(cpd.cpu().numpy().T)

    def _joint_cov(self, cv_ix: int = None) -> np.ndarray:
        """Block covariance matrix of the stacked data (N x N), brought to the host."""
        coords_d, _ = self._device_data(cv_ix)
        return ops.joint_cov(coords_d, self.mod.params.get_values(), self.n_procs, self._metric()).cpu().numpy()

    # -- back-transform ------------------------------------------------------------------------------
    def postprocess_frame(self, df: pd.DataFrame) -> pd.DataFrame:
        """``_postprocess_predictions`` (:155-205) without xarray: `df` has columns lat, lon, pred, pred_err on the
        standardised scale; returns the same columns on the original data scale -- pred * scale_fact + spatial_mean +
        OLS trend(standardised covariates) + temporal trend, pred_err * scale_fact.  ``self.covariates``: None (lon / lat
        are the covariates) or a data frame with columns lon, lat and one column per covariate of the fit."""
        field = self.mf.fields[self.i]
        attrs = field.attrs if getattr(field, "ds", None) is None else field.ds.attrs
        out = df[["lat", "lon", "pred", "pred_err"]].copy()
        out["pred"] = out["pred"] * attrs["scale_fact"] + attrs["spatial_mean"]
        out["pred_err"] = out["pred_err"] * attrs["scale_fact"]
        if self.covariates is None:
            covariates = out[["lon", "lat"]].copy()
        else:
            names = [c for c in self.covariates.columns if c not in ("lon", "lat")]
            out = out.merge(self.covariates, on=["lon", "lat"], how="left").dropna(subset=names).reset_index(drop=True)
            covariates = out[names].copy()
        for k, name in enumerate(covariates):
            covariates[name] = (covariates[name] - attrs["covariate_means"][k]) / attrs["covariate_scales"][k]
        out["pred"] = out["pred"] + attrs["spatial_model"].predict(covariates) + attrs["temporal_trend"]
        return out[["lat", "lon", "pred", "pred_err"]]

    def _postprocess_predictions(self, df: pd.DataFrame):
        """Convert prediction results to a dataset on the original data scale: undo the
        standardisation, add the OLS spatial trend and the temporal trend (host-only, xarray)."""
        field = self.mf.fields[self.i]
        attrs = field.ds.attrs
        ds = df.set_index(["lon", "lat"]).to_xarray()
        locs = df[["lon", "lat"]].copy()
        ds *= attrs["scale_fact"]
        ds["pred"] += attrs["spatial_mean"]
        if self.covariates is None:
            covariates = df[["lon", "lat"]].copy()
        else:
            ds["covariates"] = self.covariates.sel(time=field.timestamp)
            merged = (ds.to_dataframe().reset_index().merge(locs, on=["lon", "lat"], how="right")
                      .dropna(subset=["covariates"]))
            locs = merged[["lon", "lat"]].copy()
            covariates = merged[["covariates"]].copy()
        for k, name in enumerate(covariates):
            covariates[name] = (covariates[name] - attrs["covariate_means"][k]) / attrs["covariate_scales"][k]
        locs["ols_mean"] = attrs["spatial_model"].predict(covariates)
        trend = (locs.set_index(["lon", "lat"]).to_xarray()
                 .assign_coords(coords={"time": np.datetime64(field.timestamp)}))
        ds["pred"] += trend["ols_mean"]
        ds["pred"] += attrs["temporal_trend"]
        return ds

    # -- leave-one-out cross-validation ------------------------------------------------------------
    def cross_validation_frame(self, i: int) -> pd.DataFrame:
        """Closed-form LOOCV on the standardised scale: columns d1, d2, data, pred, residual, pred_err."""
        self.i = i
        params = self.mod.params.get_values()
        metric = self._metric()
        coords_d, z_d = self._device_data()
        n = int(z_d.numel())
        sizes = [int(f.coords_main.shape[0]) for f in self.mf.fields]
        start = int(sum(sizes[:i]))
        ni = sizes[i]
        factor = ops.potrf(ops.joint_cov(coords_d, params, self.n_procs, metric))
        # rows of the identity for the data of process i: V_k = L^-1 e_k  ->  (S^-1)_kk = |V_k|^2, (S^-1 z)_k = V_k . y
        import torch
        rhs = torch.zeros((ni + 1, ops.padded_ld(n)), dtype=torch.float64, device=z_d.device)[:, :n]
        rhs[:ni, start:start + ni].fill_diagonal_(1.0)
        sz, neg_diag = factor.predict(rhs, z_d, 0.0)
        factor.raise_if_failed()
        sz, diag = sz.cpu().numpy(), -neg_diag.cpu().numpy()
        z = np.asarray(self.mf.fields[i].values_main, dtype=float)
        pred = z - sz / diag
        with np.errstate(invalid="ignore", divide="ignore"):
            err = np.nan_to_num(np.sqrt(1.0 / diag))
        c = np.asarray(self.mf.fields[i].coords_main, dtype=float)
        return pd.DataFrame({"d1": c[:, 0], "d2": c[:, 1], "data": z, "pred": pred, "residual": z - pred,
                             "pred_err": err})

    def cross_validation(self, i: int, postprocess: bool = True) -> pd.DataFrame:
        """Leave-one-out cross-validation at each data location of process i: prediction at each data
        location with the corresponding value withheld.  Returns a data frame with the residuals."""
        df = self.cross_validation_frame(i)
        if not postprocess:
            return df[["d1", "d2", "data", "pred", "residual", "pred_err"]]
        df = df.rename(columns={"d1": "lat", "d2": "lon"})
        ds = self._postprocess_predictions(df[["lat", "lon", "pred", "pred_err"]].copy())
        out = (ds.to_dataframe().reset_index().dropna(subset=["pred"])
               .merge(df[["lat", "lon", "data"]], on=["lat", "lon"], how="outer"))
        out["residual"] = out["data"] - out["pred"]
        return out[["lat", "lon", "data", "pred", "residual", "pred_err"]]


def _verify_model(pred_cov: np.ndarray, pred_cross_cov: np.ndarray, joint_cov: np.ndarray):
    """Positive-definiteness check of the augmented matrix [[C_pp, C_dp^T], [C_dp, Sigma]] by a device
    Cholesky; raises LinAlgError like the reference's cho_factor-based check (:260-274)."""
    import torch
    aug = np.vstack([np.hstack([pred_cov, pred_cross_cov.T]), np.hstack([pred_cross_cov, joint_cov])])
    n = aug.shape[0]
    buf = torch.empty((n, ops.padded_ld(n)), dtype=torch.float64, device=ops.require_cuda())[:, :n]
    buf.copy_(torch.from_numpy(np.ascontiguousarray(aug)))
    ops.potrf(buf).raise_if_failed()


def prediction_coords(extents: tuple = (-125, -65, 22, 58), lon_res: float = 0.5, lat_res: float = 0.5) -> np.ndarray:
    """Prediction coordinates (land only)."""
    from data_utils import GridConfig, land_grid
    grid = GridConfig(extents=extents, lon_res=lon_res, lat_res=lat_res)
    return land_grid(grid).reset_index()[["lat", "lon"]]

"""Drop-in replacement for the reference's ``src/model.py`` (multivariate Matern model).

Same public names, call signatures, parameter ordering and quirks as the reference
(/root/reference/src/model.py:16-391), but every Matern evaluation runs in the sm_100a kernels of
libcokrig_b200.so (``ck_matern_eval``): closed forms for nu in {1/2, 3/2, 5/2, 7/2}, a Temme/Steed
K_nu otherwise.  ``h`` and results stay numpy FP64 arrays, as in the reference.

Preserved behaviour:
  * flat parameter order sigma_ii, nu_ij, len_scale_ij, nugget_ii, rho_ij      (model.py:130-152)
  * covariance adds the nugget wherever ``h == 0`` exactly                     (model.py:193-197)
  * cross_covariance scales by the product of ALL marginal sigmas              (model.py:199-207)
  * semivariance adds the nugget at every lag, including 0                     (model.py:209-213)
  * fit(): with ``guess`` the optimisation starts from the model's CURRENT values and takes only
    the bounds from ``guess``; bins whose model value is exactly 0 are dropped (model.py:277-317)
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
from scipy.optimize import minimize

from _backend import ops
from fields import EmpiricalVariogram, get_group_ids  # noqa: F401  (re-exported like the reference)


class _PatternParam:
    """An n_procs x n_procs parameter matrix of which only a fixed index pattern is free."""

    _diag_offset = None  # None: diagonal only; 0: upper triangle incl. diagonal; 1: strict upper

    def __init__(self, name: str, default: float, bounds: tuple, n_procs: int = 2) -> None:
        self.name, self.default, self.bounds, self.n_procs = name, default, bounds, n_procs
        self.values = np.full((n_procs, n_procs), np.nan)
        np.fill_diagonal(self.values, default)  # every flavour starts from a filled diagonal
        if self._diag_offset is not None:
            self._triu_index = np.triu_indices(n_procs, k=self._diag_offset)
            self.values[self._index()] = default

    def _index(self):
        if self._diag_offset is None:
            return np.diag_indices(self.n_procs)
        return self._triu_index

    def _pairs(self):
        rows, cols = self._index()
        return list(zip(rows.tolist(), cols.tolist()))

    def get_names(self):
        return [f"{self.name}_{i + 1}{j + 1}" for i, j in self._pairs()]

    def get_values(self):
        return self.values[self._index()]

    def set_values(self, x: np.ndarray):
        self.values[self._index()] = x
        return self

    def reset_values(self):
        self.values[self._index()] = self.default
        return self

    def count_params(self):
        return len(self._index()[0])

    def to_dataframe(self):
        return pd.DataFrame({"name": self.get_names(), "value": self.get_values(),
                             "bounds": [self.bounds] * self.count_params()})


class MarginalParam(_PatternParam):
    """Multivariate marginal Matern covariance parameter (diagonal entries)."""
    _diag_offset = None


class CrossParam(_PatternParam):
    """Multivariate Matern covariance parameter with cross dependence (upper triangle)."""
    _diag_offset = 0


class RhoParam(CrossParam):
    """Cross dependence only (strict upper triangle); the diagonal keeps the default value."""

    def __init__(self, name: str, default: float, bounds: tuple, n_procs: int = 2) -> None:
        super().__init__(name, default, bounds, n_procs=n_procs)  # fills diagonal + upper with default
        self._triu_index = np.triu_indices(n_procs, k=1)


class MaternParams:
    """Multivariate Matern covariance parameters.

    sigma: process standard deviation; nu: smoothness; len_scale: length scale;
    nugget: squared nugget (tau^2); rho: co-located cross-correlation coefficient(s).
    """

    def __init__(self, n_procs: int = 2) -> None:
        self.n_procs = n_procs
        self.sigma = MarginalParam("sigma", 1.0, (0.4, 3.5), n_procs=n_procs)
        self.nu = CrossParam("nu", 1.5, (0.2, 3.5), n_procs=n_procs)
        self.len_scale = CrossParam("len_scale", 5e2, (1e2, 2e3), n_procs=n_procs)
        self.nugget = MarginalParam("nugget", 0.0, (0.0, 0.2), n_procs=n_procs)
        self.rho = RhoParam("rho", np.nan if n_procs == 1 else 0.0, (-1.0, 1.0), n_procs=n_procs)
        self._params = [self.sigma, self.nu, self.len_scale, self.nugget, self.rho]
        self.n_params = sum(p.count_params() for p in self._params)

    def to_dataframe(self):
        return pd.concat([p.to_dataframe() for p in self._params], ignore_index=True)

    def get_names(self):
        return self.to_dataframe()["name"].values

    def get_values(self):
        return self.to_dataframe()["value"].values

    def set_values(self, x: np.ndarray):
        if len(x) != self.n_params:
            raise ValueError("Incorrect number of parameters in input array.")
        start = 0
        for p in self._params:
            stop = start + p.count_params()
            p.set_values(x[start:stop])
            start = stop
        return self

    def reset_values(self):
        for p in self._params:
            p.reset_values()
        return self

    def get_bounds(self):
        return self.to_dataframe()["bounds"].values

    def set_bounds(self, **kwargs):
        for name, bounds in kwargs.items():
            if name not in ("sigma", "nu", "len_scale", "nugget", "rho"):
                raise AttributeError(f"`{name}` is not a valid parameter.")
            getattr(self, name).bounds = bounds
        return self


class MultivariateMatern:
    """Multivariate Matern covariance model (Gneiting et al., 2010), Rasmussen-Williams parametrisation."""

    def __init__(self, n_procs: int = 2, params: MaternParams = None) -> None:
        self.n_procs = n_procs
        self.params = MaternParams(n_procs=n_procs) if params is None else params
        self.fit_result = None

    # -- device evaluations ------------------------------------------------------------------
    def _sigma_prod(self) -> float:
        return float(np.nanprod(self.params.sigma.values))

    def correlation(self, i: int, j: int, h: np.ndarray) -> np.ndarray:
        return _matern_correlation(self.params.nu.values[i, j], self.params.len_scale.values[i, j], h)

    def covariance(self, i: int, h: np.ndarray, use_nugget: bool = True) -> np.ndarray:
        nugget = self.params.nugget.values[i, i] if use_nugget else 0.0
        return ops.matern_eval(h, self.params.sigma.values[i, i] ** 2, self.params.nu.values[i, i],
                               self.params.len_scale.values[i, i], nugget)

    def cross_covariance(self, i: int, j: int, h: np.ndarray) -> np.ndarray:
        if i > j:
            i, j = j, i  # the cross-covariance is symmetric
        scale = self.params.rho.values[i, j] * self._sigma_prod()
        return ops.matern_eval(h, scale, self.params.nu.values[i, j], self.params.len_scale.values[i, j], 0.0)

    def semivariance(self, i: int, h: np.ndarray) -> np.ndarray:
        s2 = self.params.sigma.values[i, i] ** 2
        return s2 * (1.0 - self.correlation(i, i, h)) + self.params.nugget.values[i, i]

    def cross_semivariance(self, i: int, j: int, h: np.ndarray) -> np.ndarray:
        if i > j:
            i, j = j, i
        sill = 0.5 * np.nansum(self.params.sigma.values ** 2 + self.params.nugget.values)
        return sill - self.cross_covariance(i, j, h)

    # -- variogram tables --------------------------------------------------------------------
    def get_variogram(self, i: int, j: int, h: np.ndarray, kind: str) -> pd.DataFrame:
        """Model (cross-)covariogram or (cross-)semivariogram of processes (i, j) at lags h."""
        if kind == "covariogram":
            v = self.covariance(i, h) if i == j else self.cross_covariance(i, j, h)
        else:
            v = self.semivariance(i, h) if i == j else self.cross_semivariance(i, j, h)
        df = pd.DataFrame({"distance": h, "variogram": v, "i": i, "j": j})
        return df.set_index(["i", "j", df.index])

    def variograms(self, h: np.ndarray, kind: str = "semivariogram") -> pd.DataFrame:
        """Modelled variograms and cross-variogram(s) of the given kind at separation distances h."""
        return pd.concat([self.get_variogram(i, j, h, kind)
                          for i in range(self.n_procs) for j in range(self.n_procs) if i <= j])

    # -- fit ---------------------------------------------------------------------------------
    @staticmethod
    def _weighted_least_squares(ydata: np.ndarray, yfit: np.ndarray, bin_counts: np.ndarray) -> float:
        """Cressie (1985) weighted least squares; zero model values fall back to counts * y^2."""
        ydata, yfit, bin_counts = (np.asarray(a, dtype=float) for a in (ydata, yfit, bin_counts))
        out = np.where(yfit == 0.0, bin_counts * ydata ** 2, 0.0)
        nz = yfit != 0.0
        out[nz] = bin_counts[nz] * ((ydata[nz] - yfit[nz]) / yfit[nz]) ** 2
        return np.sum(out)

    def _map_fit(self, df_group: pd.DataFrame) -> pd.DataFrame:
        """Adds a `fit` column: the semivariogram model evaluated at `bin_center`."""
        i, j = get_group_ids(df_group)
        h = df_group["bin_center"].values
        df_group["fit"] = self.semivariance(i, h) if i == j else self.cross_semivariance(i, j, h)
        return df_group

    def _model_at_bins(self, df_vario: pd.DataFrame) -> np.ndarray:
        ii = df_vario.index.get_level_values(0).values
        jj = df_vario.index.get_level_values(1).values
        h = df_vario["bin_center"].values.astype(float)
        fit = np.empty(len(df_vario))
        for i in range(self.n_procs):
            for j in range(i, self.n_procs):
                sel = (ii == i) & (jj == j)
                if sel.any():
                    fit[sel] = self.semivariance(i, h[sel]) if i == j else self.cross_semivariance(i, j, h[sel])
        return fit

    def _composite_wls(self, p, df_vario: pd.DataFrame) -> float:
        """Composite WLS cost over all (cross-)variograms; bins with model value 0 are dropped."""
        self.params.set_values(p)
        yfit = self._model_at_bins(df_vario)
        ydata = df_vario["bin_mean"].values.astype(float)
        counts = df_vario["bin_count"].values.astype(float)
        keep = yfit != 0.0
        return _wls(ydata[keep], yfit[keep], counts[keep])

    def fit(self, estimate: EmpiricalVariogram, guess: MaternParams = None):
        """Fit the parameters to the empirical (cross-)semivariograms simultaneously by composite
        weighted least squares (L-BFGS-B, extension of Cressie 1985)."""
        if estimate.config.n_procs != self.n_procs:
            raise ValueError("Number of theoretical processes different from empirical processes.")
        if guess is None:
            init_params = self.params.reset_values().get_values()
        else:
            init_params = self.params.get_values()
            self.params.set_bounds(**{p.name: p.bounds for p in guess._params})
        bounds = self.params.get_bounds()
        optim_result = minimize(self._composite_wls, init_params, args=(estimate.df,), method="L-BFGS-B",
                                bounds=list(bounds))
        if not optim_result.success:
            warnings.warn("ERROR: optimization did not converge.")
        self.params.set_values(optim_result.x)
        self.fit_result = FittedVariogram(self, estimate, optim_result.fun)
        return self


class FittedVariogram:
    """Model parameters and theoretical variogram for the corresponding empirical variogram."""

    def __init__(self, model: MultivariateMatern, estimate: EmpiricalVariogram, cost: float) -> None:
        self.config = estimate.config
        self.timestamp = estimate.timestamp
        self.timedeltas = estimate.timedeltas
        self.df_empirical = estimate.df
        h = np.linspace(0, self.df_empirical["bin_center"].max(), 100)
        self.df_theoretical = model.variograms(h)
        self.params = model.params
        self.cost = cost
        self.cs_valid = self.cs_check()

    def cs_check(self):
        """Cauchy-Schwarz validity check placeholder (the reference returns None)."""
        return None


def _mod_bessel(nu: float, h):
    """K_nu(h) through the device Matern kernel: rho(h) = 2^(1-nu)/Gamma(nu) x^nu K_nu(x) with
    len_scale chosen such that x == h (reference: scipy.special.kv, model.py:349-350)."""
    from scipy.special import gammaln
    h = np.atleast_1d(np.asarray(h, dtype=float))
    rho = ops.matern_eval(h, 1.0, nu, np.sqrt(2.0 * nu), 0.0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return rho / np.exp((1.0 - nu) * np.log(2) - gammaln(nu) + nu * np.log(h))


def _matern_correlation(nu: float, len_scale: float, h: np.ndarray) -> np.ndarray:
    r"""Matern correlation rho(h) = 2^{1-nu}/Gamma(nu) (sqrt(2 nu) h / l)^nu K_nu(sqrt(2 nu) h / l),
    with rho(0) = 1, non-finite values mapped to 0 and the result clamped at 0 (model.py:354-385)."""
    return ops.matern_eval(np.atleast_1d(np.abs(h)), 1.0, nu, len_scale, 0.0)


def _wls(ydata: np.ndarray, yfit: np.ndarray, bin_counts: np.ndarray) -> float:
    """Weighted least squares cost of Cressie (1985)."""
    return float(np.sum(bin_counts * ((ydata - yfit) / yfit) ** 2))

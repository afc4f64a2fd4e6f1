"""Drop-in replacement for the reference's ``src/sim.py`` (synthetic bivariate Gaussian random field).

Signatures follow /root/reference/src/sim.py:11-137.  The two expensive steps of
``BivariateRandomField.__init__`` -- the joint covariance [[C11, C12], [C12^T, C22]] on the grid and
its lower Cholesky factor (:42-50) -- run on the device (``ck_joint_cov`` + ``ck_potrf``, Euclidean
metric, bit-identical distances to scipy.cdist).  The random draws stay on the host with
``numpy.random.default_rng`` so that seeds reproduce the reference's fields; ``cmat`` and
``chol_fact_lower`` are exposed as host arrays like the reference (fetched lazily).
``to_fields`` builds the MultiField without xarray, in the (x, y)-sorted order the reference gets
from its outer merge + ``to_xarray`` round trip.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from _backend import ops
from fields import MultiField
from model import MultivariateMatern


class CartesianGrid:
    """Regular Cartesian grid in Euclidean space."""

    def __init__(self, xbounds: tuple = (0, 1), ybounds: tuple = (0, 1), xcount=51, ycount=51) -> None:
        xcoords = np.linspace(*xbounds, num=xcount)
        ycoords = np.linspace(*ybounds, num=ycount)
        self.coords = pd.DataFrame(self._expand_grid(xcoords, ycoords), columns=["x", "y"])
        self.count = len(self.coords)
        self._dist = None

    @property
    def dist(self) -> np.ndarray:
        """Euclidean distance matrix of the grid nodes (count x count), computed on first access."""
        if self._dist is None:
            X = ops.coords_to_device(self.coords.values)
            self._dist = ops.distance_block(X, X, ops.METRIC_EUCLID).cpu().numpy()
        return self._dist

    def _expand_grid(self, *args) -> np.ndarray:
        """All combinations of the elements of the supplied vectors, first vector varying slowest."""
        mesh = np.meshgrid(*args, indexing="ij")
        return np.stack([m.ravel() for m in mesh], axis=1)


class BivariateRandomField:
    """Simulate and sample a bivariate Gaussian random field from the supplied model."""

    def __init__(self, model: MultivariateMatern, grid: CartesianGrid, seed: int = None) -> None:
        self.seed = seed
        self.rng = np.random.default_rng(seed)
        self.mod = model
        self.grid = grid
        self.coords = grid.coords
        X = ops.coords_to_device(self.coords.values)
        sigma = ops.joint_cov([X, X], model.params.get_values(), 2, ops.METRIC_EUCLID)
        self._cmat_dev = sigma.clone()
        self._factor = ops.potrf(sigma)
        self._factor.raise_if_failed()
        self._cmat = None
        self.chol_fact_lower = self._factor.lower().cpu().numpy()
        self.fields = self._simulate()

    @property
    def cmat(self) -> np.ndarray:
        if self._cmat is None:
            self._cmat = self._cmat_dev.cpu().numpy()
        return self._cmat

    def _joint_cov_matrix(self) -> np.ndarray:
        """Joint covariance matrix of both processes on the grid (host copy)."""
        return self.cmat

    def _simulate(self) -> list:
        n = self.grid.count
        noise_vec = self.rng.standard_normal(2 * n)
        sim_data = self.chol_fact_lower @ noise_vec
        xy = self.coords.values
        return [pd.DataFrame(np.column_stack((xy, sim_data[k * n:(k + 1) * n])), columns=["x", "y", "value"])
                for k in range(2)]

    def _split_samp_coords(self, size: int, seed: int) -> list:
        """Sample locations where half the samples are co-located and half are not."""
        n_ext = int(np.floor(1.5 * size))
        n_co = int(np.ceil(size / 2))
        n_mis = size - n_co
        assert n_ext >= n_co + 2 * n_mis
        coords = self.coords.sample(n=n_ext, random_state=seed, replace=False)
        shared = coords.iloc[:n_co, :]
        own = [coords.iloc[n_co:n_co + n_mis, :], coords.iloc[n_co + n_mis:, :]]
        return [pd.concat((shared, own[k])) for k in range(2)]

    def sample(self, size: int = None, frac: float = None, epsilon: list = [0], seed: int = None) -> list:
        """size: sample size; frac: fraction of the simulated data (overrides size); epsilon:
        measurement-error std. dev. per process; seed: sampling seed (simulation seed by default)."""
        if frac is not None:
            size = int(np.ceil(frac * self.grid.count))
        assert 1.5 * size <= self.grid.count, "Sample size is too large for semi-colocated sampling scheme."
        epsilon = np.array(epsilon)
        if epsilon.size == 1:
            epsilon = np.repeat(epsilon, 2)
        if seed is not None:
            self.rng = np.random.default_rng(seed)
        else:
            seed = self.seed
        coords = self._split_samp_coords(size, seed)
        samples = [pd.merge(self.fields[k], coords[k]) for k in range(2)]
        for k, df in enumerate(samples):
            df["value"] += self.rng.normal(scale=epsilon[k], size=size)
            df.rename(columns={"value": f"Z{k}"}, inplace=True)
        return samples

    def to_xarray(self, samples: list = None):
        if samples is None:
            for k, df in enumerate(self.fields):
                df.rename(columns={"value": f"Y{k}"}, inplace=True)
            return pd.merge(*self.fields, how="outer").set_index(["x", "y"]).to_xarray()
        return pd.merge(*samples, how="outer").set_index(["x", "y"]).to_xarray()

    def to_fields(self, samples: list, i: int = None) -> MultiField:
        """Format bivariate samples as a MultiField (rows ordered by (x, y), as the reference's
        xarray round trip orders them)."""
        coords, values = [], []
        for k in range(2):
            df = samples[k].sort_values(["x", "y"], kind="mergesort").dropna(subset=[f"Z{k}"])
            coords.append(df[["x", "y"]].values)
            values.append(df[f"Z{k}"].values)
        if i is not None:
            coords, values = [coords[i]], [values[i]]
        return MultiField.from_arrays(coords, values, type="sim")

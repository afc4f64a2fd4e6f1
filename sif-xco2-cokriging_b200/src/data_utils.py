"""The part of the reference's ``src/data_utils.py`` that the hot-path modules import
(/root/reference/src/data_utils.py:122-216, 304-328, 363-372): regular grids, land mask, main
coordinates.  Host-only data shaping; file readers / preprocessors of OCO-2, MODIS and TransCom
products are out of scope.  xarray / regionmask are imported lazily.
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd


class GridConfig:
    def __init__(self, extents: tuple = None, lon_res: float = 1, lat_res: float = 1, lon_offset: float = 0,
                 lat_offset: float = 0) -> None:
        if not (lon_offset == 0 or lat_offset == 0):
            warnings.warn("Neither offset is zero.")
        if extents is None:
            extents = (-180, 180, -90, 90)
        self.extents = extents
        self.lon_res, self.lat_res = lon_res, lat_res
        self.lon_offset, self.lat_offset = lon_offset, lat_offset
        self.lon_bounds = _prep_bounds(extents[:2], lon_res, lon_offset)
        self.lat_bounds = _prep_bounds(extents[2:], lat_res, lat_offset)


class SpatialGrid:
    def __init__(self, config: GridConfig) -> None:
        """Longitude / latitude bin edges and centre points of a regular grid."""
        self.config = config
        self.lon_bins, self.lon_centers = _prep_bins(config.lon_bounds, config.lon_res)
        self.lat_bins, self.lat_centers = _prep_bins(config.lat_bounds, config.lat_res)

    def bounds_check(self, df: pd.DataFrame) -> bool:
        inside = (self.lon_bins.min() <= df.lon.min() and self.lon_bins.max() >= df.lon.max()
                  and self.lat_bins.min() <= df.lat.min() and self.lat_bins.max() >= df.lat.max())
        if not inside:
            warnings.warn("Dataset coordinates not within grid extents; may produce unexpected behavior: "
                          f"({df.lon.min()}, {df.lon.max()}, {df.lat.min()}, {df.lat.max()})")
        return inside


def _prep_bounds(bounds: tuple, res: float, offset: float) -> tuple:
    """Bounds widened by half a cell and shifted by the offset, as (lower, upper)."""
    lo, hi = bounds
    return (lo - 0.5 * res + offset, hi + 0.5 * res + offset)


def _prep_bins(bounds: tuple, res: float):
    edges = np.arange(bounds[0], bounds[1] + res, res)
    return edges, (edges[1:] + edges[:-1]) / 2


def regrid(ds=None, df: pd.DataFrame = None, config: GridConfig = None) -> pd.DataFrame:
    """Dataset -> data frame with lon / lat snapped to the centres of a regular grid."""
    if ds is not None:
        df = ds.to_dataframe().reset_index()
    elif df is None:
        warnings.warn("No data provided.")
    grid = SpatialGrid(GridConfig() if config is None else config)
    grid.bounds_check(df)
    df["lon"] = pd.cut(df.lon, grid.lon_bins, labels=grid.lon_centers).astype(float)
    df["lat"] = pd.cut(df.lat, grid.lat_bins, labels=grid.lat_centers).astype(float)
    return df


def land_grid(config: GridConfig = None) -> pd.DataFrame:
    """Land locations on a regular grid (index [lon, lat]); masks on a 0.25 degree grid first."""
    from regionmask.defined_regions import natural_earth
    fine = SpatialGrid(GridConfig(config.extents, lon_res=0.25, lat_res=0.25))
    mask = natural_earth.land_110.mask(fine.lon_centers, fine.lat_centers)
    df_mask = (regrid(ds=mask, config=config).dropna(subset=["region"]).groupby(["lon", "lat"]).mean().reset_index())
    return df_mask[["lat", "lon"]].assign(land=lambda x: 1).set_index(["lon", "lat"])


def set_main_coords(extents: tuple = None, lon_res: float = 5, lat_res: float = 4):
    """Base coordinates used as the augmentation reference."""
    grid = SpatialGrid(GridConfig((-125, -65, 22, 58) if extents is None else extents, lon_res=lon_res, lat_res=lat_res))
    return grid.lon_centers, grid.lat_centers


def get_main_coords(ds, lon_centers: np.ndarray = None, lat_centers: np.ndarray = None):
    """The dataset restricted to the base coordinates."""
    if lon_centers is None or lat_centers is None:
        lon_centers, lat_centers = set_main_coords()
    return (ds.to_dataframe().reset_index()
            .merge(pd.DataFrame({"lat": lat_centers}), on="lat", how="inner")
            .merge(pd.DataFrame({"lon": lon_centers}), on="lon", how="inner")
            .set_index(["lon", "lat", "time"]).to_xarray())


def to_xarray(coords: np.ndarray, **kwargs):
    """Data variables on [[lat, lon]] coordinates as an xarray dataset."""
    return (pd.DataFrame({**{"lat": coords[:, 0], "lon": coords[:, 1]}, **kwargs}).set_index(["lon", "lat"]).to_xarray())

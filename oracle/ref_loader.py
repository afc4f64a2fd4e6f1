"""Loader for the UNMODIFIED reference modules (test infrastructure only).

This file is part of ``oracle/`` -- it is test infrastructure, never shipped on
the product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

It imports ``/root/reference/src/*.py`` exactly as they lie on disk, using the
stub recipe of SURVEY.md section 8(c)/B.1 (xarray, numba_scipy, geopy and
regionmask are absent from this image and take no part in hot-path arithmetic).
The reference modules are registered under private names (``_ref_model`` ...) so
that they can coexist with the same-named drop-in modules of this repository.

``/root/reference`` exists only in the build container; on the GPU box
``available()`` is False and everything that needs the real reference is skipped
(the committed fixtures under ``tests/golden`` stand in for it).
"""
from __future__ import annotations

import collections
import collections.abc
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

REF_SRC = os.environ.get("COKRIG_REFERENCE_SRC", "/root/reference/src")
_NAMES = ("stat_tools", "data_utils", "fields", "model", "point_prediction",
          "joint_prediction", "sim")
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def load() -> SimpleNamespace:
    """Import the reference modules; returns a namespace (model, fields, ...)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_SRC}")

    class _D:  # placeholder for xarray types used only in annotations
        pass

    geopy = _stub("geopy")
    geopy.distance = _stub("geopy.distance", geodesic=None)
    regionmask = _stub("regionmask")
    regionmask.defined_regions = _stub("regionmask.defined_regions", natural_earth=None)
    stubs = {
        "xarray": _stub("xarray", Dataset=_D, DataArray=_D,
                        open_dataset=lambda *a, **k: None, apply_ufunc=None),
        "numba_scipy": _stub("numba_scipy"),
        "geopy": geopy, "geopy.distance": geopy.distance,
        "regionmask": regionmask,
        "regionmask.defined_regions": regionmask.defined_regions,
    }
    had_iterable = hasattr(collections, "Iterable")
    if not had_iterable:
        collections.Iterable = collections.abc.Iterable  # src/data_utils.py:3

    saved = {}
    for k in list(stubs) + list(_NAMES):
        if k in sys.modules:
            saved[k] = sys.modules.pop(k)
    for k, v in stubs.items():
        if k not in saved or k in ("numba_scipy",):
            sys.modules[k] = v
    # a real xarray / geopy / regionmask, if ever installed, wins over the stub
    for k in ("xarray", "geopy", "geopy.distance", "regionmask", "regionmask.defined_regions"):
        if k in saved:
            sys.modules[k] = saved[k]
    sys.path.insert(0, REF_SRC)
    try:
        import importlib
        mods = {n: importlib.import_module(n) for n in _NAMES}
    finally:
        sys.path.remove(REF_SRC)
        for n in _NAMES:
            m = sys.modules.pop(n, None)
            if m is not None:
                sys.modules["_ref_" + n] = m
        for k in stubs:
            if sys.modules.get(k) is stubs[k]:
                del sys.modules[k]
        sys.modules.update(saved)
        if not had_iterable:
            del collections.Iterable
    _cache = SimpleNamespace(**mods)
    return _cache


def make_multifield(ref, coords, values, coords_main=None, values_main=None):
    """Duck-typed ``fields.MultiField`` (the real ctor needs xarray Datasets).

    coords/values: one entry per process; rows are [lat, lon] degrees or [x, y].
    """
    mf = object.__new__(ref.fields.MultiField)
    coords_main = coords if coords_main is None else coords_main
    values_main = values if values_main is None else values_main
    mf.fields = np.array([
        SimpleNamespace(coords=np.asarray(c, float), coords_main=np.asarray(cm, float),
                        values=np.asarray(v, float), values_main=np.asarray(vm, float),
                        size=len(v), timestamp=np.nan)
        for c, v, cm, vm in zip(coords, values, coords_main, values_main)])
    mf.n_procs = len(mf.fields)
    mf.n_data = int(sum(f.size for f in mf.fields))
    mf.timestamp = np.nan
    mf.timedeltas = [np.nan, np.nan]
    mf.type = "sim"
    return mf

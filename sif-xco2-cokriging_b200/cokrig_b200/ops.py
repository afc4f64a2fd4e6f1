"""Host-side wrappers: numpy / torch tensors -> raw device pointers -> C ABI (include/cokrig.h).

PyTorch is used only for device memory, streams and (in ``parallel``) torch.distributed; every
arithmetic step of the hot path runs in the hand-written sm_100a kernels of libcokrig_b200.so.
There is no CPU fallback: all entry points raise ``RuntimeError`` without a CUDA device.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import METRIC_EUCLID, METRIC_HAVERSINE, check, lib

F64 = torch.float64


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("cokrig_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def metric_id(units, fast_dist: bool) -> int:
    """Map the reference's (units, fast_dist) pair (src/fields.py:318-342) to a kernel metric."""
    if fast_dist:
        return METRIC_HAVERSINE
    if units is None:
        return METRIC_EUCLID
    raise NotImplementedError(
        "geodesic distances (geopy callback, src/fields.py:337-339) are outside the device hot path; "
        "use fast_dist=True (haversine) or dist_units=None (Euclidean)")


def to_device(x, dtype=F64) -> torch.Tensor:
    """numpy / sequence / tensor -> contiguous CUDA tensor (one H2D copy for host inputs)."""
    dev = require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64 if dtype == F64 else None))
    return torch.from_numpy(a).to(dev, dtype=dtype, non_blocking=False)


def coords_to_device(X) -> torch.Tensor:
    if isinstance(X, torch.Tensor):
        t = to_device(X)
    else:
        t = to_device(np.atleast_2d(np.asarray(X, dtype=np.float64)))
    if t.dim() != 2 or t.shape[1] != 2:
        raise ValueError(f"coordinates must be (n, 2), got {tuple(t.shape)}")
    return t


def _params(values: Sequence[float], n_procs: int):
    v = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
    need = {1: 4, 2: 11}.get(n_procs)
    if need is None:
        raise NotImplementedError(f"n_procs={n_procs}: the device path implements 1 or 2 processes")
    if v.size != need:
        raise ValueError("Incorrect number of parameters in input array.")  # src/model.py:146-147
    return v, v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def padded_ld(n: int) -> int:
    """Leading dimension rounded up so that every row starts on a 128-byte line."""
    return max(16, (int(n) + 15) // 16 * 16)


# ------------------------------------------------------------------------------------------------ K1
def matern_eval(h, scale: float, nu: float, len_scale: float, nugget: float = 0.0) -> np.ndarray:
    """scale * rho(h) (+ nugget where h == 0) elementwise; numpy in, numpy out (src/model.py:188-207)."""
    require_cuda()
    h_np = np.atleast_1d(np.asarray(h, dtype=np.float64))
    hd = to_device(h_np.ravel())
    out = torch.empty_like(hd)
    check(lib.ck_matern_eval(_ptr(hd), hd.numel(), scale, nu, len_scale, nugget, _ptr(out), _stream()), "ck_matern_eval")
    return out.cpu().numpy().reshape(h_np.shape)


def distance_block(X1: torch.Tensor, X2: torch.Tensor, metric: int) -> torch.Tensor:
    n1, n2 = X1.shape[0], X2.shape[0]
    out = torch.empty((n1, n2), dtype=F64, device=X1.device)
    check(lib.ck_distance_block(_ptr(X1), n1, _ptr(X2), n2, metric, _ptr(out), max(n2, 1), _stream()), "ck_distance_block")
    return out


def matern_block(X1: torch.Tensor, X2: torch.Tensor, metric: int, scale: float, nu: float, len_scale: float,
                 nugget: float = 0.0, symmetric: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n1, n2 = X1.shape[0], X2.shape[0]
    if out is None:
        out = torch.empty((n1, n2), dtype=F64, device=X1.device)
    ld = out.stride(0) if out.dim() == 2 and n1 > 1 else max(n2, 1)
    check(lib.ck_matern_block(_ptr(X1), n1, _ptr(X2), n2, metric, scale, nu, len_scale, nugget, _ptr(out), ld,
                              _ptr(None), 0, int(symmetric), _stream()), "ck_matern_block")
    return out


def joint_cov(coords: Sequence[torch.Tensor], params, n_procs: int, metric: int,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Joint covariance [[C00, C01], [C01^T, C11]] of the stacked data; returns an (N, N) view of a
    row-padded device buffer (src/joint_prediction.py:124-153)."""
    dev = require_cuda()
    _, pp = _params(params, n_procs)
    n0 = coords[0].shape[0]
    n1 = coords[1].shape[0] if n_procs == 2 else 0
    n = n0 + n1
    if out is None:
        ld = padded_ld(n)
        out = torch.empty((n, ld), dtype=F64, device=dev)[:, :n]
    ld = out.stride(0) if n > 1 else max(n, 1)
    c1 = coords[1] if n_procs == 2 else None
    check(lib.ck_joint_cov(_ptr(coords[0]), n0, _ptr(c1), n1, pp, n_procs, metric, _ptr(out), ld, _stream()), "ck_joint_cov")
    return out


def cross_cov(coords: Sequence[torch.Tensor], pcoords: torch.Tensor, params, n_procs: int, i_pred: int, metric: int,
              spare_rows: int = 1) -> torch.Tensor:
    """(m + spare_rows, N) TARGET-MAJOR covariances between targets and stacked data
    (transpose of src/joint_prediction.py:104-122); the spare row receives z in the solve."""
    dev = require_cuda()
    _, pp = _params(params, n_procs)
    n0 = coords[0].shape[0]
    n1 = coords[1].shape[0] if n_procs == 2 else 0
    n, m = n0 + n1, pcoords.shape[0]
    ld = padded_ld(n)
    buf = torch.empty((m + spare_rows, ld), dtype=F64, device=dev)
    c1 = coords[1] if n_procs == 2 else None
    check(lib.ck_cross_cov(_ptr(coords[0]), n0, _ptr(c1), n1, _ptr(pcoords), m, pp, n_procs, i_pred, metric,
                           _ptr(buf), ld, _stream()), "ck_cross_cov")
    return buf[:, :n]


# ------------------------------------------------------------------------------------------------ K3
class CholeskyFactor:
    """Lower Cholesky factor held on the device (in the buffer of the matrix it was computed from)
    plus the inverted diagonal blocks the blocked solves reuse."""

    def __init__(self, a: torch.Tensor, ws: torch.Tensor, info: torch.Tensor):
        self.L, self.ws, self._info = a, ws, info
        self.n = a.shape[0]
        self.ld = a.stride(0) if self.n > 1 else max(self.n, 1)

    @property
    def info(self) -> int:
        """0, or k > 0 if the leading minor of order k is not positive definite (forces a sync)."""
        return int(self._info.item())

    def raise_if_failed(self):
        k = self.info
        if k != 0:
            from scipy.linalg import LinAlgError
            raise LinAlgError(f"{k}-th leading minor of the array is not positive definite")

    def lower(self) -> torch.Tensor:
        """Dense lower-triangular L (upper triangle zeroed), like scipy.linalg.cholesky(lower=True)."""
        return torch.tril(self.L)

    def solve_lower(self, rhs: torch.Tensor) -> torch.Tensor:
        """In place rhs[c, :] <- L^{-1} rhs[c, :] for target-major rhs (nrhs, n)."""
        nrhs = rhs.shape[0]
        ldr = rhs.stride(0) if nrhs > 1 else max(self.n, rhs.stride(0))
        check(lib.ck_trsm_lower(_ptr(self.L), self.n, self.ld, _ptr(self.ws), _ptr(rhs), nrhs, ldr, _stream()), "ck_trsm_lower")
        return rhs

    def predict(self, cpd: torch.Tensor, z: torch.Tensor, c0: float):
        """cpd: (m+1, n) from cross_cov (overwritten).  Returns (pred, var) device tensors (m,)."""
        m = cpd.shape[0] - 1
        pred = torch.empty(m, dtype=F64, device=cpd.device)
        var = torch.empty(m, dtype=F64, device=cpd.device)
        check(lib.ck_potrs_predict(_ptr(self.L), self.n, self.ld, _ptr(self.ws), _ptr(cpd), m, cpd.stride(0), _ptr(z),
                                   float(c0), _ptr(pred), _ptr(var), _stream()), "ck_potrs_predict")
        return pred, var

    def schur_info(self, V: torch.Tensor, cpp: torch.Tensor) -> int:
        """Positive-definiteness of the Schur complement C_pp - V V^T (V = rows L^-1 c_j, target-major (m, n); cpp the
        (m, m) target covariance, overwritten): 0 if PD, else the order of the first non-PD leading minor.  Together
        with info == 0 of this factor it decides whether the reference's augmented matrix
        [[C_pp, C_dp^T], [C_dp, Sigma]] is PD (src/joint_prediction.py:260-274) without the (m+N)^3/3 factorisation."""
        m = V.shape[0]
        if m == 0:
            return 0
        ldv = V.stride(0) if m > 1 else max(self.n, V.stride(0))
        ldc = cpp.stride(0) if m > 1 else max(m, 1)
        check(lib.ck_gemm_nt(_ptr(V), ldv, _ptr(V), ldv, _ptr(cpp), ldc, m, m, self.n, 1, _stream()), "ck_gemm_nt")
        return potrf(cpp).info

    def logdet(self) -> torch.Tensor:
        out = torch.empty(1, dtype=F64, device=self.L.device)
        check(lib.ck_logdet(_ptr(self.L), self.n, self.ld, _ptr(out), _stream()), "ck_logdet")
        return out


def potrf_workspace(n: int, device) -> torch.Tensor:
    nbytes = int(lib.ck_potrf_workspace_bytes(n))
    return torch.empty(max(nbytes // 8, 1), dtype=F64, device=device)


def potrf(a: torch.Tensor, ws: Optional[torch.Tensor] = None) -> CholeskyFactor:
    """In-place lower Cholesky of the (n, n) device tensor `a` (row stride = ld)."""
    require_cuda()
    n = a.shape[0]
    if a.dim() != 2 or a.shape[1] != n or (n > 1 and a.stride(1) != 1):
        raise ValueError("potrf needs a square row-major device matrix")
    if ws is None:
        ws = potrf_workspace(n, a.device)
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    ld = a.stride(0) if n > 1 else max(n, 1)
    check(lib.ck_potrf(_ptr(a), n, ld, _ptr(ws), _ptr(info), _stream()), "ck_potrf")
    return CholeskyFactor(a, ws, info)


def gaussian_nll(coords: Sequence[torch.Tensor], z: torch.Tensor, params, n_procs: int, metric: int,
                 sigma_buf: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None):
    """0.5 (z' S^-1 z + logdet S + N log 2pi) entirely on the device; returns (out[3] tensor, info tensor)."""
    dev = require_cuda()
    _, pp = _params(params, n_procs)
    n0 = coords[0].shape[0]
    n1 = coords[1].shape[0] if n_procs == 2 else 0
    n = n0 + n1
    if sigma_buf is None:
        sigma_buf = torch.empty((n, padded_ld(n)), dtype=F64, device=dev)
    if ws is None:
        ws = potrf_workspace(n, dev)
    scratch = torch.empty(n, dtype=F64, device=dev)
    out = torch.empty(3, dtype=F64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    c1 = coords[1] if n_procs == 2 else None
    check(lib.ck_nll(_ptr(coords[0]), n0, _ptr(c1), n1, pp, n_procs, metric, _ptr(z), _ptr(sigma_buf),
                     sigma_buf.stride(0), _ptr(ws), _ptr(scratch), _ptr(out), _ptr(info), _stream()), "ck_nll")
    return out, info


# ------------------------------------------------------------------------------------------------ K2
VARIO_GUARD = 2.0e-14  # must match CK_VARIO_GUARD (csrc/ck_vario.cu)
VARIO_LIST_CAPACITY = 1 << 16


def vario_tiling(na: int, nb: int):
    """(tile rows, tile columns, points per tile row, points per tile column) of the K2 pair tiling."""
    va, vb = ctypes.c_int(0), ctypes.c_int(0)
    check(lib.ck_vario_tile_shape(ctypes.byref(va), ctypes.byref(vb)), "ck_vario_tile_shape")
    return int(lib.ck_vario_tile_rows(na)), int(lib.ck_vario_tile_cols(nb)), va.value, vb.value


def _read_pairs(pairs: torch.Tensor, count: torch.Tensor, what: str):
    n = int(count.item())
    if n > pairs.shape[0]:
        raise RuntimeError(f"{what}: {n} ambiguous pairs exceed the list capacity {pairs.shape[0]} "
                           "(degenerate geometry: many pairs exactly on a decision boundary)")
    return pairs[:n].cpu().numpy()


def vario_extrema(Xa: torch.Tensor, Xb: torch.Tensor, metric: int, same_field: bool, max_dist: float,
                  tile_rows=(0, -1), combine=None) -> dict:
    """Device pass 1: {'min': min non-zero distance, 'max': max distance, 'count': pairs} over pairs
    with d <= max_dist, plus 'candidates': (a, b) index pairs inside the guard band of the extrema
    (haversine only; None for the exact Euclidean metric).

    tile_rows: this rank's share of the pair-tile rows ((0, -1) = all).  combine: callable that maps the
    local (min, max, count) to the global one (MIN / MAX / SUM across ranks, parallel.py); the candidate
    list is then this rank's part of the global guard band."""
    na, nb = Xa.shape[0], Xb.shape[0]
    out = torch.empty(3, dtype=F64, device=Xa.device)
    ws = torch.empty(int(lib.ck_vario_minmax_workspace_bytes(na, nb)) // 8 + 1, dtype=F64, device=Xa.device)
    check(lib.ck_vario_minmax(_ptr(Xa), na, _ptr(Xb), nb, metric, int(same_field), float(max_dist), tile_rows[0],
                              tile_rows[1], _ptr(out), _ptr(ws), _stream()), "ck_vario_minmax")
    mn, mx, cnt = out.cpu().tolist()
    if combine is not None:
        mn, mx, cnt = combine(mn, mx, cnt)
    res = {"min": mn, "max": mx, "count": int(cnt), "candidates": None}
    if metric == METRIC_HAVERSINE and cnt > 0:
        pairs = torch.empty((VARIO_LIST_CAPACITY, 2), dtype=torch.int64, device=Xa.device)
        count = torch.zeros(1, dtype=torch.int64, device=Xa.device)
        lo, hi = mn * (1.0 + 2 * VARIO_GUARD), mx * (1.0 - 2 * VARIO_GUARD)
        check(lib.ck_vario_candidates(_ptr(Xa), na, _ptr(Xb), nb, metric, int(same_field), float(max_dist), lo, hi,
                                      tile_rows[0], tile_rows[1], _ptr(ws), _ptr(pairs), VARIO_LIST_CAPACITY, _ptr(count),
                                      _stream()), "ck_vario_candidates")
        res["candidates"] = _read_pairs(pairs, count, "variogram extrema")
    return res


def vario_minmax(Xa: torch.Tensor, Xb: torch.Tensor, metric: int, same_field: bool, max_dist: float):
    r = vario_extrema(Xa, Xb, metric, same_field, max_dist)
    return r["min"], r["max"], r["count"]


def vario_bin(Xa: torch.Tensor, va: torch.Tensor, mean_a: float, Xb: torch.Tensor, vb: torch.Tensor, mean_b: float,
              metric: int, same_field: bool, covariogram: bool, max_dist: float, edges: np.ndarray, guard: bool = True,
              tile_rows=(0, -1), combine=None):
    """Device pass 2: per-bin (counts int64, sums float64, flagged) with pandas.cut(include_lowest=True)
    semantics; `flagged` = (a, b) index pairs left undecided inside the guard band (or None).

    tile_rows / combine: row-block multi-GPU partition.  Every rank bins its own pair-tile rows into the
    per-tile partial arrays (zeros elsewhere); `combine(sums_partials, counts_partials)` sums them across
    ranks in place (exact: one non-zero contributor per entry) and the fixed-order tile reduction then
    gives counts and sums that are bit-identical for any number of ranks."""
    e = np.ascontiguousarray(np.asarray(edges, dtype=np.float64))
    n_bins = e.size - 1
    na, nb = Xa.shape[0], Xb.shape[0]
    counts = torch.zeros(n_bins, dtype=torch.int64, device=Xa.device)
    sums = torch.zeros(n_bins, dtype=F64, device=Xa.device)
    use_guard = guard and metric == METRIC_HAVERSINE
    if na == 0 or nb == 0:
        return counts.cpu().numpy(), sums.cpu().numpy(), (np.empty((0, 2), dtype=np.int64) if use_guard else None)
    nbytes = int(lib.ck_vario_bin_workspace_bytes(na, nb, n_bins))
    ws = torch.empty(nbytes // 8 + 1, dtype=F64, device=Xa.device)
    pairs = torch.empty((VARIO_LIST_CAPACITY, 2), dtype=torch.int64, device=Xa.device) if use_guard else None
    count = torch.zeros(1, dtype=torch.int64, device=Xa.device)
    check(lib.ck_vario_bin_tiles(_ptr(Xa), _ptr(va), na, float(mean_a), _ptr(Xb), _ptr(vb), nb, float(mean_b), metric,
                                 int(same_field), int(covariogram), float(max_dist),
                                 e.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_bins, tile_rows[0], tile_rows[1],
                                 _ptr(pairs), VARIO_LIST_CAPACITY if use_guard else 0, _ptr(count), _ptr(ws), _stream()),
          "ck_vario_bin_tiles")
    if combine is not None:
        tiles = int(lib.ck_vario_tile_rows(na)) * int(lib.ck_vario_tile_cols(nb))
        off = int(lib.ck_vario_bin_partials_offset(n_bins)) // 8
        psum = ws[off: off + tiles * n_bins]
        off_c = off + ((tiles * n_bins * 8 + 255) // 256 * 256) // 8
        pcnt = ws[off_c: off_c + (tiles * n_bins * 4 + 7) // 8].view(torch.int32)[: tiles * n_bins]
        combine(psum, pcnt)
    check(lib.ck_vario_bin_reduce(na, nb, n_bins, _ptr(ws), _ptr(counts), _ptr(sums), _stream()), "ck_vario_bin_reduce")
    flagged = _read_pairs(pairs, count, "variogram binning") if use_guard else None
    return counts.cpu().numpy(), sums.cpu().numpy(), flagged


# ------------------------------------------------------------------------------------------------ K4
def local_predict(coords: Sequence[torch.Tensor], values: Sequence[torch.Tensor], pcoords: torch.Tensor, params,
                  n_procs: int, i_pred: int, metric: int, max_dist: float, cv: bool = False, params_pred=None,
                  c0: Optional[float] = None, sigma: Optional[torch.Tensor] = None):
    """Batched local-neighbourhood cokriging: returns (pred, sd, k, info) numpy arrays over targets
    (src/point_prediction.py:127-249).  info: 0 ok, >0 local matrix not PD (pred = sd = NaN),
    -1 valid prediction but the augmented matrix is not PD (the reference only warns).

    params: the model of the local covariance matrix (the reference freezes it at construction);
    params_pred: the model of the target-to-neighbour vector (default: params); c0: the target variance
    the caller computed (default: sigma_i^2 + nugget_i of params_pred).
    sigma: optional joint covariance of the stacked data on the device (``joint_cov(coords, params, ...)``, the
    counterpart of the reference's stored ``Predictor.Sigma``): local matrices are then gathered from it
    instead of being re-computed from the coordinates (faster; needs N^2 memory)."""
    dev = require_cuda()
    pv, pp = _params(params, n_procs)
    qv, qp = _params(params if params_pred is None else params_pred, n_procs)
    if c0 is None:
        c0 = qv[i_pred] ** 2 + qv[8 + i_pred] if n_procs == 2 else qv[0] ** 2 + qv[3]
    m = pcoords.shape[0]
    n0 = coords[0].shape[0]
    n1 = coords[1].shape[0] if n_procs == 2 else 0
    c1 = coords[1] if n_procs == 2 else None
    z1 = values[1] if n_procs == 2 else None
    k = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
    seg = torch.zeros(8 * max(m, 1), dtype=torch.int32, device=dev)
    check(lib.ck_local_count(_ptr(coords[0]), n0, _ptr(c1), n1, _ptr(pcoords), m, n_procs, i_pred, metric,
                             float(max_dist), int(cv), _ptr(k), _ptr(seg), _stream()), "ck_local_count")
    kmax = int(k[:m].max().item()) if m else 0
    nbytes = int(lib.ck_local_predict_workspace_bytes(m, kmax))
    ws = torch.empty(nbytes // 8 + 1, dtype=F64, device=dev)
    pred = torch.empty(max(m, 1), dtype=F64, device=dev)
    sd = torch.empty(max(m, 1), dtype=F64, device=dev)
    info = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
    check(lib.ck_local_predict(_ptr(coords[0]), _ptr(values[0]), n0, _ptr(c1), _ptr(z1), n1, _ptr(pcoords), m, pp, qp,
                               n_procs, i_pred, metric, float(max_dist), int(cv), float(c0), _ptr(sigma),
                               0 if sigma is None else (sigma.stride(0) if sigma.shape[0] > 1 else max(sigma.shape[1], 1)),
                               _ptr(k), _ptr(seg), kmax,
                               _ptr(pred), _ptr(sd), _ptr(info), _ptr(ws), _stream()), "ck_local_predict")
    return (pred[:m].cpu().numpy(), sd[:m].cpu().numpy(), k[:m].cpu().numpy(), info[:m].cpu().numpy())


# ------------------------------------------------------------------------------------------------ temporal statistics
def xcor_lags(Z1, Z2, lags, tau=None, detrend: bool = False):
    """Masked cross-correlation along the last axis of two equal-shape arrays for every integer lag in `lags`, in one
    device pass (src/stat_tools.py:128-160, :181-233).  Returns (xcor with shape (nlag,) + Z1.shape[:-1], index of the
    lag with the largest |xcor| per cell, that xcor) as numpy arrays."""
    dev = require_cuda()
    a = np.ascontiguousarray(np.asarray(Z1, dtype=np.float64))
    b = np.ascontiguousarray(np.asarray(Z2, dtype=np.float64))
    if a.shape != b.shape or a.ndim < 1:
        raise ValueError("Z1 and Z2 must have the same shape with time as the last axis")
    lg = np.ascontiguousarray(np.asarray(list(lags), dtype=np.int32))
    T = a.shape[-1]
    ncell = int(a.size // max(T, 1))
    ad, bd = to_device(a.reshape(ncell, T)), to_device(b.reshape(ncell, T))
    xc = torch.empty((len(lg), ncell), dtype=F64, device=dev)
    bi = torch.zeros(max(ncell, 1), dtype=torch.int32, device=dev)
    bx = torch.empty(max(ncell, 1), dtype=F64, device=dev)
    check(lib.ck_xcor_lags(_ptr(ad), _ptr(bd), ncell, T, lg.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(lg),
                           int(tau) if tau else 0, int(bool(detrend)), _ptr(xc), _ptr(bi), _ptr(bx), _stream()), "ck_xcor_lags")
    cell_shape = a.shape[:-1]
    return (xc.cpu().numpy().reshape((len(lg),) + cell_shape), bi[:ncell].cpu().numpy().reshape(cell_shape),
            bx[:ncell].cpu().numpy().reshape(cell_shape))

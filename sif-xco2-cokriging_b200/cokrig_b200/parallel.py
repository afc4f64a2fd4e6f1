"""Multi-GPU partitioning of the cokriging hot path: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch) for the plumbing, the hand-written kernels of libcokrig_b200.so for every local step.

The reference is single-process (its only parallelism is a multiprocessing.Pool over prediction
targets, src/point_prediction.py:69-81); this module is where the path shards naturally:

  VarioShard             K2: pair tiles row-block partitioned over ranks; per-tile partials combined by an
                         exact sum, then the single-GPU fixed-order tile reduction -> counts AND sums are
                         bit-identical for any number of ranks.
  shard_windows          C4: independent weekly windows round-robin over ranks (no data-path collective).
  fd_gradient            C3: the P+1 objective evaluations of one finite-difference gradient, one theta per rank.
  BlockCyclicCokriging   C5: the large joint system as ONE augmented array [Sigma ; C^T + z], square tiles,
                         2-D block-cyclic over a P x Q grid; right-looking tile Cholesky with NCCL panel
                         broadcasts and one-panel look-ahead; factor and solve in a single sweep.

Every compute step goes through a small "kernel set" object.  The product kernel set is CudaKernels
(ctypes -> libcokrig_b200.so; raises without a CUDA device -- there is no CPU fallback).  The CPU test
tier (gloo, world_size 2) injects a numpy kernel set that lives under tests/ to exercise the
partitioning, communication order and reductions without a GPU.
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

F64 = torch.float64


# ------------------------------------------------------------------------------------------------ grid
def grid_shape(world: int) -> Tuple[int, int]:
    """P x Q process grid for `world` ranks: as square as possible with P >= Q (2 -> 2 x 1, 4 -> 2 x 2, 8 -> 4 x 2).
    More process ROWS split the panel work (column update + TRSM product of every tile column, the critical path of
    the sweep) over more ranks; measured on C3: 2 x 1 250 ms vs 1 x 2 261 ms, 4 x 2 103 ms vs 2 x 4 114 ms
    (profiles/r02_mg_c3_*gpu_grid_shapes.log)."""
    q = int(np.floor(np.sqrt(world)))
    while world % q:
        q -= 1
    return world // q, q


class ProcessGrid:
    """rank = p * Q + q.  Column groups (fixed q: the P ranks of one process column) carry the diagonal-tile
    broadcast and the exchange of the panel tiles that become B operands; row groups (fixed p: the Q ranks of one
    process row) carry the panel broadcast.  A rank therefore receives only the panel rows of its own process row
    and the panel tiles of its own tile columns -- what ScaLAPACK does with row / column communicators."""

    def __init__(self, P: Optional[int] = None, Q: Optional[int] = None):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if P is None or Q is None:
            P, Q = grid_shape(self.world)
        if P * Q != self.world:
            raise ValueError(f"process grid {P}x{Q} does not match world size {self.world}")
        self.P, self.Q = P, Q
        self.p, self.q = divmod(self.rank, Q)
        self.col_groups = [None] * Q
        self.row_groups = [None] * P
        if self.world > 1 and P > 1:
            for qq in range(Q):  # every rank must create every group, in the same order
                self.col_groups[qq] = dist.new_group([pp * Q + qq for pp in range(P)])
        if self.world > 1 and Q > 1:
            for pp in range(P):
                self.row_groups[pp] = dist.new_group([pp * Q + qq for qq in range(Q)])

    def rank_of(self, p: int, q: int) -> int:
        return p * self.Q + q


def local_tiles(ntiles: int, nprocs: int, rank: int) -> int:
    """Number of tiles t in [0, ntiles) with t % nprocs == rank (== ck_mg_local_tiles)."""
    return 0 if ntiles <= rank else (ntiles - rank + nprocs - 1) // nprocs


def first_local_after(k: int, nprocs: int, rank: int) -> int:
    """Smallest local tile index l with global index l * nprocs + rank > k."""
    return (k - rank) // nprocs + 1 if k >= rank else 0


# ------------------------------------------------------------------------------------------------ C4 / C3
def shard_windows(n_windows: int, rank: int, world: int) -> List[int]:
    """Independent systems (weekly windows, BASELINE config 4) handled by `rank`: round-robin."""
    return list(range(rank, n_windows, world))


def gather_window_results(local: dict, n_windows: int) -> Optional[list]:
    """Collect {window index: result} from all ranks on rank 0, in window order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local[w] for w in range(n_windows)]
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, parts, dst=0)
    if dist.get_rank() != 0:
        return None
    merged = {}
    for part in parts:
        merged.update(part)
    return [merged[w] for w in range(n_windows)]


def fd_gradient(objective: Callable[[np.ndarray], float], theta: np.ndarray, eps: float = 1.4901161193847656e-08,
                device=None) -> Tuple[float, np.ndarray]:
    """Objective value and 2-point forward-difference gradient (what scipy's L-BFGS-B does for the
    reference fit, src/model.py:306-312) with the P+1 evaluations dealt round-robin to the ranks; every
    rank returns the same (f, grad).  Evaluation j=0 is theta itself, j>0 perturbs parameter j-1."""
    theta = np.asarray(theta, dtype=np.float64)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    vals = torch.zeros(theta.size + 1, dtype=F64, device=device)
    for j in range(rank, theta.size + 1, world):
        t = theta.copy()
        if j > 0:
            t[j - 1] += eps
        vals[j] = float(objective(t))
    if world > 1:
        dist.all_reduce(vals)  # every slot has exactly one non-zero contributor: exact
    v = vals.cpu().numpy()
    return float(v[0]), (v[1:] - v[0]) / eps


# ------------------------------------------------------------------------------------------------ K2
class VarioShard:
    """Row-block partition of the K2 pair tiles.  `tile_rows(...)` gives this rank's share of
    the tile rows of A, balanced by pair count (for a marginal variogram only tiles above the diagonal do
    work, so the split is triangular); the combine_* callables plug into ops.vario_extrema / vario_bin."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    @staticmethod
    def split(ta: int, tb: int, va: int, vb: int, same_field: bool, world: int) -> List[int]:
        """Boundaries b[0..world] over the `ta` tile rows with ~equal numbers of active tiles."""
        if same_field:  # tile (i, j) is active unless b0 + vb - 1 <= a0  <=>  (j + 1) * vb - 1 > i * va
            w = np.array([tb - min(tb, (i * va + 1) // vb) for i in range(ta)], dtype=np.float64)
        else:
            w = np.full(ta, float(tb))
        c = np.concatenate([[0.0], np.cumsum(w)])
        bounds = [int(np.searchsorted(c, c[-1] * r / world, side="left")) for r in range(world + 1)]
        bounds[0], bounds[-1] = 0, ta
        for r in range(1, world + 1):
            bounds[r] = max(bounds[r], bounds[r - 1])
        return bounds

    def tile_rows(self, ta: int, tb: int, va: int, vb: int, same_field: bool) -> Tuple[int, int]:
        b = self.split(ta, tb, va, vb, same_field, self.world)
        return b[self.rank], b[self.rank + 1]

    def combine_extrema(self, mn: float, mx: float, cnt: float, device=None):
        if self.world == 1:
            return mn, mx, cnt
        t = torch.tensor([mn, -mx], dtype=F64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        c = torch.tensor([cnt], dtype=F64, device=device)
        dist.all_reduce(c, group=self.group)
        return float(t[0]), float(-t[1]), float(c[0])

    def combine_partials(self, psum: torch.Tensor, pcnt: torch.Tensor) -> None:
        if self.world == 1:
            return
        dist.all_reduce(psum, group=self.group)  # one non-zero contributor per entry -> exact
        dist.all_reduce(pcnt, group=self.group)

    def gather_pairs(self, pairs: Optional[np.ndarray]) -> Optional[np.ndarray]:
        """Union of the guard-band index pairs found by every rank (identical on all ranks, rank order)."""
        if pairs is None or self.world == 1:
            return pairs
        parts = [None] * self.world
        dist.all_gather_object(parts, np.asarray(pairs, dtype=np.int64).reshape(-1, 2), group=self.group)
        return np.concatenate(parts, axis=0)


# ------------------------------------------------------------------------------------------------ C5
class CudaKernels:
    """The product kernel set: every method is one or a few C-ABI calls on CUDA tensors."""

    def __init__(self, device=None):
        from . import ops
        from ._lib import check, lib
        self.ops, self.lib, self.check = ops, lib, check
        self.device = ops.require_cuda() if device is None else torch.device(device)
        lo, hi = torch.cuda.Stream.priority_range()
        self.main = torch.cuda.current_stream(self.device)
        self.panel = torch.cuda.Stream(self.device, priority=hi)

    # memory / ordering ------------------------------------------------------------------------
    def empty(self, *shape, dtype=F64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=F64):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def to_device(self, a):
        return self.ops.to_device(a)

    def stream(self, which: str):
        return torch.cuda.stream(self.panel if which == "panel" else self.main)

    def event(self):
        e = torch.cuda.Event()
        e.record()
        return e

    def wait(self, event) -> None:
        if event is not None:
            torch.cuda.current_stream(self.device).wait_event(event)

    def sync(self) -> None:
        torch.cuda.synchronize(self.device)

    # compute -----------------------------------------------------------------------------------
    def pack_size(self, tb: int) -> int:
        return tb * tb + int(self.lib.ck_potrf_workspace_bytes(tb)) // 8

    def assemble(self, coords, targets, z, params, n_procs, i_pred, metric, tb, grid: ProcessGrid, local):
        o = self.ops
        _, pp = o._params(params, n_procs)
        c1 = coords[1] if n_procs == 2 else None
        n1 = coords[1].shape[0] if n_procs == 2 else 0
        self.check(self.lib.ck_mg_assemble(o._ptr(coords[0]), coords[0].shape[0], o._ptr(c1), n1, o._ptr(targets),
                                           targets.shape[0], o._ptr(z), pp, n_procs, i_pred, metric, tb, grid.P, grid.p,
                                           grid.Q, grid.q, o._ptr(local), local.stride(0), o._stream()), "ck_mg_assemble")

    def potrf_tile(self, tile: torch.Tensor, pack: torch.Tensor, info: torch.Tensor) -> None:
        """In-place Cholesky of the (tb x tb) view `tile`; pack <- [L tile (dense tb x tb) | inverted diagonal blocks]."""
        o, tb = self.ops, tile.shape[0]
        self.check(self.lib.ck_potrf(o._ptr(tile), tb, tile.stride(0), o._ptr(pack[tb * tb:]), o._ptr(info), o._stream()),
                   "ck_potrf")
        pack[: tb * tb].view(tb, tb).copy_(tile)

    # INT8 tensor-core path (FP64-equivalent products, csrc/ck_ozaki.cu) ---------------------------
    def _int8_ok(self, m: int, n: int, k: int) -> bool:
        """Worth a persistent tcgen05 launch: path enabled, K in range, enough 128 x 64 tiles to fill the machine."""
        min_tiles = int(os.environ.get("CK_MG_INT8_MIN_TILES", "600"))
        return bool(self.lib.ck_oz_active(1 << 20)) and k % 32 == 0 and 256 <= k <= 1024 and (m // 128) * (n // 64) >= min_tiles

    def _oz_scratch(self, rows_a: int, rows_b: int, k: int):
        """Slice / scale buffers, one set per stream (main / panel), grown on demand and reused."""
        need = (int(self.lib.ck_oz_slices_bytes(rows_a, k, 0)), int(self.lib.ck_oz_slices_bytes(rows_b, k, 1)),
                int(self.lib.ck_oz_scales_len(rows_a)), int(self.lib.ck_oz_scales_len(rows_b)))
        if not hasattr(self, "_oz"):
            self._oz = {}
        key = torch.cuda.current_stream(self.device).cuda_stream
        have = self._oz.get(key)
        if have is None or any(h.numel() < n for h, n in zip(have, need)):
            self._oz[key] = have = (self.empty(need[0], dtype=torch.uint8), self.empty(need[1], dtype=torch.uint8),
                                    self.empty(need[2]), self.empty(need[3]))
        return have

    def _oz_product(self, A, B, C, tb=0, gi0=0, gis=1, gj0=0, gjs=1) -> None:
        """C -= A B^T through digit slices; tb > 0: block-cyclic mask of ck_mg_update."""
        o = self.ops
        m, n, k = A.shape[0], B.shape[0], A.shape[1]
        fa, fb, sa, sb = self._oz_scratch(m, n, k)
        st = o._stream()
        # optional split of the SMs between the panel stream and the trailing update (CK_MG_PANEL_SMS = R > 0): the
        # persistent kernel of the main stream leaves R SMs to the look-ahead work of the next tile column
        # (a launch argument of the update: no process-wide state)
        r = int(os.environ.get("CK_MG_PANEL_SMS", str(getattr(self, "panel_sms", 0))))
        cap = 0
        if r > 0:
            nsm = torch.cuda.get_device_properties(self.device).multi_processor_count
            on_panel = torch.cuda.current_stream(self.device) == self.panel
            cap = r if on_panel else max(nsm - r, 1)
        self.check(self.lib.ck_oz_split(o._ptr(A), A.stride(0), m, k, o._ptr(fa), None, o._ptr(sa), st), "ck_oz_split")
        self.check(self.lib.ck_oz_split(o._ptr(B), B.stride(0), n, k, None, o._ptr(fb), o._ptr(sb), st), "ck_oz_split")
        if tb:
            self.check(self.lib.ck_oz_mg_update(o._ptr(fa), o._ptr(sa), m, o._ptr(fb), o._ptr(sb), n, k, o._ptr(C), C.stride(0),
                                                tb, gi0, gis, gj0, gjs, cap, st), "ck_oz_mg_update")
        else:
            self.check(self.lib.ck_oz_gemm(o._ptr(fa), o._ptr(sa), m, o._ptr(fb), o._ptr(sb), n, k, o._ptr(C), C.stride(0), 0, cap, st),
                       "ck_oz_gemm")

    def trsm(self, pack: torch.Tensor, tb: int, rows: torch.Tensor, out: torch.Tensor) -> None:
        """rows <- rows L^-T for the (nrows x tb) view `rows` (row stride = its leading dimension); `out` (nrows x tb,
        contiguous) receives a copy of the result (the panel staging buffer)."""
        o = self.ops
        nrows = rows.shape[0]
        if self._int8_ok(nrows, tb, tb):
            # one INT8 product with the explicitly inverted tile: X^T = solve(L, I) (target-major), rows L^-T = rows (X)^T
            xt = torch.eye(tb, dtype=F64, device=self.device)
            self.check(self.lib.ck_trsm_lower(o._ptr(pack), tb, tb, o._ptr(pack[tb * tb:]), o._ptr(xt), tb, tb, o._stream()),
                       "ck_trsm_lower")
            negx = (-xt).t().contiguous()      # row j = -(row j of L^-1)
            out.zero_()
            self._oz_product(rows, negx, out)  # out = 0 - rows (-X)^T
            rows.copy_(out)
            return
        self.check(self.lib.ck_trsm_lower(o._ptr(pack), tb, tb, o._ptr(pack[tb * tb:]), o._ptr(rows), nrows,
                                          rows.stride(0), o._stream()), "ck_trsm_lower")
        out.copy_(rows)

    def update(self, A, B, C, tb, gi0, gis, gj0, gjs) -> None:
        o = self.ops
        if self._int8_ok(C.shape[0], C.shape[1], A.shape[1]):
            self._oz_product(A, B, C, tb, gi0, gis, gj0, gjs)
            return
        self.check(self.lib.ck_mg_update(o._ptr(A), A.stride(0), o._ptr(B), B.stride(0), o._ptr(C), C.stride(0), C.shape[0],
                                         C.shape[1], A.shape[1], tb, gi0, gis, gj0, gjs, o._stream()), "ck_mg_update")

    def row_dots(self, V: torch.Tensor, y: torch.Tensor):
        o = self.ops
        vy, vv = self.empty(V.shape[0]), self.empty(V.shape[0])
        self.check(self.lib.ck_row_dots(o._ptr(V), V.stride(0), V.shape[0], V.shape[1], o._ptr(y), 0, o._ptr(vy), o._ptr(vv),
                                        o._stream()), "ck_row_dots")
        return vy, vv


class BlockCyclicCokriging:
    """Joint simple cokriging (src/joint_prediction.py:50-78) of one large system on a P x Q grid of GPUs.

    Layout: see csrc/ck_mg.cu.  Per tile column k (right-looking):
        diagonal owner      potrf(tile k,k)                                  -> broadcast down process column k%Q
        process column k%Q  panel rows I>k:  A_Ik <- A_Ik L_kk^-T (TRSM)     -> broadcast along each process ROW (a rank gets
                                                                                only the tiles I = p mod P: its A operand)
        every rank          panel tiles J = q mod Q (its B operand)          <- all-gather inside its process COLUMN of the
                                                                                tiles each member holds after the row broadcast
        every rank          A_IJ -= L_Ik L_Jk^T on its tiles I>k, J>k, J<=I  (one launch: ck_oz_mg_update / ck_mg_update)
    with one-panel look-ahead: the owners of column k+1 update that column first, factor it and start its
    broadcasts on a high-priority stream while everybody's trailing update of step k is still running.
    The target rows (C^T and z) ride along as extra row tiles, so after the sweep they hold L^-1 c and
    L^-1 z:  pred = (L^-1 c).(L^-1 z),  var = c0 - |L^-1 c|^2, reduced over the process row in rank order.
    """

    def __init__(self, grid: ProcessGrid, tile: int = 1024, kernels=None, lookahead: bool = True):
        if tile % 128:
            raise ValueError("tile must be a multiple of 128")
        self.g, self.tb = grid, int(tile)
        self.k = kernels if kernels is not None else CudaKernels()
        self.lookahead = lookahead
        # with look-ahead the persistent INT8 update kernel of the main stream leaves this many SMs to the panel stream
        # (measured on C5, 8 GPUs: 1.83 s without the split, 1.60 s with 40 SMs; CK_MG_PANEL_SMS overrides)
        if isinstance(self.k, CudaKernels):
            self.k.panel_sms = 40 if lookahead else 0
        # adaptive split (default on): per tile column the SM share of the panel stream is chosen so that the trailing update
        # of step k and the panel chain of step k + 1, which run side by side, take the same time (CK_MG_PANEL_SMS pins it)
        self.adaptive_split = lookahead and "CK_MG_PANEL_SMS" not in os.environ
        self.timings = {}
        self._tracing = bool(int(os.environ.get("CK_MG_TRACE", "0"))) and isinstance(self.k, CudaKernels)
        self._trace_ev = []

    # -- layout ---------------------------------------------------------------------------------
    def _layout(self, N: int, m: int):
        tb, g = self.tb, self.g
        self.N, self.m = N, m
        self.TC = (N + tb - 1) // tb
        self.cap = tb - 1
        self.TE = (m + self.cap - 1) // self.cap
        self.TR = self.TC + self.TE
        self.LRt = local_tiles(self.TR, g.P, g.p)
        self.LCt = local_tiles(self.TC, g.Q, g.q)
        self.LRmax = local_tiles(self.TR, g.P, 0)

    def local_bytes(self, N: int, m: int) -> int:
        self._layout(N, m)
        tb = self.tb
        return 8 * (self.LRt * tb * self.LCt * tb + 2 * self.g.P * self.LRmax * tb * tb + self.LCt * tb * tb)

    def _panel_share(self, k: int, nsm: int = 148) -> int:
        """SMs left to the panel stream while trailing update k and panel chain k + 1 run side by side.  Model (fitted to the
        C3 trace on 2 GPUs, profiles/r02_mg_trace_*): a tile product (tb^3 MACs through the INT8 kernel) costs `tau` SM-ms,
        the trailing update has ~ rows x cols / 2 of them per rank, the panel chain two per local row tile (column update +
        TRSM product) plus a fixed latency-bound part (tile Cholesky, its inverse, the broadcasts)."""
        g = self.g
        rows = max(self.TR - k - 1, 0)
        cols = max(self.TC - k - 1, 0)
        w_main = rows * cols / (2.0 * g.P * g.Q)
        w_panel = 2.0 * max(self.TR - k - 2, 0) / g.P
        tau, fixed = 3.65 * (self.tb / 1024.0) ** 3, 1.7
        # floor 24 (CK_MG_PANEL_MIN): measured on C3 / 2 GPUs with the dynamically scheduled update kernel 16 -> 254.9 ms,
        # 24 -> 247.4, 32 -> 248.2, 40 -> 253.7 (profiles/r02z_mg_panel_min_sweep_2gpu.log); with the static tile partition of
        # round 1 anything below 40 lost time (an update kernel whose CTAs are not all resident at launch finished late)
        floor = int(os.environ.get("CK_MG_PANEL_MIN", "24"))
        best, best_t = floor, float("inf")
        # the NCCL kernels of the look-ahead broadcasts occupy SMs of their own, hence a floor
        for r in range(floor, nsm - 23, 4):
            t = max(w_main * tau / (nsm - r), w_panel * tau / r + fixed)
            if t < best_t:
                best, best_t = r, t
        return best

    # -- the sweep --------------------------------------------------------------------------------
    def solve(self, coords: Sequence, z, targets, params, n_procs: int, i_pred: int, metric: int):
        """coords / z / targets: host arrays replicated on every rank.  Returns (pred, var, info) as numpy
        arrays / int, identical on every rank; info > 0 = order of the first non-PD leading minor."""
        K = self.k
        coords_d = [K.to_device(np.ascontiguousarray(np.asarray(c, dtype=np.float64))) for c in coords[:n_procs]]
        z_d = K.to_device(np.ascontiguousarray(np.hstack([np.asarray(v, dtype=np.float64) for v in z[:n_procs]])))
        t_d = K.to_device(np.ascontiguousarray(np.asarray(targets, dtype=np.float64)))
        pred_d, var_d, info = self.solve_device(coords_d, z_d, t_d, params, n_procs, i_pred, metric)
        K.sync()
        pred, var = pred_d.cpu().numpy(), var_d.cpu().numpy()
        inf = info.cpu().numpy()
        TC, tb = self.TC, self.tb
        bad = np.nonzero(inf[:TC] > 0)[0]
        first_bad = int(bad[0] * tb + inf[bad[0]]) if bad.size else 0
        if first_bad > self.N:
            first_bad = 0  # cannot happen: the pad is the identity
        self.timings = self._elapsed(self._ev)
        return pred, var, first_bad

    def solve_device(self, coords_d: Sequence, z_d, t_d, params, n_procs: int, i_pred: int, metric: int):
        """The sweep on device-resident inputs (replicated on every rank): returns (pred, var, info) as device tensors
        (m,), (m,), (tile columns,) without synchronising the host -- info[k] > 0 flags tile column k."""
        K, g, tb = self.k, self.g, self.tb
        params = np.asarray(params, dtype=np.float64)
        N, m = int(z_d.shape[0]), int(t_d.shape[0])
        self._layout(N, m)
        P, Q, p, q = g.P, g.Q, g.p, g.q
        TC, LRt, LCt = self.TC, self.LRt, self.LCt
        sig = params[i_pred] if n_procs == 2 else params[0]
        nug = params[8 + i_pred] if n_procs == 2 else params[3]
        c0 = sig * sig + nug

        local = K.empty(max(LRt * tb, 1), max(LCt * tb, 1))
        stage = K.empty(2, max(self.LRmax, 1), tb, tb)             # double-buffered: panel tiles of my process row (A operand)
        bcols = K.empty(2, max(LCt, 1), tb, tb)                     # double-buffered: panel tiles of my tile columns (B operand)
        cmax = max(LCt, 1)  # one process row may hold ALL of my column tiles (J = q mod Q has a fixed residue mod P when P | Q)
        self._xsend = K.empty(cmax, tb, tb) if P > 1 else None      # column exchange: my contribution / everybody's
        self._xrecv = K.empty(P, cmax, tb, tb) if P > 1 else None
        packs = [K.empty(K.pack_size(tb)) for _ in range(2)]
        info = K.zeros(max(TC, 1), dtype=torch.int32)
        self._build_exchange_plans()
        ev = {"t0": self._mark()}
        self._trace_ev = []
        K.assemble(coords_d, t_d, z_d, params, n_procs, i_pred, metric, tb, g, local)
        ev["t1"] = self._mark()
        with K.stream("main"):
            assembled = K.event()  # the panel stream must not touch `local` before the assembly kernels are done

        main_done = [None, None]  # main_done[b]: last trailing update that read stage[b] / bcols[b] has finished
        panel_ready = self._factor_panel(0, local, stage[0], bcols[0], packs[0], info, None, assembled) if TC else None
        for k in range(TC):
            buf = k % 2
            nxt = None
            if self.adaptive_split and isinstance(K, CudaKernels):
                K.panel_sms = self._panel_share(k)
            if self.lookahead and k + 1 < TC:
                # column k+1 first (its owners), then its panel, all on the panel stream
                nxt = self._factor_panel(k + 1, local, stage[1 - buf], bcols[1 - buf], packs[1 - buf], info,
                                         (k, stage[buf], bcols[buf], panel_ready), main_done[1 - buf])
            with K.stream("main"):
                K.wait(panel_ready)
                skip = (k + 1) if (self.lookahead and k + 1 < TC) else None
                self._trace(k, "main_begin")
                self._trailing_update(k, local, stage[buf], bcols[buf], skip_col=skip)
                self._trace(k, "main_end")
                main_done[buf] = K.event()
            if not self.lookahead and k + 1 < TC:
                nxt = self._factor_panel(k + 1, local, stage[1 - buf], bcols[1 - buf], packs[1 - buf], info, None, main_done[buf])
            panel_ready = nxt
        ev["t2"] = self._mark()

        # -- predictions from the target tiles
        part = K.zeros(2, max(m, 1))
        with K.stream("main"):
            for li in range(LRt):
                I = li * P + p
                if I < TC or LCt == 0:
                    continue
                t_lo = (I - TC) * self.cap
                nt = min(self.cap, m - t_lo)
                rows = local[li * tb: li * tb + nt, : LCt * tb]
                y = local[li * tb + tb - 1, : LCt * tb]
                vy, vv = K.row_dots(rows, y)
                part[0, t_lo: t_lo + nt] = vy
                part[1, t_lo: t_lo + nt] = vv
            if g.world > 1:
                allp = K.empty(g.world, 2, max(m, 1))
                dist.all_gather_into_tensor(allp.view(-1), part.view(-1))
                dist.all_reduce(info, op=dist.ReduceOp.MAX)
                total = allp[0].clone()
                for r in range(1, g.world):  # fixed rank order
                    total += allp[r]
            else:
                total = part
            pred = total[0, :m]
            var = c0 - total[1, :m]
        ev["t3"] = self._mark()
        self._ev = ev
        self._keep = (local, stage)  # factor stays resident (logdet / diagnostics)
        return pred, var, info

    def logdet(self) -> float:
        """2 sum log L_kk over the diagonal tiles (after solve()), summed over ranks."""
        local, _ = self._keep
        g, tb = self.g, self.tb
        s = self.k.zeros(1)
        for k in range(self.TC):
            if k % g.P == g.p and k % g.Q == g.q:
                li, lj = k // g.P, k // g.Q
                d = torch.diagonal(local[li * tb:(li + 1) * tb, lj * tb:(lj + 1) * tb])
                s += 2.0 * torch.log(d).sum()
        if g.world > 1:
            dist.all_reduce(s)
        return float(s.item())

    def gather_factor_rows(self, rows: Sequence[int]) -> np.ndarray:
        """Rows `rows` (stacked data indices) of the lower Cholesky factor L, assembled from the ranks that
        hold their tiles (after solve()); zeros right of the diagonal.  Every rank gets the same array.
        For sampled checks of L L^T against Sigma at sizes where nothing else can hold the matrix."""
        local, _ = self._keep
        g, tb = self.g, self.tb
        rows = [int(r) for r in rows]
        out = self.k.zeros(len(rows), self.TC * tb)
        for a, r in enumerate(rows):
            I, off = divmod(r, tb)
            if I % g.P != g.p:
                continue
            li = I // g.P
            for lj in range(self.LCt):
                J = lj * g.Q + g.q
                if J > I:
                    break
                out[a, J * tb:(J + 1) * tb] = local[li * tb + off, lj * tb:(lj + 1) * tb]
            if I % g.Q == g.q:  # diagonal tile: only columns <= off are part of L
                out[a, I * tb + off + 1:(I + 1) * tb] = 0.0
        if g.world > 1:
            dist.all_reduce(out)
        return out[:, : self.N].cpu().numpy()

    # -- pieces -----------------------------------------------------------------------------------
    def _factor_panel(self, k: int, local, stage_b, bcols_b, pack, info, pending, stage_free):
        """Panel stream: [apply `pending` = (k-1, its stage, its column tiles, its ready event) to column k on its owners]
        -> potrf(k,k) -> broadcast down the process column -> TRSM of the rows below -> row broadcast -> column exchange.
        Returns the event after which stage_b / bcols_b hold this rank's share of panel k."""
        K, g, tb = self.k, self.g, self.tb
        P, Q, p, q = g.P, g.Q, g.p, g.q
        qk, pk = k % Q, k % P
        ljk = k // Q
        with K.stream("panel"):
            K.wait(stage_free)
            if pending is not None:
                kp, stage_prev, bcols_prev, ready_prev = pending
                K.wait(ready_prev)
                self._trace(k, "panel_begin")
                if q == qk:  # bring column k up to date with panel k-1 (rows I >= k)
                    li0 = first_local_after(k - 1, P, p)
                    if li0 < self.LRt:
                        A = stage_prev[li0: self.LRt].view(-1, tb)
                        B = bcols_prev[ljk]  # tile k of panel k-1: column k is one of my tile columns
                        C = local[li0 * tb: self.LRt * tb, ljk * tb:(ljk + 1) * tb]
                        K.update(A, B, C, tb, li0 * P + p, P, k, Q)
            l0 = first_local_after(k, P, p)
            self._trace(k, "pending_done")
            if q == qk:
                if p == pk:
                    lik = k // P
                    K.potrf_tile(local[lik * tb:(lik + 1) * tb, ljk * tb:(ljk + 1) * tb], pack, info[k: k + 1])
                if P > 1:
                    dist.broadcast(pack, src=g.rank_of(pk, qk), group=g.col_groups[qk])
                self._trace(k, "potrf_done")
                if l0 < self.LRt:
                    rows = local[l0 * tb: self.LRt * tb, ljk * tb:(ljk + 1) * tb]
                    K.trsm(pack, tb, rows, stage_b[l0: self.LRt].view(-1, tb))
                self._trace(k, "trsm_done")
            # A operand: the panel tiles of my process row, from the member of my row that owns tile column k
            if Q > 1 and l0 < self.LRt:
                dist.broadcast(stage_b[l0: self.LRt], src=g.rank_of(p, qk), group=g.row_groups[p])
            self._trace(k, "row_bcast_done")
            # B operand: the panel tiles J of my tile columns (J > k); tile J sits with process row J mod P after the row
            # broadcast, so the members of my process column all-gather their shares
            lj0 = first_local_after(k, Q, q)
            if lj0 < self.LCt:
                if P == 1:
                    bcols_b[lj0: self.LCt].copy_(stage_b[lj0 * Q + q: (self.LCt - 1) * Q + q + 1: Q])
                else:
                    plan = self._exchange_plan(k)  # index tensors were uploaded before the sweep: no host sync here
                    n_js, cnt = plan["n"], plan["cnt"]
                    send = self._xsend[:cnt]                           # my share, padded to the largest share
                    if plan["mine"] is not None:
                        torch.index_select(stage_b, 0, plan["mine"], out=send[: plan["mine"].numel()])
                    if plan["holder"] is not None:
                        # P divides Q (e.g. 2 x 4): every tile of my columns sits with ONE process row -> a broadcast from it
                        if p == plan["holder"]:
                            bcols_b[lj0: self.LCt].copy_(send[:n_js])
                        dist.broadcast(bcols_b[lj0: self.LCt], src=g.rank_of(plan["holder"], q), group=g.col_groups[q])
                    else:
                        recv = self._xrecv.view(-1, tb, tb)[: P * cnt]  # [member of my process column][tile]
                        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=g.col_groups[q])
                        torch.index_select(recv, 0, plan["sel"], out=bcols_b[lj0: self.LCt])
            self._trace(k, "panel_ready")
            return K.event()

    def _build_exchange_plans(self) -> None:
        """Per tile column k: which panel tiles of my tile columns each member of my process column holds after the row
        broadcast, as index tensors ON THE DEVICE.  Built (and uploaded) once per layout, before the sweep -- creating a
        device tensor from a Python list inside the sweep would block the host on the panel stream and serialise the
        look-ahead (measured on 8 GPUs: profiles/r02_mg_trace_c3_8gpu_serialized_rank0.json)."""
        g = self.g
        P, Q, p, q = g.P, g.Q, g.p, g.q
        key = (self.TC, self.TR, P, Q, p, q)
        if getattr(self, "_plans_key", None) == key or P == 1:
            return
        dev = self.k.device
        plans = []
        for k in range(self.TC):
            lj0 = first_local_after(k, Q, q)
            Js = [lj * Q + q for lj in range(lj0, self.LCt)]
            if not Js:
                plans.append(None)
                continue
            share = [[J for J in Js if J % P == pp] for pp in range(P)]
            cnt = max(len(sh) for sh in share)
            holders = [pp for pp in range(P) if share[pp]]
            where = {J: pp * cnt + i for pp in range(P) for i, J in enumerate(share[pp])}
            plans.append({"n": len(Js), "cnt": cnt, "holder": holders[0] if len(holders) == 1 else None,
                          "mine": [J // P for J in share[p]] or None, "sel": [where[J] for J in Js]})
        # one upload for everything
        flat, spans = [], []
        for pl in plans:
            if pl is None:
                continue
            for name in ("mine", "sel"):
                if pl[name] is not None:
                    spans.append((pl, name, len(flat), len(pl[name])))
                    flat.extend(pl[name])
        buf = torch.tensor(flat if flat else [0], dtype=torch.int64).to(dev)
        for pl, name, off, n in spans:
            pl[name] = buf[off: off + n]
        self._plans, self._plans_key = plans, key

    def _exchange_plan(self, k: int) -> dict:
        return self._plans[k]

    def _trailing_update(self, k: int, local, stage_b, bcols_b, skip_col=None) -> None:
        """A_IJ -= L_Ik L_Jk^T on my tiles with I > k, J > k (J <= I), minus column `skip_col` (done by look-ahead)."""
        K, g, tb = self.k, self.g, self.tb
        P, Q, p, q = g.P, g.Q, g.p, g.q
        li0 = first_local_after(k, P, p)
        lj0 = first_local_after(k, Q, q)
        if skip_col is not None and lj0 < self.LCt and lj0 * Q + q == skip_col:
            lj0 += 1
        if li0 >= self.LRt or lj0 >= self.LCt:
            return
        A = stage_b[li0: self.LRt].view(-1, tb)
        B = bcols_b[lj0: self.LCt].view(-1, tb)
        C = local[li0 * tb: self.LRt * tb, lj0 * tb: self.LCt * tb]
        K.update(A, B, C, tb, li0 * P + p, P, lj0 * Q + q, Q)

    # -- timing -----------------------------------------------------------------------------------
    def _trace(self, k: int, what: str) -> None:
        """CK_MG_TRACE=1: one CUDA event per (tile column, point of the schedule) on the current stream; `trace_table()`
        turns them into milliseconds since the start of the sweep (diagnosis of the panel / update overlap)."""
        if self._tracing:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.k.device))
            self._trace_ev.append((k, what, e))

    def trace_table(self):
        torch.cuda.synchronize(self.k.device)
        t0 = self._ev["t1"]
        return [(k, what, t0.elapsed_time(e)) for k, what, e in self._trace_ev]

    def _mark(self):
        if isinstance(self.k, CudaKernels):
            e = torch.cuda.Event(enable_timing=True)
            e.record(self.k.main)
            return e
        import time
        return time.perf_counter()

    def _elapsed(self, ev) -> dict:
        names = [("assemble", "t0", "t1"), ("factor_solve", "t1", "t2"), ("reduce", "t2", "t3")]
        out = {}
        for name, a, b in names:
            if isinstance(ev[a], float):
                out[name + "_ms"] = 1e3 * (ev[b] - ev[a])
            else:
                out[name + "_ms"] = ev[a].elapsed_time(ev[b])
        return out


# ------------------------------------------------------------------------------------------------ native handle API
class NativeBlockCyclic:
    """The same sweep through the C handle API of libcokrig_b200.so (csrc/ck_mgctx.cu: ck_mg_create / ck_mg_joint_cov /
    ck_mg_potrf / ck_mg_potrs_predict / ck_mg_destroy) -- schedule, streams, events and the NCCL row / column
    communicators all live behind the C ABI; this class only allocates the workspace and moves the 128-byte NCCL id
    between the ranks (torch.distributed, any backend).  It is what a C / C++ host program would call directly
    (INTEGRATION.md); `BlockCyclicCokriging` is the torch.distributed twin whose orchestration the gloo CPU tests cover."""

    def __init__(self, P: Optional[int] = None, Q: Optional[int] = None, tile: int = 1024, device=None):
        from . import ops
        from ._lib import check, lib
        self.ops, self.lib, self.check = ops, lib, check
        self.device = ops.require_cuda() if device is None else torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if P is None or Q is None:
            P, Q = grid_shape(self.world)
        uid = (ctypes.c_ubyte * 128)()
        if self.world > 1:
            if self.rank == 0:
                check(lib.ck_mg_unique_id(uid), "ck_mg_unique_id")
            box = [bytes(uid)]
            dist.broadcast_object_list(box, src=0)
            uid = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.ck_mg_create(ctypes.byref(self._h), self.world, self.rank, int(P), int(Q), int(tile), uid), "ck_mg_create")
        self.P, self.Q, self.tile = int(P), int(Q), int(tile)
        self._ws = None
        self.timings = {}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            self.lib.ck_mg_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def local_bytes(self, N: int, m: int) -> int:
        return int(self.lib.ck_mg_workspace_bytes(self._h, int(N), int(m)))

    def solve_device(self, coords_d: Sequence, z_d, t_d, params, n_procs: int, i_pred: int, metric: int):
        """Device-resident replicated inputs -> (pred, var, info) device tensors, nothing synchronised."""
        o, lib = self.ops, self.lib
        _, pp = o._params(params, n_procs)
        N, m = int(z_d.shape[0]), int(t_d.shape[0])
        need = self.local_bytes(N, m)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        c1 = coords_d[1] if n_procs == 2 else None
        n1 = int(coords_d[1].shape[0]) if n_procs == 2 else 0
        pred = torch.empty(max(m, 1), dtype=F64, device=self.device)
        var = torch.empty(max(m, 1), dtype=F64, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        st = o._stream()
        self.check(lib.ck_mg_joint_cov(self._h, o._ptr(coords_d[0]), int(coords_d[0].shape[0]), o._ptr(c1), n1, o._ptr(t_d), m,
                                       o._ptr(z_d), pp, n_procs, i_pred, metric, o._ptr(self._ws), self._ws.numel(), st),
                   "ck_mg_joint_cov")
        self.check(lib.ck_mg_potrf(self._h, st), "ck_mg_potrf")
        self.check(lib.ck_mg_potrs_predict(self._h, o._ptr(pred), o._ptr(var), o._ptr(info), st), "ck_mg_potrs_predict")
        return pred[:m], var[:m], info

    def solve(self, coords: Sequence, z, targets, params, n_procs: int, i_pred: int, metric: int):
        """Host arrays replicated on every rank -> (pred, var, info) as numpy arrays / int, identical on every rank."""
        o = self.ops
        coords_d = [o.to_device(np.ascontiguousarray(np.asarray(c, dtype=np.float64))) for c in coords[:n_procs]]
        z_d = o.to_device(np.ascontiguousarray(np.hstack([np.asarray(v, dtype=np.float64) for v in z[:n_procs]])))
        t_d = o.to_device(np.ascontiguousarray(np.asarray(targets, dtype=np.float64)))
        pred, var, info = self.solve_device(coords_d, z_d, t_d, params, n_procs, i_pred, metric)
        out = pred.cpu().numpy(), var.cpu().numpy(), int(info.item())
        t = (ctypes.c_double * 3)()
        self.check(self.lib.ck_mg_times_ms(self._h, t), "ck_mg_times_ms")
        self.timings = {"assemble_ms": t[0], "factor_solve_ms": t[1], "reduce_ms": t[2]}
        return out

    def logdet(self) -> float:
        out = torch.zeros(1, dtype=F64, device=self.device)
        self.check(self.lib.ck_mg_logdet(self._h, self.ops._ptr(out), self.ops._stream()), "ck_mg_logdet")
        return float(out.item())

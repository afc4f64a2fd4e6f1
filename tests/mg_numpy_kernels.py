"""Numpy stand-in for cokrig_b200.parallel.CudaKernels -- TEST INFRASTRUCTURE ONLY.

The product kernel set (CudaKernels) issues C-ABI calls on CUDA tensors and raises without a GPU.
This class has the same methods but computes each local step with numpy on CPU torch tensors, so the
world_size-2/4 `gloo` tests can exercise the block-cyclic layout, the look-ahead ordering, the
broadcast pattern and the rank-ordered reductions of BlockCyclicCokriging without a device.
Assembly goes through the CPU oracle (the checker), tile by tile, following the layout documented in
csrc/ck_mg.cu.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

import cokrig_oracle as orc

F64 = torch.float64


class NumpyKernels:
    def __init__(self):
        self.device = torch.device("cpu")

    # memory / ordering: a single in-order "stream"
    def empty(self, *shape, dtype=F64):
        return torch.full(shape, float("nan"), dtype=dtype) if dtype == F64 else torch.zeros(*shape, dtype=dtype)

    def zeros(self, *shape, dtype=F64):
        return torch.zeros(*shape, dtype=dtype)

    def to_device(self, a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))

    def stream(self, which: str):
        return contextlib.nullcontext()

    def event(self):
        return object()

    def wait(self, event) -> None:
        pass

    def sync(self) -> None:
        pass

    # compute
    def pack_size(self, tb: int) -> int:
        return tb * tb

    def assemble(self, coords, targets, z, params, n_procs, i_pred, metric, tb, grid, local):
        name = "haversine" if metric == 1 else "euclidean"
        P = orc.Params(list(params), n_procs)
        cs = [c.numpy() for c in coords]
        sigma = orc.joint_cov(P, cs, name)
        cdp = orc.pred_cross_cov(P, i_pred, cs, targets.numpy(), name)  # (N, m)
        N, m = sigma.shape[0], targets.shape[0]
        cap = tb - 1
        TC, TE = -(-N // tb), -(-m // cap)
        glob = np.zeros(((TC + TE) * tb, TC * tb))
        glob[:N, :N] = sigma
        for d in range(N, TC * tb):
            glob[d, d] = 1.0
        for e in range(TE):
            t_lo, t_hi = e * cap, min(m, (e + 1) * cap)
            r0 = (TC + e) * tb
            glob[r0: r0 + (t_hi - t_lo), :N] = cdp[:, t_lo:t_hi].T
            glob[r0 + tb - 1, :N] = z.numpy()
        loc = local.numpy()
        for I in range(grid.p, TC + TE, grid.P):
            for J in range(grid.q, TC, grid.Q):
                if I < TC and J > I:
                    continue  # never referenced: stays NaN so that any use shows up in the results
                li, lj = I // grid.P, J // grid.Q
                loc[li * tb:(li + 1) * tb, lj * tb:(lj + 1) * tb] = glob[I * tb:(I + 1) * tb, J * tb:(J + 1) * tb]

    def potrf_tile(self, tile, pack, info):
        tb = tile.shape[0]
        a = np.tril(tile.numpy())
        a = a + np.tril(a, -1).T
        try:
            L = np.linalg.cholesky(a)
        except np.linalg.LinAlgError:
            info[0] = 1
            L = np.eye(tb)
        t = tile.numpy()
        il = np.tril_indices(tb)
        t[il] = L[il]  # the upper triangle is left untouched, like ck_potrf
        pack[: tb * tb].view(tb, tb).copy_(torch.from_numpy(np.ascontiguousarray(L)))

    def trsm(self, pack, tb, rows, out):
        L = pack[: tb * tb].view(tb, tb).numpy()
        r = rows.numpy()
        r[...] = np.linalg.solve(np.tril(L), r.T).T
        out.copy_(rows)

    def update(self, A, B, C, tb, gi0, gis, gj0, gjs):
        a, b, c = A.numpy(), B.numpy(), C.numpy()
        full = a @ b.T
        for li in range(c.shape[0] // tb):
            for lj in range(c.shape[1] // tb):
                I, J = gi0 + li * gis, gj0 + lj * gjs
                if J > I:
                    continue
                blk = full[li * tb:(li + 1) * tb, lj * tb:(lj + 1) * tb]
                if J == I:
                    blk = np.tril(blk)
                c[li * tb:(li + 1) * tb, lj * tb:(lj + 1) * tb] -= blk

    def row_dots(self, V, y):
        v = V.numpy()
        return torch.from_numpy(v @ y.numpy()), torch.from_numpy(np.sum(v * v, axis=1))

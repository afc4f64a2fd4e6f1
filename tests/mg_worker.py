"""Per-rank bodies of the world_size > 1 CPU tests (gloo) -- TEST INFRASTRUCTURE ONLY.

`run_ranks(world, name, *args)` spawns `world` processes, initialises torch.distributed (gloo,
127.0.0.1) in each, runs WORKERS[name](rank, world, *args) and returns the per-rank results.
"""
from __future__ import annotations

import os
import socket
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "sif-xco2-cokriging_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, name, args, queue):
    try:
        torch.set_num_threads(1)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        out = WORKERS[name](rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        queue.put((rank, "ok", out))
    except Exception:  # noqa: BLE001
        queue.put((rank, "error", traceback.format_exc()))


def run_ranks(world: int, name: str, *args, timeout: float = 120.0):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, name, args, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, out = queue.get(timeout=timeout)
            if status != "ok":
                raise RuntimeError(f"rank {rank} failed:\n{out}")
            results[rank] = out
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    return [results[r] for r in range(world)]


# ------------------------------------------------------------------------------------------------ workers
def _block_cyclic(rank, world, P, Q, tile, n0, n1, m, params, n_procs, i_pred, metric, lookahead, seed):
    from cokrig_b200 import parallel
    from mg_numpy_kernels import NumpyKernels
    rng = np.random.default_rng(seed)
    if metric == 1:
        coords = [np.c_[rng.uniform(25, 50, n), rng.uniform(-120, -70, n)] for n in (n0, n1)][:n_procs]
        targets = np.c_[rng.uniform(25, 50, m), rng.uniform(-120, -70, m)]
    else:
        coords = [rng.uniform(0, 1, (n, 2)) for n in (n0, n1)][:n_procs]
        targets = rng.uniform(0, 1, (m, 2))
    if m > 2:
        targets[1] = coords[i_pred][3]  # a target on a datum: nugget in c (src/joint_prediction.py:118)
    z = [rng.standard_normal(len(c)) for c in coords]
    grid = parallel.ProcessGrid(P, Q)
    solver = parallel.BlockCyclicCokriging(grid, tile=tile, kernels=NumpyKernels(), lookahead=lookahead)
    pred, var, info = solver.solve(coords, z, targets, params, n_procs, i_pred, metric)
    n_all = sum(len(c) for c in coords)
    rows = [0, 1, min(130, n_all - 1), n_all // 2, n_all - 1]
    return {"pred": pred, "var": var, "info": info, "logdet": solver.logdet(), "coords": coords, "z": z,
            "rows": rows, "L_rows": solver.gather_factor_rows(rows),
            "targets": targets, "local_bytes": solver.local_bytes(sum(len(c) for c in coords), m)}


def _block_cyclic_not_pd(rank, world, P, Q, tile):
    from cokrig_b200 import parallel
    from mg_numpy_kernels import NumpyKernels
    rng = np.random.default_rng(5)
    coords = [rng.uniform(0, 1, (200, 2)), rng.uniform(0, 1, (150, 2))]
    z = [rng.standard_normal(200), rng.standard_normal(150)]
    params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .0, .0, -1.3]  # |rho| > 1: not a valid model
    solver = parallel.BlockCyclicCokriging(parallel.ProcessGrid(P, Q), tile=tile, kernels=NumpyKernels())
    _, _, info = solver.solve(coords, z, rng.uniform(0, 1, (10, 2)), params, 2, 0, 0)
    return info


def _fd_gradient(rank, world, theta):
    from cokrig_b200 import parallel
    calls = []

    def objective(t):
        calls.append(t.copy())
        return float(np.sum(np.sin(t)) + t[0] * t[-1])
    f, g = parallel.fd_gradient(objective, np.asarray(theta))
    return {"f": f, "g": g, "calls": len(calls)}


def _windows(rank, world, n_windows):
    from cokrig_b200 import parallel
    mine = parallel.shard_windows(n_windows, rank, world)
    local = {w: np.full(3, float(w)) for w in mine}
    merged = parallel.gather_window_results(local, n_windows)
    return {"mine": mine, "merged": None if merged is None else np.vstack(merged)}


def _vario_shard(rank, world, ta, tb, va, vb, same_field, n_bins, seed):
    """The cross-rank combine of K2: each rank owns a block of tile rows; partials of the other rows are zero."""
    from cokrig_b200 import parallel
    sh = parallel.VarioShard()
    lo, hi = sh.tile_rows(ta, tb, va, vb, same_field)
    rng = np.random.default_rng(seed)  # same stream on every rank = the "single GPU" partials
    psum_all = rng.standard_normal((ta, tb, n_bins))
    pcnt_all = rng.integers(0, 1 << 20, (ta, tb, n_bins)).astype(np.int32)
    psum = torch.zeros(ta, tb, n_bins, dtype=torch.float64)
    pcnt = torch.zeros(ta, tb, n_bins, dtype=torch.int32)
    psum[lo:hi] = torch.from_numpy(psum_all[lo:hi])
    pcnt[lo:hi] = torch.from_numpy(pcnt_all[lo:hi])
    sh.combine_partials(psum.view(-1), pcnt.view(-1))
    mn, mx, cnt = sh.combine_extrema(10.0 + rank, 100.0 - rank, float(hi - lo))
    pairs = sh.gather_pairs(np.array([[rank, rank + 1]], dtype=np.int64))
    return {"rows": (lo, hi), "sum_exact": bool(np.array_equal(psum.numpy(), psum_all)),
            "cnt_exact": bool(np.array_equal(pcnt.numpy(), pcnt_all)), "extrema": (mn, mx, cnt), "pairs": pairs}


WORKERS = {"block_cyclic": _block_cyclic, "block_cyclic_not_pd": _block_cyclic_not_pd, "fd_gradient": _fd_gradient,
           "windows": _windows, "vario_shard": _vario_shard}

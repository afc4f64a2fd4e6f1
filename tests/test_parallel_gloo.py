"""world_size > 1 on CPU (gloo): the host-side logic of cokrig_b200.parallel -- process grid, block-cyclic
layout, look-ahead order, broadcast pattern, rank-ordered reductions -- with the numpy kernel set of
tests/mg_numpy_kernels.py standing in for the CUDA kernels (the product path has no CPU fallback).
The checker is the CPU oracle's joint_predict (src/joint_prediction.py:50-78)."""
import numpy as np
import pytest

import cokrig_oracle as orc
from mg_worker import run_ranks

HALF = [1.0, 0.8, 1.5, 1.5, 1.5, 0.3, 0.3, 0.3, 0.02, 0.02, -0.2]
HALF_KM = [1.0, 0.8, 1.5, 1.5, 1.5, 500.0, 500.0, 500.0, 0.02, 0.02, -0.2]


def _check_against_oracle(res, params, n_procs, i_pred, metric):
    r0 = res[0]
    for r in res[1:]:  # identical on every rank
        assert np.array_equal(r["pred"], r0["pred"]) and np.array_equal(r["var"], r0["var"])
        assert r["info"] == r0["info"] and r["logdet"] == r0["logdet"]
    assert r0["info"] == 0
    P = orc.Params(params, n_procs)
    name = "haversine" if metric == 1 else "euclidean"
    pred, err, _ = orc.joint_predict(P, i_pred, r0["coords"], r0["z"], r0["targets"], name)
    # predictions of a zero-mean field cross zero while the solve error is normwise: relative to the field scale, and
    # pointwise relative where the prediction is not small (same criterion as tests/test_gpu_parallel.py)
    assert np.max(np.abs(r0["pred"] - pred)) / np.max(np.abs(pred)) < 1e-9
    big = np.abs(pred) > 0.1 * np.max(np.abs(pred))
    assert np.max(np.abs(r0["pred"][big] - pred[big]) / np.abs(pred[big])) < 1e-9
    assert np.max(np.abs(r0["var"] - err ** 2)) < 1e-9 * (params[i_pred if n_procs == 2 else 0] ** 2)
    sigma = orc.joint_cov(P, r0["coords"], name)
    L = np.linalg.cholesky(sigma)
    for r in res:  # rows of the factor gathered from their owners
        assert np.max(np.abs(r["L_rows"] - L[r["rows"]])) < 1e-12
    assert abs(r0["logdet"] - np.linalg.slogdet(sigma)[1]) < 1e-9 * abs(np.linalg.slogdet(sigma)[1]) + 1e-9


@pytest.mark.parametrize("world,P,Q,lookahead", [(2, 1, 2, True), (2, 2, 1, True), (4, 2, 2, True), (2, 1, 2, False)])
def test_block_cyclic_matches_oracle(world, P, Q, lookahead):
    # N = 700 with 128-tiles: 6 tile columns (the last one ragged); 300 targets -> 3 target tile rows (ragged)
    res = run_ranks(world, "block_cyclic", P, Q, 128, 380, 320, 300, HALF, 2, 1, 0, lookahead, 11)
    _check_against_oracle(res, HALF, 2, 1, 0)


def test_block_cyclic_many_tile_columns_when_p_divides_q():
    # 2 x 2 grid, 12 tile columns: every panel tile of a rank's columns sits with one process row (broadcast branch of the
    # column exchange); 4 x 1 below exercises the all-gather branch with unequal shares
    res = run_ranks(4, "block_cyclic", 2, 2, 128, 800, 700, 200, HALF, 2, 0, 0, True, 5)
    _check_against_oracle(res, HALF, 2, 0, 0)
    res = run_ranks(4, "block_cyclic", 4, 1, 128, 500, 400, 150, HALF, 2, 1, 0, True, 6)
    _check_against_oracle(res, HALF, 2, 1, 0)


def test_block_cyclic_haversine_univariate_single_tile_column():
    # N = 100 < tile: one tile column, all of it on process column 0; rank 1 holds target rows only
    uni = [1.3, 1.5, 400.0, 0.05]
    res = run_ranks(2, "block_cyclic", 2, 1, 128, 100, 0, 140, uni, 1, 0, 1, True, 3)
    _check_against_oracle(res, uni, 1, 0, 1)


def test_block_cyclic_more_ranks_than_tiles():
    # 1 x 2 grid but a single tile column: process column 1 owns no columns at all
    res = run_ranks(2, "block_cyclic", 1, 2, 256, 120, 90, 40, HALF_KM, 2, 0, 1, True, 7)
    _check_against_oracle(res, HALF_KM, 2, 0, 1)


def test_block_cyclic_reports_first_bad_minor_on_every_rank():
    infos = run_ranks(2, "block_cyclic_not_pd", 1, 2, 128)
    assert infos[0] == infos[1] and infos[0] > 0


def test_fd_gradient_round_robin():
    theta = np.linspace(0.2, 1.4, 11)
    res = run_ranks(2, "fd_gradient", theta)
    f = float(np.sum(np.sin(theta)) + theta[0] * theta[-1])
    g = np.cos(theta)
    g[0] += theta[-1]
    g[-1] += theta[0]
    for r in res:
        assert r["f"] == f
        assert np.allclose(r["g"], g, atol=1e-6)
    assert np.array_equal(res[0]["g"], res[1]["g"])
    assert res[0]["calls"] + res[1]["calls"] == theta.size + 1 and abs(res[0]["calls"] - res[1]["calls"]) <= 1


def test_windows_round_robin_and_gather():
    res = run_ranks(2, "windows", 7)
    assert res[0]["mine"] == [0, 2, 4, 6] and res[1]["mine"] == [1, 3, 5]
    assert res[1]["merged"] is None
    assert np.array_equal(res[0]["merged"][:, 0], np.arange(7.0))


@pytest.mark.parametrize("same_field", [True, False])
def test_vario_shard_combine_is_exact_and_rank_independent(same_field):
    res = run_ranks(2, "vario_shard", 9, 9, 64, 64, same_field, 5, 1)
    (lo0, hi0), (lo1, hi1) = res[0]["rows"], res[1]["rows"]
    assert lo0 == 0 and hi0 == lo1 and hi1 == 9
    if same_field:
        assert hi0 < 5  # triangular pair count: the first rank gets fewer, longer rows
    for r in res:
        assert r["sum_exact"] and r["cnt_exact"]
        assert r["extrema"] == (10.0, 100.0, 9.0)
        assert np.array_equal(r["pairs"], np.array([[0, 1], [1, 2]]))


def test_vario_split_balances_active_tiles():
    from cokrig_b200 import parallel
    for world in (1, 2, 4, 8):
        for same in (True, False):
            b = parallel.VarioShard.split(157, 157, 64, 64, same, world)
            assert b[0] == 0 and b[-1] == 157 and all(x <= y for x, y in zip(b, b[1:]))
            w = [157 - min(157, (i * 64 + 1) // 64) if same else 157 for i in range(157)]
            loads = [sum(w[b[r]: b[r + 1]]) for r in range(world)]
            assert max(loads) <= 1.15 * (sum(w) / world) + 157


def test_grid_shape_and_tile_ownership():
    from cokrig_b200 import parallel
    assert parallel.grid_shape(8) == (4, 2) and parallel.grid_shape(4) == (2, 2)
    assert parallel.grid_shape(2) == (2, 1) and parallel.grid_shape(1) == (1, 1)
    for ntiles in (0, 1, 5, 8, 13):
        for nprocs in (1, 2, 3, 4):
            owned = [parallel.local_tiles(ntiles, nprocs, r) for r in range(nprocs)]
            assert owned == [len(range(r, ntiles, nprocs)) for r in range(nprocs)]
            for r in range(nprocs):
                for k in range(-1, ntiles):
                    l0 = parallel.first_local_after(k, nprocs, r)
                    assert l0 * nprocs + r > k and (l0 == 0 or (l0 - 1) * nprocs + r <= k)

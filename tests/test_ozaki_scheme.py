"""CPU checks of the INT8 update scheme (numpy model of csrc/ck_ozaki.cu in tests/ozaki_model.py): the digit recoding
is exact, the sliced product is at least as accurate as a plain FP64 product, and a blocked Cholesky + cokriging
solve whose every update goes through the scheme reproduces the oracle's predictions like LAPACK does."""
import numpy as np

import cokrig_oracle as orc
import ozaki_model as oz


def test_digits_recode_the_fixed_point_value_exactly():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 64)) * np.exp(rng.uniform(-30, 30, (50, 1)))
    x[3] = 0.0
    digits, scales = oz.split(x)
    assert all(((d >= -128) & (d <= 127)).all() for d in digits)
    rec = sum(d * 2.0 ** (-8 * p) for p, d in enumerate(digits)) * scales[:, None]
    rowmax = np.maximum(np.abs(x).max(axis=1, keepdims=True), 1e-300)
    assert (np.abs(rec - x) / rowmax).max() <= 2.0 ** -56 / 0.49
    assert (rec[3] == 0).all() and scales[3] == 0


def test_sliced_product_is_fp64_accurate():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((96, 1024)), rng.standard_normal((80, 1024))
    ld = np.longdouble
    exact = a.astype(ld) @ b.T.astype(ld)
    rowmax = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :]
    err_oz = np.abs(oz.product(a, b).astype(ld) - exact).astype(float) / rowmax
    err_fp = np.abs((a @ b.T).astype(ld) - exact).astype(float) / rowmax
    assert err_oz.max() < 32 * np.sqrt(1024) * 2.0 ** -56
    assert err_oz.max() < err_fp.max()  # fixed point with exact accumulation beats sequential FP64 rounding


def test_cokriging_through_the_scheme_matches_the_oracle():
    params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .01, .01, -.6]
    P = orc.Params(params)
    grid = orc.expand_grid(xcount=24, ycount=24)
    _, _, z = orc.sim_fields(P, grid, seed=1)
    pc = np.random.default_rng(7).uniform(0, 1, (200, 2))
    sigma = orc.joint_cov(P, [grid, grid], "euclidean")
    cdp = orc.pred_cross_cov(P, 1, [grid, grid], pc, "euclidean")
    ref_pred, ref_err, _ = orc.joint_predict(P, 1, [grid, grid], z, pc, "euclidean")
    L = oz.cholesky_blocked(sigma, 256, oz.product)
    V = oz.solve_blocked(L, np.c_[cdp, np.hstack(z)].T, 256, oz.product)
    pred, var = V[:-1] @ V[-1], 1.01 - (V[:-1] ** 2).sum(axis=1)
    assert np.abs(pred - ref_pred).max() / np.abs(ref_pred).max() < 1e-11
    assert np.abs(var - ref_err ** 2).max() < 1e-11
    from scipy.linalg import cholesky
    assert np.abs(L - cholesky(sigma, lower=True)).max() < 1e-12


# ------------------------------------------------------------------------------------------------ tile order (host-only ABI)
def _tile_order(m, n, lower=0, tb=0, gi0=0, gis=1, gj0=0, gjs=1):
    import ctypes
    from cokrig_b200._lib import lib
    nvirt = ctypes.c_int64()
    count = lib.ck_oz_tile_order(m, n, lower, tb, gi0, gis, gj0, gjs, None, 0, ctypes.byref(nvirt), None)
    assert count >= 0
    ij = (ctypes.c_int * (2 * max(count, 1)))()
    lim = (ctypes.c_int64 * max(count, 1))()
    assert lib.ck_oz_tile_order(m, n, lower, tb, gi0, gis, gj0, gjs, ij, count, None, lim) == count
    return np.array(ij[: 2 * count]).reshape(-1, 2), np.array(lim[:count]), nvirt.value


def test_tile_order_covers_every_tile_exactly_once():
    """csrc/ck_ozaki.cu oz_decode (the order the dynamic scheduler of the INT8 update kernel hands tiles out):
    rectangular updates visit every 128 x 64 tile once; lower updates exactly the tiles that touch the lower triangle."""
    for (m, n) in ((8832, 32768), (1000, 900), (128, 64), (5000, 131)):
        ij, lim, nvirt = _tile_order(m, n)
        ni, nj = -(-m // 128), -(-n // 64)
        assert len(ij) == ni * nj == len({(i, j) for i, j in ij}) and nvirt >= len(ij)
        assert ij[:, 0].max() == ni - 1 and ij[:, 1].max() == nj - 1 and (lim == -1).all()
    for n in (38976, 16384, 3000, 1500, 128, 4223):
        ij, lim, nvirt = _tile_order(n, n, lower=1)
        ni, nj = -(-n // 128), -(-n // 64)
        want = {(i, j) for i in range(ni) for j in range(nj) if 64 * j <= 128 * i + 127}
        assert {(i, j) for i, j in ij} == want and len(ij) == len(want)
        assert (lim == 128 * ij[:, 0]).all()        # first row of the tile: columns <= row
        if n >= 16384:
            assert nvirt <= 1.2 * len(want)   # few skipped virtual tiles: the static fallback stays balanced too


def test_tile_order_block_cyclic_mask_matches_ck_mg_update():
    """Block-cyclic mode: local tile (li, lj) = global tile (gi0 + li gis, gj0 + lj gjs); tiles with J > I are skipped,
    J == I keeps its lower triangle (include/cokrig.h, ck_mg_update contract)."""
    tb = 512
    for (lrt, lct, gi0, gis, gj0, gjs) in ((5, 4, 1, 2, 0, 2), (3, 7, 2, 4, 1, 1), (4, 4, 0, 1, 0, 1), (2, 3, 7, 2, 8, 2)):
        ij, lim, _ = _tile_order(lrt * tb, lct * tb, 0, tb, gi0, gis, gj0, gjs)
        got = {(i, j) for i, j in ij}
        want = set()
        for I in range(lrt * tb // 128):
            for j in range(lct * tb // 64):
                li, lj = (I * 128) // tb, (j * 64) // tb
                Ig, Jg = gi0 + li * gis, gj0 + lj * gjs
                if Jg < Ig or (Jg == Ig and (j * 64 - lj * tb) <= (I * 128 + 127 - li * tb)):
                    want.add((I, j))
        assert got == want and len(ij) == len(want)
        for (I, j), l in zip(ij, lim):
            li, lj = (I * 128) // tb, (j * 64) // tb
            diag = gi0 + li * gis == gj0 + lj * gjs
            assert l == (lj * tb + (I * 128 - li * tb) if diag else -1)


def test_tile_order_keeps_the_tiles_in_flight_inside_one_l2_window():
    """What the dynamic scheduler relies on: ANY 148 consecutive tiles of the order (one per SM) touch at most two
    super-rows' worth of A blocks and a dozen B blocks -- a few tens of MB of slices, L2-resident."""
    ij, _, _ = _tile_order(38976, 38976, lower=1)
    full = ij[ij[:, 0] < 16 * (305 // 16)]     # the last, partial super-row (one row block) is a plain row sweep
    worst_a = worst_b = 0
    worst_mb = 0.0
    for s in range(0, len(ij) - 148, 37):
        w = ij[s: s + 148]
        worst_mb = max(worst_mb, (len(set(w[:, 0])) * 128 + len(set(w[:, 1])) * 64) * 1024 * 7 / 2 ** 20)
    for s in range(0, len(full) - 148, 37):
        w = full[s: s + 148]
        worst_a = max(worst_a, len(set(w[:, 0])))
        worst_b = max(worst_b, len(set(w[:, 1])))
    assert worst_a <= 32 and worst_b <= 48, (worst_a, worst_b)
    assert worst_mb < 80, worst_mb

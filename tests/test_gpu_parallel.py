"""GPU tests of the multi-GPU building blocks (cokrig_b200.parallel with the CUDA kernel set):
on ONE device the block-cyclic sweep degenerates to a 1 x 1 grid and still runs every mg kernel
(ck_mg_assemble, ck_mg_update, ck_row_dots, tile potrf / trsm); with >= 2 devices the same checks plus
the sharded variogram run under torchrun (tools/mg_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cokrig_oracle as orc
from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu

HALF = [1.0, 0.8, 1.5, 1.5, 1.5, 0.3, 0.3, 0.3, 0.02, 0.02, -0.2]
GENERIC_KM = [1.0, 0.8, 0.75, 1.0, 1.25, 500.0, 500.0, 500.0, 0.02, 0.02, -0.2]


def _inputs(metric, n0, n1, m, seed):
    rng = np.random.default_rng(seed)
    if metric == 1:
        coords = [np.c_[rng.uniform(25, 50, n), rng.uniform(-120, -70, n)] for n in (n0, n1)]
        targets = np.c_[rng.uniform(25, 50, m), rng.uniform(-120, -70, m)]
    else:
        coords = [rng.uniform(0, 1, (n, 2)) for n in (n0, n1)]
        targets = rng.uniform(0, 1, (m, 2))
    targets[1] = coords[0][3]
    return coords, [rng.standard_normal(n0), rng.standard_normal(n1)], targets


@pytest.mark.parametrize("metric,params,tile,lookahead", [(0, HALF, 128, True), (1, GENERIC_KM, 256, True), (0, HALF, 256, False)])
def test_block_cyclic_single_rank_vs_oracle(metric, params, tile, lookahead):
    from cokrig_b200 import parallel
    coords, z, targets = _inputs(metric, 700, 650, 420, 17)
    solver = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=tile, lookahead=lookahead)
    for i_pred in (0, 1):
        pred, var, info = solver.solve(coords, z, targets, params, 2, i_pred, metric)
        rp, re, _ = orc.joint_predict(orc.Params(params), i_pred, coords, z, targets, "haversine" if metric else "euclidean")
        assert info == 0
        # predictions cross zero on this random field (values down to 1e-3 of the field scale) while the solve
        # error is normwise (kappa(Sigma) ~ 1e4-1e5): compare relative to the field scale, and pointwise
        # relative where the prediction is not small
        assert np.max(np.abs(pred - rp)) / np.max(np.abs(rp)) < 1e-9
        big = np.abs(rp) > 0.1 * np.max(np.abs(rp))
        assert relerr(pred[big], rp[big]) < 1e-9
        assert np.max(np.abs(var - re ** 2)) < 1e-9
    sigma = orc.joint_cov(orc.Params(params), coords, "haversine" if metric else "euclidean")
    assert abs(solver.logdet() / np.linalg.slogdet(sigma)[1] - 1) < 1e-10


def test_block_cyclic_single_rank_int8_updates_vs_oracle():
    """The same sweep with the trailing updates on the INT8 tensor cores (ck_oz_split + ck_oz_mg_update)."""
    from cokrig_b200 import parallel
    from cokrig_b200._lib import lib
    coords, z, targets = _inputs(0, 1500, 1400, 600, 23)
    os.environ["CK_MG_INT8_MIN_TILES"] = "48"  # also the tile-column updates and the inverted-tile TRSM of this small case
    try:
        launches0 = lib.ck_launch_count()
        solver = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=256, lookahead=True)
        pred, var, info = solver.solve(coords, z, targets, HALF, 2, 1, 0)
        n_int8 = lib.ck_launch_count() - launches0
        lib.ck_oz_configure(0, -1)
        launches0 = lib.ck_launch_count()
        pred_d, var_d, _ = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=256).solve(coords, z, targets, HALF, 2, 1, 0)
        n_dmma = lib.ck_launch_count() - launches0
    finally:
        lib.ck_oz_configure(1, -1)
        del os.environ["CK_MG_INT8_MIN_TILES"]
    rp, re, _ = orc.joint_predict(orc.Params(HALF), 1, coords, z, targets, "euclidean")
    assert info == 0 and n_int8 > n_dmma  # the INT8 path launches two splits + one product per big update
    assert np.max(np.abs(pred - rp)) / np.max(np.abs(rp)) < 1e-9
    assert np.max(np.abs(var - re ** 2)) < 1e-9
    assert np.max(np.abs(pred - pred_d)) / np.max(np.abs(rp)) < 1e-11


@pytest.mark.parametrize("metric,params,tile", [(0, HALF, 128), (1, GENERIC_KM, 256)])
def test_native_handle_api_single_rank_vs_oracle(metric, params, tile):
    """csrc/ck_mgctx.cu (ck_mg_create / ck_mg_joint_cov / ck_mg_potrf / ck_mg_potrs_predict / ck_mg_logdet) on a 1 x 1 grid:
    the C-side schedule against the oracle and against the torch.distributed twin (same kernels: same bits expected up
    to the order of the launches, which does not change any sum)."""
    from cokrig_b200 import parallel
    coords, z, targets = _inputs(metric, 700, 650, 420, 17)
    native = parallel.NativeBlockCyclic(1, 1, tile=tile)
    twin = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=tile)
    for i_pred in (0, 1):
        pred, var, info = native.solve(coords, z, targets, params, 2, i_pred, metric)
        rp, re, _ = orc.joint_predict(orc.Params(params), i_pred, coords, z, targets, "haversine" if metric else "euclidean")
        assert info == 0
        assert np.max(np.abs(pred - rp)) / np.max(np.abs(rp)) < 1e-9
        assert np.max(np.abs(var - re ** 2)) < 1e-9
        p2, v2, _ = twin.solve(coords, z, targets, params, 2, i_pred, metric)
        assert np.max(np.abs(pred - p2)) <= 1e-13 * np.max(np.abs(rp)) and np.max(np.abs(var - v2)) <= 1e-13
    sigma = orc.joint_cov(orc.Params(params), coords, "haversine" if metric else "euclidean")
    assert abs(native.logdet() / np.linalg.slogdet(sigma)[1] - 1) < 1e-10
    assert set(native.timings) == {"assemble_ms", "factor_solve_ms", "reduce_ms"}
    native.close()


def test_native_handle_api_int8_updates_and_non_pd():
    from cokrig_b200 import parallel
    from cokrig_b200._lib import lib
    coords, z, targets = _inputs(0, 1500, 1400, 600, 23)
    os.environ["CK_MG_INT8_MIN_TILES"] = "48"  # read by ck_mg_create: INT8 trailing updates, column updates and inverted-tile TRSM
    try:
        launches0 = lib.ck_launch_count()
        native = parallel.NativeBlockCyclic(1, 1, tile=256)
        pred, var, info = native.solve(coords, z, targets, HALF, 2, 1, 0)
        n_int8 = lib.ck_launch_count() - launches0
        twin = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=256, lookahead=True)
        p2, v2, _ = twin.solve(coords, z, targets, HALF, 2, 1, 0)
    finally:
        del os.environ["CK_MG_INT8_MIN_TILES"]
    launches0 = lib.ck_launch_count()
    plain = parallel.NativeBlockCyclic(1, 1, tile=256)
    plain.solve(coords, z, targets, HALF, 2, 1, 0)
    assert n_int8 > lib.ck_launch_count() - launches0  # two splits + one product per big update
    rp, re, _ = orc.joint_predict(orc.Params(HALF), 1, coords, z, targets, "euclidean")
    assert info == 0
    assert np.max(np.abs(pred - rp)) / np.max(np.abs(rp)) < 1e-9 and np.max(np.abs(var - re ** 2)) < 1e-9
    assert np.max(np.abs(pred - p2)) / np.max(np.abs(rp)) < 1e-12
    bad = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .0, .0, -1.3]
    c2, z2, t2 = _inputs(0, 300, 280, 20, 2)
    _, _, info_bad = parallel.NativeBlockCyclic(1, 1, tile=128).solve(c2, z2, t2, bad, 2, 0, 0)
    _, _, info_twin = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=128).solve(c2, z2, t2, bad, 2, 0, 0)
    assert info_bad > 0 and info_bad == info_twin
    native.close()
    plain.close()


def test_c_example_program_vs_oracle(tmp_path):
    """examples/mg_cokrige.c: a plain C99 host (no Python, no torch in the process) drives ck_mg_* and must print the
    oracle's predictions for the same inputs (libc rand() stream reproduced through ctypes)."""
    import ctypes
    import re
    import shutil
    from cokrig_b200 import _lib
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not shutil.which("gcc") or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("gcc / CUDA headers not available")
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "mg_cokrige")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                           os.path.join(ROOT, "examples", "mg_cokrige.c"), "-L", libdir, "-lcokrig_b200", "-L", os.path.join(cuda, "lib64"),
                           "-lcudart", "-lm", "-o", exe])
    n0, n1, m = 1500, 1400, 300
    import torch
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    env = dict(os.environ, LD_LIBRARY_PATH=":".join([libdir, os.path.join(cuda, "lib64"), tl, os.environ.get("LD_LIBRARY_PATH", "")]))
    r = subprocess.run([exe, "0", "1", "/tmp/unused", str(n0), str(n1), str(m)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = [float(v) for v in re.search(r"pred\[0\.\.3\] = ([-0-9.e ]+);", r.stdout).group(1).split()]
    logdet = float(re.search(r"logdet ([-0-9.e]+)", r.stdout).group(1))
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(7)
    h = np.array([libc.rand() for _ in range(2 * (n0 + n1 + m) + n0 + n1)], dtype=np.float64) / 2147483647.0
    xy0, xy1 = h[: 2 * n0].reshape(n0, 2), h[2 * n0: 2 * (n0 + n1)].reshape(n1, 2)
    xyp = h[2 * (n0 + n1): 2 * (n0 + n1 + m)].reshape(m, 2)
    z = h[2 * (n0 + n1 + m):]
    params = [1.0, 0.8, 1.5, 1.5, 1.5, 0.1, 0.1, 0.1, 0.02, 0.02, -0.2]
    rp, _, _ = orc.joint_predict(orc.Params(params), 0, [xy0, xy1], [z[:n0], z[n0:]], xyp, "euclidean")
    assert np.max(np.abs(np.array(got) - rp[:4])) < 1e-9 * np.max(np.abs(rp))
    sigma = orc.joint_cov(orc.Params(params), [xy0, xy1], "euclidean")
    assert abs(logdet / np.linalg.slogdet(sigma)[1] - 1) < 1e-9


def test_block_cyclic_single_rank_reports_non_pd():
    from cokrig_b200 import parallel
    coords, z, targets = _inputs(0, 300, 280, 20, 2)
    bad = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .0, .0, -1.3]
    _, _, info = parallel.BlockCyclicCokriging(parallel.ProcessGrid(1, 1), tile=128).solve(coords, z, targets, bad, 2, 0, 0)
    assert info > 0


def test_gemm_nt_entry_point_vs_torch():
    import torch
    from cokrig_b200 import ops
    from cokrig_b200._lib import check, lib
    g = torch.Generator(device="cuda").manual_seed(1)
    for (m, n, k) in ((257, 130, 96), (128, 128, 128), (1, 5, 3), (300, 64, 513)):
        a = torch.randn(m, k, dtype=torch.float64, device="cuda", generator=g)
        b = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
        c = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
        ref = c - a @ b.T
        check(lib.ck_gemm_nt(ops._ptr(a), k, ops._ptr(b), k, ops._ptr(c), n, m, n, k, 1, ops._stream()))
        assert (c - ref).abs().max().item() < 1e-11 * k
        check(lib.ck_gemm_nt(ops._ptr(a), k, ops._ptr(b), k, ops._ptr(c), n, m, n, k, 0, ops._stream()))
        assert (c - a @ b.T).abs().max().item() < 1e-11 * k


def test_two_gpus_under_torchrun():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run tools/mg_check.py under torchrun on a multi-GPU box)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mg_check.py"), "--points", "3000", "--targets", "1000", "--tile", "512", "--native"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]

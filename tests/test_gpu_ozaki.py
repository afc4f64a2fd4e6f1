"""GPU tests of the INT8 tensor-core update path (csrc/ck_ozaki.cu) through the C ABI: the digit split is an
error-free recoding of a 55-bit fixed-point rounding, the tcgen05 product equals the FP64 product to FP64 rounding
level on plain, ragged, badly scaled and lower-masked cases, and ck_potrf / ck_trsm_lower / ck_potrs_predict give
the same factor, solves and predictions with the INT8 path on as with the FP64 DMMA kernel and LAPACK.

Tolerances: products 4 sqrt(k) 2^-53 (row max of A) x (row max of B) (the FP64 rounding level of the numpy reference); factor 1e-12,
predictions 1e-9 relative (BASELINE.json north star).
"""
import numpy as np
import pytest

import cokrig_oracle as orc

pytestmark = pytest.mark.gpu
S = 7


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cokrig_b200 import _lib
    return _lib


@pytest.fixture()
def int8_path_small(lib):
    """Hand even small trailing matrices to the INT8 kernel; restore the defaults afterwards."""
    lib.lib.ck_oz_configure(1, 256)
    yield
    lib.lib.ck_oz_configure(1, 1024)


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def _split(lib, x, want_a=True, want_b=True):
    import torch
    rows, k = x.shape
    fa = torch.zeros(max(lib.lib.ck_oz_slices_bytes(rows, k, 0), 16), dtype=torch.uint8, device="cuda") if want_a else None
    fb = torch.zeros(max(lib.lib.ck_oz_slices_bytes(rows, k, 1), 16), dtype=torch.uint8, device="cuda") if want_b else None
    sc = torch.zeros(lib.lib.ck_oz_scales_len(rows), dtype=torch.float64, device="cuda")
    lib.check(lib.lib.ck_oz_split(x.data_ptr(), x.stride(0), rows, k, fa.data_ptr() if want_a else None,
                                  fb.data_ptr() if want_b else None, sc.data_ptr(), _stream()), "ck_oz_split")
    return fa, fb, sc


def test_split_is_a_55_bit_fixed_point_recoding(lib):
    import torch
    rng = np.random.default_rng(0)
    rows, k = 200, 96
    x = rng.standard_normal((rows, k)) * np.exp(rng.uniform(-20, 5, (rows, 1)))
    x[7] = 0.0                      # an all-zero row
    x[11, 5] = 0.98 * 2.0 ** 3      # the rounding threshold of the exponent choice
    x[12, :] = 2.0 ** -40 * x[12, :]
    x[12, 3] = 1.0                  # one large entry: the small ones lose their low bits, nothing else
    fa, fb, sc = _split(lib, torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    kcn, w = k // 32, 2.0 ** (-8.0 * np.arange(S))
    da = fa.cpu().numpy().view(np.int8).reshape(-1, kcn, S, 2, 16, 8, 16).astype(np.float64)
    ra = np.tensordot(da, w, axes=([2], [0])).transpose(0, 3, 4, 1, 2, 5).reshape(-1, k) * sc.cpu().numpy()[:, None]
    db = fb.cpu().numpy().view(np.int8).reshape(-1, kcn, 2, S, 8, 8, 16).astype(np.float64)
    rb = np.tensordot(db, w, axes=([3], [0])).transpose(0, 3, 4, 1, 2, 5).reshape(-1, k)
    rb = rb * sc.cpu().numpy()[: rb.shape[0], None]
    rowmax = np.maximum(np.abs(x).max(axis=1, keepdims=True), 1e-300)
    bound = 2.0 ** -56 / 0.49  # half a unit of the 55-bit grid, rows scaled into (0.49, 0.98]
    assert (np.abs(ra[:rows] - x) / rowmax).max() <= bound
    assert (np.abs(rb[:rows] - x) / rowmax).max() <= bound
    assert (ra[rows:] == 0).all() and (ra[7] == 0).all()
    s = sc.cpu().numpy()
    assert (s[rows:] == 0).all() and s[7] == 0 and np.all(np.log2(s[s > 0]) % 1 == 0)  # powers of two


@pytest.mark.parametrize("m,n,k,lower,wide", [
    (128, 64, 32, False, False), (128, 64, 1024, False, False), (1, 1, 32, False, False), (130, 70, 96, False, False),
    (1000, 900, 256, False, False), (1024, 1024, 1024, True, False), (1500, 1500, 512, True, True),
    (3000, 700, 1024, True, False), (2304, 5000, 1024, False, True)])
def test_int8_product_equals_fp64_product(lib, m, n, k, lower, wide):
    import torch
    rng = np.random.default_rng(m + 3 * n + k)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((n, k))
    if wide:  # 13 decades of dynamic range inside rows: small entries keep their absolute, not relative, accuracy
        a *= np.exp(rng.uniform(-30, 0, (m, k)))
        b *= np.exp(rng.uniform(-30, 0, (n, k)))
    c0 = rng.standard_normal((m, n))
    ref = c0 - a @ b.T
    if lower:
        ref = np.where(np.tril(np.ones((m, n), dtype=bool)), ref, c0)
    ldc = n + (n % 2) + 2
    cbuf = torch.zeros((m, ldc), dtype=torch.float64, device="cuda")
    cbuf[:, :n] = torch.from_numpy(c0).cuda()
    fa, _, sa = _split(lib, torch.from_numpy(a).cuda(), True, False)
    _, fb, sb = _split(lib, torch.from_numpy(b).cuda(), False, True)
    lib.check(lib.lib.ck_oz_gemm(fa.data_ptr(), sa.data_ptr(), m, fb.data_ptr(), sb.data_ptr(), n, k, cbuf.data_ptr(), ldc,
                                 int(lower), 0, _stream()), "ck_oz_gemm")
    torch.cuda.synchronize()
    got = cbuf[:, :n].cpu().numpy()
    # every entry against the FP64 numpy product: the bound is dominated by the rounding of that reference itself
    rowmax = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :]
    assert (np.abs(got - ref) / (16 * np.sqrt(k) * 2.0 ** -53 * rowmax + 2.0 ** -51 * np.abs(ref))).max() < 1.0
    # 32 rows against an extended-precision (x87 long double, 64-bit mantissa) product: operands rounded to 2^-56
    # of their row maximum and dropped digit groups below that, exact integer accumulation, one final rounding of C.
    # A missing kept group would show up at >= 2e-13 of (row max) x (row max) at k = 1024.
    rows = np.sort(rng.choice(m, size=min(m, 32), replace=False))
    ld = np.longdouble
    ref_x = c0[rows].astype(ld) - a[rows].astype(ld) @ b.T.astype(ld)
    if lower:
        ref_x = np.where(np.arange(n)[None, :] <= rows[:, None], ref_x, c0[rows].astype(ld))
    err_x = np.abs(got[rows].astype(ld) - ref_x).astype(np.float64)
    bound_x = 32 * np.sqrt(k) * 2.0 ** -56 * rowmax[rows] + 2.0 ** -52 * np.maximum(np.abs(ref[rows]), np.abs(c0[rows]))
    assert (err_x / bound_x).max() < 1.0
    assert (cbuf[:, n:] == 0).all().item()  # nothing written past column n


def test_int8_product_rejects_bad_arguments(lib):
    import torch
    x = torch.zeros((128, 48), dtype=torch.float64, device="cuda")
    sc = torch.zeros(128, dtype=torch.float64, device="cuda")
    buf = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    assert lib.lib.ck_oz_split(x.data_ptr(), 48, 128, 48, buf.data_ptr(), None, sc.data_ptr(), _stream()) == lib.CK_ERR_ARG  # k % 32
    assert lib.lib.ck_oz_split(x.data_ptr(), 48, 128, 32, None, None, sc.data_ptr(), _stream()) == lib.CK_ERR_ARG          # no output
    assert lib.lib.ck_oz_gemm(buf.data_ptr(), sc.data_ptr(), 128, buf.data_ptr(), sc.data_ptr(), 64, 32, x.data_ptr(), 32, 0,
                              0, _stream()) == lib.CK_ERR_ARG                                                                   # ldc < n


def _spd(n, seed, nugget=0.01):
    xy = np.random.default_rng(seed).uniform(0, 1, (n, 2))
    return np.exp(-orc.distance_matrix(xy, xy, units=None) / 0.2) + nugget * np.eye(n)


@pytest.mark.parametrize("n", [1100, 2049, 3000])
def test_potrf_and_solve_with_int8_updates_vs_lapack(lib, int8_path_small, n):
    import torch
    from scipy.linalg import cholesky, solve_triangular
    from cokrig_b200 import ops
    launches0 = lib.lib.ck_launch_count()
    A = _spd(n, n)
    Lref = cholesky(A, lower=True)
    buf = torch.empty((n, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]
    buf.copy_(torch.from_numpy(A))
    f = ops.potrf(buf)
    assert f.info == 0
    L = f.lower().cpu().numpy()
    assert np.abs(L - Lref).max() / np.abs(Lref).max() < 1e-12
    assert np.abs(L @ L.T - A).max() < 1e-13 * n
    m = 1200  # >= 1024 right-hand sides: the solve takes the INT8 path as well
    B = np.random.default_rng(n + 1).standard_normal((m, n))
    rb = torch.empty((m, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]
    rb.copy_(torch.from_numpy(B))
    V = f.solve_lower(rb).cpu().numpy()
    Vref = solve_triangular(Lref, B.T, lower=True).T
    assert np.abs(V - Vref).max() / np.abs(Vref).max() < 1e-11
    # the same problem with the INT8 path off: both paths agree to rounding level
    lib.lib.ck_oz_configure(0, -1)
    buf2 = torch.empty((n, ops.padded_ld(n)), dtype=torch.float64, device="cuda")[:, :n]
    buf2.copy_(torch.from_numpy(A))
    L2 = ops.potrf(buf2).lower().cpu().numpy()
    lib.lib.ck_oz_configure(1, -1)
    assert np.abs(L - L2).max() / np.abs(Lref).max() < 1e-13
    assert lib.lib.ck_launch_count() > launches0


def test_joint_prediction_with_int8_updates_vs_oracle(lib, int8_path_small):
    """C1-style bivariate system (N = 2 x 1296), 1100 targets: predictions and variances within the north-star
    tolerance of the oracle with every big update on the INT8 tensor cores."""
    from cokrig_b200 import METRIC_EUCLID, ops
    params = [1, 1, 1.5, 1.5, 1.5, .2, .2, .2, .01, .01, -.6]
    P = orc.Params(params)
    grid = orc.expand_grid(xcount=36, ycount=36)
    _, _, z = orc.sim_fields(P, grid, seed=1)
    pc = np.random.default_rng(7).uniform(0, 1, (1100, 2))
    cd = [ops.coords_to_device(grid), ops.coords_to_device(grid)]
    factor = ops.potrf(ops.joint_cov(cd, params, 2, METRIC_EUCLID))
    cpd = ops.cross_cov(cd, ops.coords_to_device(pc), params, 2, 1, METRIC_EUCLID)
    pred, var = factor.predict(cpd, ops.to_device(np.hstack(z)), 1.01)
    ref_pred, ref_err, _ = orc.joint_predict(P, 1, [grid, grid], z, pc, "euclidean")
    assert factor.info == 0
    assert np.max(np.abs(pred.cpu().numpy() - ref_pred)) / np.abs(ref_pred).max() < 1e-9
    assert np.max(np.abs(var.cpu().numpy() - ref_err ** 2)) < 1e-9 * 1.01


def test_factor_and_solve_at_scale_with_lookahead(lib):
    """N = 18 000, 1 100 right-hand sides: large enough for the look-ahead beside the persistent INT8 kernel (panel chain
    of the next aggregate on a side stream).  Size-independent checks -- sampled rows of L L^T against Sigma, solve
    residuals, bit-reproducibility -- and agreement with the FP64 DMMA path on the same inputs."""
    import torch
    from cokrig_b200 import METRIC_EUCLID, ops
    n = 9000
    xy = ops.coords_to_device(np.random.default_rng(4).uniform(0, 1, (n, 2)))
    params = [1, .8, 1.5, 1.5, 1.5, .05, .05, .05, .02, .02, -.2]
    S0 = ops.joint_cov([xy, xy], params, 2, METRIC_EUCLID)
    f = ops.potrf(S0.clone())
    assert f.info == 0
    L = f.lower()
    idx = torch.randint(0, 2 * n, (100,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    assert ((L[idx] @ L.T) - S0[idx]).abs().max().item() < 1e-11
    m = 1100
    B = torch.randn((m, 2 * n), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    rb = torch.empty((m, ops.padded_ld(2 * n)), dtype=torch.float64, device="cuda")[:, : 2 * n]
    rb.copy_(B)
    V = f.solve_lower(rb)
    assert ((V[:50] @ L.T) - B[:50]).abs().max().item() < 1e-9
    f2 = ops.potrf(S0.clone())
    assert torch.equal(torch.tril(f2.L), L), "factorisation must be bit-reproducible (look-ahead included)"
    lib.lib.ck_oz_configure(0, -1)
    try:
        fd = ops.potrf(S0.clone())
        rb2 = torch.empty((m, ops.padded_ld(2 * n)), dtype=torch.float64, device="cuda")[:, : 2 * n]
        rb2.copy_(B)
        Vd = fd.solve_lower(rb2)
    finally:
        lib.lib.ck_oz_configure(1, -1)
    assert (fd.lower() - L).abs().max().item() / L.abs().max().item() < 1e-13
    assert (Vd - V).abs().max().item() / V.abs().max().item() < 1e-11


def test_concurrent_streams_share_no_scheduler_state(lib):
    """Batched windows (BASELINE config 4): several factorisations + solves enqueued on different CUDA streams run their
    INT8 update kernels side by side.  Each launch's dynamic tile scheduler draws from a counter slot that belongs to its
    (device, stream), so the concurrent results must equal -- bit for bit -- the ones computed alone on the default stream."""
    import torch
    from cokrig_b200 import METRIC_EUCLID, ops
    n, m, nsys = 1536, 700, 6   # N = 3072 per system: trailing updates on the INT8 path (>= 2048 rows)
    params = [1, .8, 1.5, 1.5, 1.5, .07, .07, .07, .02, .02, -.2]
    assert lib.lib.ck_oz_active(2 * n) == 1
    systems = []
    for s in range(nsys):
        rng = np.random.default_rng(100 + s)
        xy = ops.coords_to_device(rng.uniform(0, 1, (n, 2)))
        B = ops.to_device(rng.standard_normal((m, 2 * n)))
        systems.append((ops.joint_cov([xy, xy], params, 2, METRIC_EUCLID), B))

    def run(S0, B):
        f = ops.potrf(S0.clone())
        rb = torch.empty((m, ops.padded_ld(2 * n)), dtype=torch.float64, device="cuda")[:, : 2 * n]
        rb.copy_(B)
        return f, f.solve_lower(rb)

    alone = [run(S0, B) for S0, B in systems]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    for rep in range(2):
        together = []
        for s, (S0, B) in enumerate(systems):
            st = streams[s % len(streams)]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                together.append(run(S0, B))
        torch.cuda.synchronize()
        for (f1, v1), (f2, v2) in zip(alone, together):
            assert f2.info == 0
            assert torch.equal(torch.tril(f1.L), torch.tril(f2.L)) and torch.equal(v1, v2)

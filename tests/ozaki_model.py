"""numpy model of the INT8 update scheme of csrc/ck_ozaki.cu (test infrastructure, CPU only).

Same arithmetic as the kernels: per-row power-of-two scaling to (0.49, 0.98], ONE rounding to a 55-bit fixed-point
integer, exact recoding into 7 balanced base-256 digits, exact integer slice products G_g = sum_{p+q=g} A_p B_q^T
for g = 0..6 (int32 range checked), exact integer recombination three groups at a time + two FP64 fused multiply-adds,
exact power-of-two scaling."""
import numpy as np

S = 7
QBITS = 7 + 8 * (S - 1)


def split(x: np.ndarray):
    """-> (digits [S] float arrays holding integers in [-128, 127], most significant first; scales 2^(e-7))."""
    amax = np.abs(x).max(axis=1)
    f, e = np.frexp(amax)
    e = e + (f > 0.98)
    q = np.rint(x * np.exp2(QBITS - e)[:, None]).astype(np.int64)
    digits = []
    for _ in range(S):
        d = ((q + 128) & 255) - 128
        q = (q - d) >> 8
        digits.append(d.astype(np.float64))
    assert not q.any(), "leading digit out of range"
    scales = np.where(amax > 0, np.exp2(e - 7.0), 0.0)
    return digits[::-1], scales


def product(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """FP64-equivalent a @ b.T through the sliced integer products (K <= 1024 per call, like one kernel launch)."""
    assert a.shape[1] == b.shape[1] <= 1024
    da, sa = split(a)
    db, sb = split(b)
    G = []
    for g in range(S):
        G.append(sum(da[p] @ db[g - p].T for p in range(g + 1)))  # exact: |G| < 2^31 << 2^53
        assert np.abs(G[-1]).max() < 2 ** 31
    # recombination as in the kernel epilogue: three groups at a time in exact integer arithmetic (|t| < 2^45),
    # then two fused multiply-adds (the scalings by powers of two are exact, each addition rounds once)
    t0 = G[0] * 65536.0 + G[1] * 256.0 + G[2]
    t1 = G[3] * 65536.0 + G[4] * 256.0 + G[5]
    v = t1 * 2.0 ** -24 + t0
    v = G[6] * 2.0 ** -32 + v
    return v * 2.0 ** -16 * sa[:, None] * sb[None, :]


def cholesky_blocked(sigma: np.ndarray, nb: int, gemm) -> np.ndarray:
    """Right-looking blocked Cholesky whose trailing updates go through `gemm(a, b) = a @ b.T`."""
    from scipy.linalg import cholesky, solve_triangular
    a = sigma.copy()
    n = a.shape[0]
    for k in range(0, n, nb):
        k1 = min(k + nb, n)
        a[k:k1, k:k1] = cholesky(a[k:k1, k:k1], lower=True)
        if k1 < n:
            a[k1:, k:k1] = solve_triangular(a[k:k1, k:k1], a[k1:, k:k1].T, lower=True).T
            a[k1:, k1:] -= gemm(a[k1:, k:k1], a[k1:, k:k1])
    return np.tril(a)


def solve_blocked(L: np.ndarray, rhs: np.ndarray, nb: int, gemm) -> np.ndarray:
    """Target-major forward substitution V = rhs L^-T with the updates through `gemm`."""
    from scipy.linalg import solve_triangular
    r = rhs.copy()
    n = L.shape[0]
    for k in range(0, n, nb):
        k1 = min(k + nb, n)
        r[:, k:k1] = solve_triangular(L[k:k1, k:k1], r[:, k:k1].T, lower=True).T
        if k1 < n:
            r[:, k1:] -= gemm(r[:, k:k1], L[k1:, k:k1])
    return r

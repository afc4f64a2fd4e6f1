// ck_mg.cu -- per-rank assembly of the 2-D block-cyclic AUGMENTED cokriging system (multi-GPU path).
//
// Global array (never materialised in one place), square tiles of `tb` elements:
//
//        cols: TC = ceil(N / tb) tile columns over the stacked data (process 0 then 1), padded to TC*tb
//   rows 0 .. TC-1        data tiles      Sigma (lower tiles J <= I only), padded with the identity
//   rows TC .. TC+TE-1    target tiles    tb-1 prediction targets each: c^T = Cov(target, data)  (ck_cross_cov rows)
//                                         + the stacked data vector z in the LAST row of every tile
//
// A right-looking Cholesky sweep over the tile columns, applied to ALL rows, leaves L in the data
// tiles and V = C L^-T (rows L^-1 c) and y = L^-1 z in the target tiles -- the factorisation and the
// triangular solve of src/joint_prediction.py:68-73 in one pass, with no second streaming of L.
// Tile (I, J) lives on rank (I mod P, J mod Q) at local tile (I div P, J div Q).  Every rank
// assembles its own tiles straight from the (replicated) coordinates: Sigma is never communicated.
#include "ck_common.cuh"

// entries outside the valid (rows x cols) part of a tile: identity on the diagonal of diagonal tiles, else 0
__global__ void __launch_bounds__(256) ck_tile_pad_kernel(double* __restrict__ t, long long ld, int tb, int vrows, int vcols,
                                                          int diag) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= tb) return;
  if (r < vrows && c < vcols) return;
  t[(long long)r * ld + c] = (diag && r == c) ? 1.0 : 0.0;
}

extern "C" ck_i64 ck_mg_local_tiles(ck_i64 ntiles, ck_i64 nprocs, ck_i64 rank) {
  if (ntiles <= rank) return 0;
  return (ntiles - rank + nprocs - 1) / nprocs;
}

extern "C" int ck_mg_assemble(const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* xyp, ck_i64 m,
                              const double* z, const double* params, int n_procs, int i_pred, int metric, ck_i64 tb,
                              int P, int p, int Q, int q, double* local, ck_i64 ld, void* stream) {
  CkParams prm;
  int rc = ck_unpack_params(params, n_procs, &prm);
  if (rc) return rc;
  if (n_procs == 1) n1 = 0;
  CK_REQUIRE(n0 >= 0 && n1 >= 0 && m >= 0, "negative size");
  CK_REQUIRE(i_pred >= 0 && i_pred < n_procs, "i_pred out of range");
  CK_REQUIRE(tb >= 128 && tb % 128 == 0, "tile size must be a multiple of 128 (got %lld)", (long long)tb);
  CK_REQUIRE(P >= 1 && Q >= 1 && p >= 0 && p < P && q >= 0 && q < Q, "bad process grid (%d,%d) of %dx%d", p, q, P, Q);
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  const ck_i64 N = n0 + n1;
  if (N == 0) return CK_OK;
  CK_REQUIRE(xy0 && (n1 == 0 || xy1) && local && z, "null pointer");
  CK_REQUIRE(m == 0 || xyp, "null target coordinates");
  const ck_i64 cap = tb - 1;
  const ck_i64 TC = (N + tb - 1) / tb, TE = (m + cap - 1) / cap;
  const ck_i64 lct = ck_mg_local_tiles(TC, Q, q);
  CK_REQUIRE(ld >= lct * tb, "ld (%lld) < local columns (%lld)", (long long)ld, (long long)(lct * tb));
  cudaStream_t st = ck_stream(stream);
  const double* xy[2] = {xy0, xy1};
  const ck_i64 nn[2] = {n0, n1};
  const ck_i64 start[3] = {0, n0, N};  // stacked index range of process j: [start[j], start[j+1])
  CkMatern S[2][2], C[2];
  for (int a = 0; a < n_procs; ++a)
    for (int b = 0; b < n_procs; ++b)
      if ((rc = ck_block_matern(prm, a < b ? a : b, a < b ? b : a, 1, &S[a][b]))) return rc;
  for (int b = 0; b < n_procs; ++b)
    if ((rc = ck_block_matern(prm, i_pred, b, 1, &C[b]))) return rc;
  (void)nn;
  const dim3 pad_grid((unsigned)((tb + 255) / 256), (unsigned)tb);

  for (ck_i64 I = p; I < TC + TE; I += P) {
    const ck_i64 li = I / P;
    for (ck_i64 J = q; J < TC; J += Q) {
      const ck_i64 lj = J / Q;
      if (I < TC && J > I) continue;  // above the diagonal: never referenced
      double* tile = local + li * tb * ld + lj * tb;
      const ck_i64 c_lo = J * tb, c_hi = (c_lo + tb < N) ? c_lo + tb : N;
      if (I < TC) {
        const ck_i64 r_lo = I * tb, r_hi = (r_lo + tb < N) ? r_lo + tb : N;
        for (int a = 0; a < n_procs; ++a) {
          const ck_i64 ra = r_lo > start[a] ? r_lo : start[a], rb = r_hi < start[a + 1] ? r_hi : start[a + 1];
          if (ra >= rb) continue;
          for (int b = 0; b < n_procs; ++b) {
            const ck_i64 ca = c_lo > start[b] ? c_lo : start[b], cb = c_hi < start[b + 1] ? c_hi : start[b + 1];
            if (ca >= cb) continue;
            rc = ck_block_launch(xy[a] + 2 * (ra - start[a]), rb - ra, xy[b] + 2 * (ca - start[b]), cb - ca, metric, S[a][b], 1,
                                 tile + (ra - r_lo) * ld + (ca - c_lo), ld, nullptr, 0, 0, st);
            if (rc) return rc;
          }
        }
        if (r_hi - r_lo < tb || c_hi - c_lo < tb) {
          ck_tile_pad_kernel<<<pad_grid, 256, 0, st>>>(tile, ld, (int)tb, (int)(r_hi - r_lo), (int)(c_hi - c_lo), I == J);
          CK_LAUNCH_CHECK();
        }
      } else {
        const ck_i64 t_lo = (I - TC) * cap, t_hi = (t_lo + cap < m) ? t_lo + cap : m;
        for (int b = 0; b < n_procs; ++b) {
          const ck_i64 ca = c_lo > start[b] ? c_lo : start[b], cb = c_hi < start[b + 1] ? c_hi : start[b + 1];
          if (ca >= cb || t_hi <= t_lo) continue;
          rc = ck_block_launch(xyp + 2 * t_lo, t_hi - t_lo, xy[b] + 2 * (ca - start[b]), cb - ca, metric, C[b], 1,
                               tile + (ca - c_lo), ld, nullptr, 0, 0, st);
          if (rc) return rc;
        }
        // rows past the last target and columns past N are zero; the last row carries z
        ck_tile_pad_kernel<<<pad_grid, 256, 0, st>>>(tile, ld, (int)tb, (int)(t_hi - t_lo), (int)(c_hi - c_lo), 0);
        CK_LAUNCH_CHECK();
        CK_CUDA(cudaMemcpyAsync(tile + (tb - 1) * ld, z + c_lo, (size_t)(c_hi - c_lo) * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  return CK_OK;
}

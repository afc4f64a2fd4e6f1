/* cokrig.h -- C ABI of libcokrig_b200.so: the B200 (sm_100a) cokriging hot path.
 *
 * The reference (91Mrwu/sif-xco2-cokriging) is pure Python and has no FFI of its own; its hot path
 * is the set of numpy/scipy/sklearn/pandas calls cited below (paths relative to the reference root).
 * Each entry point here replaces one of those call groups; the drop-in Python modules under
 * sif-xco2-cokriging_b200/src bind them with ctypes (see INTEGRATION.md for the binding stubs).
 *
 * Conventions
 *   - plain C types only: pointers, sizes, scalars.  No torch / C++ types cross this boundary.
 *   - every `const double*` / `double*` named *_dev (and all matrix/vector arguments unless marked
 *     HOST) is a DEVICE pointer on the current CUDA device; `params` vectors are HOST pointers.
 *   - matrices are FP64, row-major, with an explicit leading dimension `ld` (elements per row).
 *   - coordinates are (n x 2) row-major FP64: [lat, lon] in degrees for CK_METRIC_HAVERSINE,
 *     [x, y] for CK_METRIC_EUCLID (src/fields.py:318-342).
 *   - `params` is the reference's flat parameter vector, MaternParams.get_values() order
 *     (src/model.py:130-152): n_procs=2: sigma11,sigma22, nu11,nu12,nu22, len11,len12,len22,
 *     nugget11,nugget22, rho12 (11 values); n_procs=1: sigma, nu, len_scale, nugget (4 values).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; nothing synchronises unless stated.
 *   - return value: CK_OK (0), or a negative CK_ERR_* code; ck_last_error() gives the text.
 *     Numerical failures (non-positive pivot) are reported through the `info` outputs, LAPACK
 *     style, never through the return value.
 *   - no hidden device allocations: scratch comes from the caller via *_workspace_bytes().
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns CK_ERR_CUDA.
 */
#ifndef COKRIG_H
#define COKRIG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t ck_i64;

#define CK_OK 0
#define CK_ERR_ARG (-1)         /* invalid argument (null pointer, negative size, bad enum, ld too small) */
#define CK_ERR_CUDA (-2)        /* CUDA runtime error (launch failure, no device, ...) */
#define CK_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels implement (e.g. n_procs > 2) */

#define CK_METRIC_EUCLID 0    /* scipy.spatial.distance.cdist(X1, X2)              src/fields.py:342 */
#define CK_METRIC_HAVERSINE 1 /* sklearn haversine_distances(radians(X)) * 6371     src/fields.py:334-336 */

int ck_version(void);
const char* ck_last_error(void);
/* Number of kernels this library has launched in the calling process so far (monotonic). */
long long ck_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  Matern (cross-)covariance assembly
 * ---------------------------------------------------------------------------------------------- */

/* Elementwise covariance of a distance array: out[k] = scale * rho(h[k] | nu, len_scale)
 * (+ nugget where h[k] == 0).  Replaces model._matern_correlation + the scaling in
 * MultivariateMatern.correlation/covariance/cross_covariance  (src/model.py:188-207, 354-385).
 * scale = sigma_i^2 (covariance) or rho_ij * prod(sigma) (cross_covariance); pass nugget = 0 for
 * cross blocks or use_nugget=False; scale = 1, nugget = 0 gives the bare correlation. */
int ck_matern_eval(const double* h_dev, ck_i64 n, double scale, double nu, double len_scale, double nugget,
                   double* out_dev, void* stream);

/* Pairwise distances only: out[i*ld + j] = d(xy1[i], xy2[j]).
 * Replaces fields.distance_matrix(X1, X2, units, fast_dist)  (src/fields.py:318-342). */
int ck_distance_block(const double* xy1_dev, ck_i64 n1, const double* xy2_dev, ck_i64 n2, int metric,
                      double* out_dev, ck_i64 ld, void* stream);

/* Fused coordinates -> distance -> Matern -> scale/nugget for one block:
 *   out[i*ld + j] = scale * rho(d(xy1[i], xy2[j])) (+ nugget where d == 0).
 * If out_t != NULL the transposed block is written as well: out_t[j*ld_t + i] = out[i*ld + j].
 * symmetric != 0 requires xy1 == xy2 (same points): only tiles on/above the diagonal are evaluated
 * and mirrored into the lower triangle of `out` (out_t is ignored).
 * Replaces distance_matrix + MultivariateMatern.covariance/cross_covariance on one block
 * (src/joint_prediction.py:94-122, src/point_prediction.py:98-113, src/sim.py:45-50). */
int ck_matern_block(const double* xy1_dev, ck_i64 n1, const double* xy2_dev, ck_i64 n2, int metric, double scale,
                    double nu, double len_scale, double nugget, double* out_dev, ck_i64 ld, double* out_t_dev,
                    ck_i64 ld_t, int symmetric, void* stream);

/* Joint covariance of the stacked data vector, block order [[C00, C01], [C01^T, C11]]
 * (n_procs = 2, N = n0 + n1) or the single block C00 (n_procs = 1, N = n0; xy_b ignored).
 * Replaces joint_prediction.Predictor._joint_cov (src/joint_prediction.py:124-153),
 * point_prediction.Predictor._cov_blocks (src/point_prediction.py:98-113) and
 * sim.BivariateRandomField._joint_cov_matrix (src/sim.py:45-50). */
int ck_joint_cov(const double* xy0_dev, ck_i64 n0, const double* xy1_dev, ck_i64 n1, const double* params /*HOST*/,
                 int n_procs, int metric, double* sigma_dev, ck_i64 ld, void* stream);

/* Covariances between m prediction targets of process i_pred and the stacked data, TARGET-MAJOR:
 *   cpd[c*ld + k] = Cov(Y_i(target c), Z(stacked datum k)),  k over process 0 then process 1
 * (own-process columns use covariance(use_nugget=True), the others cross_covariance).
 * This is the transpose of joint_prediction.Predictor._pred_cross_cov (src/joint_prediction.py:104-122). */
int ck_cross_cov(const double* xy0_dev, ck_i64 n0, const double* xy1_dev, ck_i64 n1, const double* xyp_dev, ck_i64 m,
                 const double* params /*HOST*/, int n_procs, int i_pred, int metric, double* cpd_dev, ck_i64 ld,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3  dense FP64 Cholesky, triangular solve, simple-cokriging prediction
 * ---------------------------------------------------------------------------------------------- */

/* Scratch needed by ck_potrf for an n x n matrix: the inverses of the 128 x 128 diagonal blocks of L (kept for
 * the blocked triangular solves) and, for n >= 2048, the int8 digit slices of one n x 1024 panel (INT8 tensor-core
 * updates, below).  ck_trsm_lower / ck_potrs_predict reuse the slice scratch: solves sharing one factor must not
 * run concurrently. */
size_t ck_potrf_workspace_bytes(ck_i64 n);

/* In-place lower Cholesky A = L L^T (row-major; the strict upper triangle is not referenced and is
 * left untouched).  *info_dev (device int) = 0 on success, k > 0 if the leading minor of order k is
 * not positive definite (LAPACK dpotrf convention; the factor is then invalid).
 * Replaces scipy.linalg.cho_factor(lower=True) / cholesky(lower=True)
 * (src/joint_prediction.py:68-69, src/point_prediction.py:209-210, src/sim.py:42). */
int ck_potrf(double* a_dev, ck_i64 n, ck_i64 ld, void* ws_dev, int* info_dev, void* stream);

/* Forward substitution with many right-hand sides stored TARGET-MAJOR: on entry rhs is (nrhs x n),
 * row c holding one right-hand side; on exit row c holds L^{-1} rhs_c.  Needs the ws of ck_potrf. */
int ck_trsm_lower(const double* l_dev, ck_i64 n, ck_i64 ld, const void* ws_dev, double* rhs_dev, ck_i64 nrhs,
                  ck_i64 ld_rhs, void* stream);

/* Simple cokriging from the factor: cpd is ((m+1) x n) as produced by ck_cross_cov with one spare
 * row; z (n) is the stacked data vector.  On exit rows 0..m-1 of cpd hold V = L^{-1} c and row m
 * holds y = L^{-1} z;  pred[c] = V_c . y  (= c^T Sigma^{-1} z),  var[c] = c0 - |V_c|^2
 * (= c0 - c^T Sigma^{-1} c, may be slightly negative; the caller applies nan_to_num(sqrt())).
 * Replaces cho_solve + the two matmuls + diagonal of src/joint_prediction.py:68-78. */
int ck_potrs_predict(const double* l_dev, ck_i64 n, ck_i64 ld, const void* ws_dev, double* cpd_dev, ck_i64 m,
                     ck_i64 ld_c, const double* z_dev, double c0, double* pred_dev, double* var_dev, void* stream);

/* logdet(A) = 2 sum_k log L_kk from the factor -> *out_dev (device double). */
int ck_logdet(const double* l_dev, ck_i64 n, ck_i64 ld, double* out_dev, void* stream);

/* Gaussian negative log-likelihood 0.5 (z^T Sigma^-1 z + logdet Sigma + N log 2pi) of the stacked
 * data under the model: assembles Sigma into sigma_dev (N x N, ld), factors it in place, solves.
 * scratch_dev: at least N doubles.  out_dev[0] = nll, out_dev[1] = quadratic form, out_dev[2] = logdet.
 * No counterpart in the current reference src/ (the likelihood fit of the north star). */
int ck_nll(const double* xy0_dev, ck_i64 n0, const double* xy1_dev, ck_i64 n1, const double* params /*HOST*/,
           int n_procs, int metric, const double* z_dev, double* sigma_dev, ck_i64 ld, void* ws_dev,
           double* scratch_dev, double* out_dev, int* info_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  empirical (cross-)semivariogram pair binning
 * ---------------------------------------------------------------------------------------------- */

/* Guard band.  Haversine distances from the device differ from sklearn's (glibc sin/asin) by a few
 * ulp, so a pair sitting within ~8 ulp of a decision boundary (max_dist, a bin edge, the extrema that
 * define the edges) could be decided differently.  Those pairs are not decided on the device: they are
 * returned as index pairs and the host re-evaluates them with libm (fields._host_distance), which
 * makes bin centres and bin counts bit-identical to the reference.  Euclidean distances are
 * bit-identical to scipy.cdist and need no guard (lists stay empty). */

size_t ck_vario_minmax_workspace_bytes(ck_i64 na, ck_i64 nb);

/* Pair tiles.  All K2 passes visit the pairs in tiles of (tile_rows x tile_cols) points of (A x B); the
 * tile decomposition depends only on (na, nb).  Every pass takes a tile-ROW range
 * [tile_row_begin, tile_row_end): (0, -1) = all rows (single GPU); on several GPUs each rank passes its
 * share of the rows of A (row-block partition, cokrig_b200/parallel.py) and the per-tile partials are
 * combined afterwards, so results do not depend on the number of ranks. */
ck_i64 ck_vario_tile_rows(ck_i64 na);
ck_i64 ck_vario_tile_cols(ck_i64 nb);
int ck_vario_tile_shape(int* rows_per_tile /*HOST*/, int* cols_per_tile /*HOST*/);

/* Pass 1: over all pairs (a < b if same_field, else all a, b) with d <= max_dist (+ guard band):
 *   out_dev[0] = min{d : d > 0} (+inf if none), out_dev[1] = max d (-inf if none),
 *   out_dev[2] = number of such pairs (as double, exact below 2^53);  ws_dev keeps per-tile extrema.
 * Replaces the two reductions of fields._construct_variogram_bins (src/fields.py:394-395). */
int ck_vario_minmax(const double* xya_dev, ck_i64 na, const double* xyb_dev, ck_i64 nb, int metric, int same_field,
                    double max_dist, ck_i64 tile_row_begin, ck_i64 tile_row_end, double* out_dev, void* ws_dev,
                    void* stream);

/* Index pairs (a, b) with 0 < d <= lo or d >= hi (and d <= max_dist + guard): the candidates for the
 * exact extrema.  pairs_dev holds 2*capacity entries; *count_dev may exceed capacity (then only the
 * first `capacity` pairs were stored).  ws_dev is the workspace filled by ck_vario_minmax. */
int ck_vario_candidates(const double* xya_dev, ck_i64 na, const double* xyb_dev, ck_i64 nb, int metric, int same_field,
                        double max_dist, double lo, double hi, ck_i64 tile_row_begin, ck_i64 tile_row_end,
                        const void* ws_dev, ck_i64* pairs_dev, ck_i64 capacity, unsigned long long* count_dev,
                        void* stream);

size_t ck_vario_bin_workspace_bytes(ck_i64 na, ck_i64 nb, int n_bins);

/* Pass 2: bin every retained pair with pandas.cut(include_lowest=True) semantics on `edges`
 * (n_bins + 1 ascending edges, HOST): bin k <=> edges[k] < d <= edges[k+1], d == edges[0] -> bin 0,
 * d > edges[n_bins] dropped.  Cloud value 0.5 (ra - rb)^2 (covariogram == 0) or ra * rb, with
 * ra = va - mean_a, rb = vb - mean_b (src/fields.py:378-386).  counts_dev: n_bins uint64 (bit-exact);
 * sums_dev: n_bins FP64 sums accumulated in a fixed, launch-geometry-independent order.
 * flagged_dev (may be NULL: no guard): pairs inside the guard band are appended there instead of
 * being binned (2*flag_capacity entries, *flag_count_dev as in ck_vario_candidates).
 * Replaces _variogram_cloud + pd.cut + groupby.agg  (src/fields.py:192-222). */
int ck_vario_bin(const double* xya_dev, const double* va_dev, ck_i64 na, double mean_a, const double* xyb_dev,
                 const double* vb_dev, ck_i64 nb, double mean_b, int metric, int same_field, int covariogram,
                 double max_dist, const double* edges /*HOST*/, int n_bins, unsigned long long* counts_dev,
                 double* sums_dev, ck_i64* flagged_dev, ck_i64 flag_capacity, unsigned long long* flag_count_dev,
                 void* ws_dev, void* stream);

/* The two stages of ck_vario_bin, for the row-block multi-GPU partition: _tiles fills the per-tile
 * partials of tile rows [tile_row_begin, tile_row_end) inside ws_dev (all other partials are set to
 * zero); the ranks then sum their partial arrays (each entry has exactly one non-zero contributor, so the
 * sum is exact and order-free) and _reduce combines the tiles in the fixed order of the single-GPU path:
 * counts and sums are bit-identical for any number of ranks.  The partial arrays start
 * ck_vario_bin_partials_offset(n_bins) bytes into ws_dev: tiles*n_bins doubles, then (256-byte aligned)
 * tiles*n_bins uint32, tiles = ck_vario_tile_rows(na) * ck_vario_tile_cols(nb). */
size_t ck_vario_bin_partials_offset(int n_bins);
int ck_vario_bin_tiles(const double* xya_dev, const double* va_dev, ck_i64 na, double mean_a, const double* xyb_dev,
                       const double* vb_dev, ck_i64 nb, double mean_b, int metric, int same_field, int covariogram,
                       double max_dist, const double* edges /*HOST*/, int n_bins, ck_i64 tile_row_begin,
                       ck_i64 tile_row_end, ck_i64* flagged_dev, ck_i64 flag_capacity, unsigned long long* flag_count_dev,
                       void* ws_dev, void* stream);
int ck_vario_bin_reduce(ck_i64 na, ck_i64 nb, int n_bins, const void* ws_dev, unsigned long long* counts_dev,
                        double* sums_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  batched local-neighbourhood cokriging (point prediction)
 * ---------------------------------------------------------------------------------------------- */

size_t ck_local_predict_workspace_bytes(ck_i64 m, ck_i64 kmax);

/* Pass 1: k_dev[c] = number of data (both processes) within max_dist of target c
 * (cv != 0: data of process i_pred at distance exactly 0 are excluded, src/point_prediction.py:140-142), and
 * seg_dev[8 c .. 8 c + 7] = the same count split by (process, quarter of the index range) -- the offsets pass 2 needs to
 * compact the neighbours in the reference's order without counting again.  Coordinate arrays 16-byte aligned. */
int ck_local_count(const double* xy0_dev, ck_i64 n0, const double* xy1_dev, ck_i64 n1, const double* xyp_dev, ck_i64 m,
                   int n_procs, int i_pred, int metric, double max_dist, int cv, int* k_dev, int* seg_dev, void* stream);

/* Pass 2: for every target: gather neighbours (process 0 then 1, index order), re-compute the local covariance from
 * coordinates, factor, solve.  pred/sd follow src/point_prediction.py:200-241: NaN when there is no neighbour or the
 * local matrix is not positive definite (info_dev[c] > 0); sd = sqrt(max(c0 - w.c, 0)) with NaN -> 0 (info_dev[c] = -1
 * when c0 - w.c <= 0: the reference's augmented-matrix check fails there and it warns).
 * params_sigma: the model the local covariance MATRIX is built from (the reference freezes Sigma when the Predictor is
 * constructed, src/point_prediction.py:42); params_pred (NULL = params_sigma): the model of the target-to-neighbour
 * VECTOR, evaluated at call time (src/point_prediction.py:115-125); c0 = covariance(i, 0, use_nugget=True) as the caller
 * computed it (src/point_prediction.py:66).  kmax = max_c k_dev[c]; k_dev / seg_dev from pass 1.
 * sigma_dev (optional, NULL = off): the joint covariance of the stacked data as ck_joint_cov writes it ((n0 + n1)^2, row-major,
 * leading dimension ld_sigma) -- the device counterpart of the reference's stored Predictor.Sigma (src/point_prediction.py:98-113).
 * When given, the entries of every local matrix are GATHERED from it (src/point_prediction.py:159-179 does the same with
 * np.ix_) instead of being re-computed from the coordinates: faster, at the price of N^2 memory. */
int ck_local_predict(const double* xy0_dev, const double* z0_dev, ck_i64 n0, const double* xy1_dev, const double* z1_dev,
                     ck_i64 n1, const double* xyp_dev, ck_i64 m, const double* params_sigma /*HOST*/,
                     const double* params_pred /*HOST or NULL*/, int n_procs, int i_pred, int metric, double max_dist, int cv,
                     double c0, const double* sigma_dev, ck_i64 ld_sigma, const int* k_dev, const int* seg_dev, ck_i64 kmax,
                     double* pred_dev, double* sd_dev, int* info_dev, void* ws_dev, void* stream);

/* Profiling aid: when set to a device buffer of 6 int64 counters, every following ck_local_predict launch adds the
 * cycles thread 0 of each CTA spent in [0] neighbour scan, [1] covariance entries, [2] DMMA main loops, [3] diagonal
 * blocks, [4] panel products + stores, [5] the final reductions.  NULL switches it off (default).  Only a
 * -DCK_LOCAL_PROFILE build carries the counters; the product build returns CK_ERR_UNSUPPORTED for a non-NULL buffer. */
int ck_local_debug_buffer(void* dev_counters);

/* ------------------------------------------------------------------------------------------------
 * Temporal cross-correlation: src/stat_tools.py:128-160 compute_xcor_nd for EVERY lag of `lags` in one
 * pass over two (ncell x T) row-major arrays (NaN = missing), optionally preceded by the per-cell linear detrend of
 * apply_detrend (:56-75), plus the arg-max over lags of optim_lag_nd (:181-233).  numpy.ma semantics: Python-slice lag
 * windows (negative lags included), sum(XY) over jointly valid entries, sum(XX) / sum(YY) over each series' own valid
 * entries; NaN for fewer than tau joint entries (tau > 0), empty support or zero denominator.
 * xcor_dev: nlag x ncell; best_idx_dev / best_xcor_dev (optional): index into `lags` of the largest |xcor| (0 if all NaN)
 * and its value.  lags: HOST array, at most 64 entries.
 * ---------------------------------------------------------------------------------------------- */
int ck_xcor_lags(const double* z1_dev, const double* z2_dev, ck_i64 ncell, int t_len, const int* lags /*HOST*/, int nlag,
                 int tau, int detrend, double* xcor_dev, int* best_idx_dev, double* best_xcor_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU building blocks (one process per GPU; the collectives are NCCL broadcasts issued by the host
 * side, cokrig_b200/parallel.py).  The large system is the AUGMENTED array
 *     [ Sigma (N x N, lower tiles) ; C^T (targets x N) + z ]      square tiles of `tb` elements,
 * 2-D block-cyclic over a P x Q process grid: tile (I, J) on rank (I mod P, J mod Q), local tile
 * (I div P, J div Q).  One right-looking Cholesky sweep over the tile columns applied to all rows
 * leaves L, V = L^-1 c and y = L^-1 z (src/joint_prediction.py:68-73 without a second pass over L).
 * No counterpart in the reference (single process); SURVEY 8(e).
 * ---------------------------------------------------------------------------------------------- */

/* Number of tiles t in [0, ntiles) with t mod nprocs == rank. */
ck_i64 ck_mg_local_tiles(ck_i64 ntiles, ck_i64 nprocs, ck_i64 rank);

/* Assemble this rank's tiles of the augmented array from the replicated coordinates (K1 per tile).
 * Row tiles 0 .. TC-1 (TC = ceil(N/tb)) hold Sigma (tiles with J <= I only; the pad beyond N is the
 * identity); row tiles TC + e hold tb-1 targets each (rows of ck_cross_cov) and z in their last row.
 * local_dev: (local row tiles * tb) x ld, ld >= local column tiles * tb. */
int ck_mg_assemble(const double* xy0_dev, ck_i64 n0, const double* xy1_dev, ck_i64 n1, const double* xyp_dev, ck_i64 m,
                   const double* z_dev, const double* params /*HOST*/, int n_procs, int i_pred, int metric, ck_i64 tb,
                   int P, int p, int Q, int q, double* local_dev, ck_i64 ld, void* stream);

/* C = A B^T (subtract == 0) or C -= A B^T: A (m x k), B (n x k), C (m x n), row-major, FP64 DMMA. */
int ck_gemm_nt(const double* a_dev, ck_i64 lda, const double* b_dev, ck_i64 ldb, double* c_dev, ck_i64 ldc, ck_i64 m,
               ck_i64 n, ck_i64 k, int subtract, void* stream);

/* Trailing update of the local part of the block-cyclic array: C -= A B^T on the local tiles whose
 * global indices satisfy J <= I (J == I: lower triangle of the tile only).  Local tile (li, lj) of C is
 * global tile (row_tile0 + li * row_tile_step, col_tile0 + lj * col_tile_step); m, n whole tiles. */
int ck_mg_update(const double* a_dev, ck_i64 lda, const double* b_dev, ck_i64 ldb, double* c_dev, ck_i64 ldc, ck_i64 m,
                 ck_i64 n, ck_i64 k, ck_i64 tb, ck_i64 row_tile0, ck_i64 row_tile_step, ck_i64 col_tile0,
                 ck_i64 col_tile_step, void* stream);

/* Row-wise partial sums over the local columns: out_vy[r] = v_r . y_r, out_vv[r] = |v_r|^2 with
 * v_r = v + r*ldv, y_r = y + r*ldy (ldy = 0: one shared y), fixed summation order. */
int ck_row_dots(const double* v_dev, ck_i64 ldv, ck_i64 nrows, ck_i64 ncols, const double* y_dev, ck_i64 ldy,
                double* out_vy_dev, double* out_vv_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU handle API: the whole block-cyclic sweep driven from C (ck_mg_create /
 * ck_mg_joint_cov / ck_mg_potrf / ck_mg_potrs_predict / ck_mg_destroy).  One process (or host
 * thread) per GPU; every rank makes the same sequence of calls with the same arguments (replicated
 * coordinates and data).  The context owns a world, a process-row and a process-column NCCL
 * communicator (bound at run time: dlopen of libnccl.so.2, env CK_NCCL_LIB overrides; world = 1
 * never touches NCCL) and a high-priority panel stream for the one-panel look-ahead.  Together the
 * three calls replace, for one system spread over P x Q GPUs, what the reference does in one
 * process: _joint_cov / _pred_cross_cov (src/joint_prediction.py:94-153), cho_factor, cho_solve and
 * the prediction / variance products (src/joint_prediction.py:68-78).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ck_mg_ctx ck_mg_ctx;

/* 128-byte NCCL unique id (HOST buffer): call on ONE rank and hand the bytes to every rank's
 * ck_mg_create through whatever channel the host program has (MPI, a socket, a file). */
int ck_mg_unique_id(void* id128 /*HOST*/);

/* Collective over all `world` ranks: rank = p * Q + q on a P x Q process grid, square tiles of `tile`
 * elements (multiple of 128, at most 1024).  The current CUDA device of the calling thread becomes the
 * context's device.  nccl_unique_id128 may be NULL when world == 1.  Env knobs read here:
 * CK_MG_LOOKAHEAD (default 1), CK_MG_PANEL_SMS (default: adaptive per tile column), CK_MG_INT8_MIN_TILES. */
int ck_mg_create(ck_mg_ctx** out /*HOST*/, int world, int rank, int P, int Q, ck_i64 tile, const void* nccl_unique_id128 /*HOST*/);
/* Destroys the communicators and the panel stream.  The caller synchronises the device (or at least the streams it passed
 * to the ck_mg_* calls) first; a context is used by one host thread at a time. */
int ck_mg_destroy(ck_mg_ctx* h);
int ck_mg_grid(const ck_mg_ctx* h, int* pq4 /*HOST: P, Q, p, q*/);

/* Device bytes this rank needs for a system of n_data stacked data and m targets: its tiles of the
 * augmented array, the double-buffered panel operands, exchange buffers and the int8 slice scratch. */
size_t ck_mg_workspace_bytes(ck_mg_ctx* h, ck_i64 n_data, ck_i64 m);

/* Assemble this rank's tiles of [Sigma ; C^T + z] (ck_mg_assemble) inside ws_dev, which must stay
 * alive until the predictions have been read.  Arguments as ck_joint_cov + ck_cross_cov. */
int ck_mg_joint_cov(ck_mg_ctx* h, const double* xy0_dev, ck_i64 n0, const double* xy1_dev, ck_i64 n1, const double* xyp_dev,
                    ck_i64 m, const double* z_dev, const double* params /*HOST*/, int n_procs, int i_pred, int metric,
                    void* ws_dev, size_t ws_bytes, void* stream);

/* The sweep: right-looking tile Cholesky over all rows of the augmented array with NCCL panel broadcasts
 * (diagonal tile down the process column, panel rows along each process row, panel columns inside each
 * process column) and one-panel look-ahead on the context's panel stream; trailing updates on `stream`.
 * Leaves L in the data tiles, L^-1 c and L^-1 z in the target tiles. */
int ck_mg_potrf(ck_mg_ctx* h, void* stream);

/* pred[c] = (L^-1 c).(L^-1 z), var[c] = c0 - |L^-1 c|^2 for all m targets on EVERY rank (per-rank row
 * partial sums all-gathered and added in rank order: deterministic); *info_dev (device int) = 0 or the
 * order of the first non-positive-definite leading minor, as ck_potrf. */
int ck_mg_potrs_predict(ck_mg_ctx* h, double* pred_dev, double* var_dev, int* info_dev, void* stream);

/* logdet(Sigma) = 2 sum log L_kk over the diagonal tiles of all ranks -> *out_dev on every rank. */
int ck_mg_logdet(ck_mg_ctx* h, double* out_dev, void* stream);

/* Milliseconds of the last assemble / sweep / predict phases on this rank (synchronises on their events). */
int ck_mg_times_ms(ck_mg_ctx* h, double* out3 /*HOST*/);

/* This rank's part of the array (diagnostics: sampled checks of L L^T): local tile (li, lj) = global
 * tile (li P + p, lj Q + q) at local_dev + li*tile*ld + lj*tile. */
int ck_mg_local_factor(ck_mg_ctx* h, double** local_dev /*HOST*/, ck_i64* ld /*HOST*/, ck_i64* local_row_tiles /*HOST*/,
                       ck_i64* local_col_tiles /*HOST*/);

/* ------------------------------------------------------------------------------------------------
 * K3 (INT8 tensor-core path)  FP64-equivalent rank-k updates  C -= A B^T  by fixed-slice error-free
 * splitting (7 balanced base-256 digits per operand, 28 int8 x int8 -> int32 slice products in TMEM,
 * FP64 recombination in the epilogue).  ck_potrf / ck_trsm_lower use these internally for their big
 * trailing updates (same calls the reference makes: scipy cho_factor / cho_solve,
 * src/joint_prediction.py:68-73); they are exported for tests and tools.
 * ---------------------------------------------------------------------------------------------- */

/* Process-wide switches of the INT8 path (negative value = leave unchanged): enabled (default 1, env CK_OZAKI)
 * and the smallest trailing dimension handed to it (default 1024, env CK_OZ_MIN_ROWS; matrices smaller than
 * twice that, or than 2048, are factored by the FP64 DMMA kernel alone).  The workspace size does not depend on
 * these switches. */
int ck_oz_configure(int enabled, ck_i64 min_rows);

/* 1 if ck_potrf / ck_trsm_lower hand the big updates of an n x n system to the INT8 kernel under the current switches. */
int ck_oz_active(ck_i64 n);

/* Profiling aid: 16 clock64 stamps (thread 0) of the phases of every following diagonal-block kernel launch
 * (load, 4 x [32-column elimination, panel, trailing update], inverse assembly, store); NULL = off. */
int ck_potf2_debug_buffer(void* dev_stamps);

/* Profiling aid: when set to a device buffer of 8 x 148 int64 counters, every ck_oz_gemm launch stores per-CTA
 * cycle counts there ([0] MMA-issue thread total, [1] waiting for operands, [2] waiting for TMEM, [4] epilogue
 * waiting, [5] epilogue busy).  NULL switches it off (default). */
int ck_oz_debug_buffer(void* dev_counters);

/* Bytes of one slice buffer for a (rows x k) panel: fmt 0 = "A" operand format (128-row blocks: the
 * rows of C), fmt 1 = "B" operand format (64-row blocks: the columns of C). */
size_t ck_oz_slices_bytes(ck_i64 rows, ck_i64 k, int fmt);
/* Length (doubles) of the per-row scale vector for `rows` rows (rows rounded up to 128). */
ck_i64 ck_oz_scales_len(ck_i64 rows);

/* Split the FP64 panel src (rows x k, row-major, leading dimension ld; k a multiple of 32, <= 1024) into
 * int8 digit slices.  fmt_a_dev / fmt_b_dev (either may be NULL) receive the two operand formats,
 * scales_dev the per-row power-of-two scales. */
int ck_oz_split(const double* src_dev, ck_i64 ld, ck_i64 rows, ck_i64 k, void* fmt_a_dev, void* fmt_b_dev,
                double* scales_dev, void* stream);

/* C (m x n, FP64, row-major) -= A B^T from split operands: A-format slices + scales of the m-row panel,
 * B-format slices + scales of the n-row panel.  lower != 0: only entries with column <= row are updated.
 * max_ctas > 0 caps the persistent grid (one CTA per SM) of THIS launch, so that a caller running a second stream
 * (panel look-ahead) can leave SMs to it; 0 = every SM.  It is a launch argument, not process state: concurrent host
 * threads / streams cannot disturb each other. */
int ck_oz_gemm(const void* a_slices_dev, const double* a_scales_dev, ck_i64 m, const void* b_slices_dev,
               const double* b_scales_dev, ck_i64 n, ck_i64 k, double* c_dev, ck_i64 ldc, int lower, int max_ctas,
               void* stream);

/* HOST-ONLY (no CUDA call): the order in which ck_oz_gemm (tb == 0) / ck_oz_mg_update (tb > 0) visit the 128 x 64 tiles of C
 * for a problem shape -- what the kernel's dynamic scheduler hands out.  Writes up to `cap` (row block I, column block j)
 * pairs to ij_out and, per tile, the largest column its first row may update (-1: unmasked) to col_limits_out (either may
 * be NULL); *nvirt_out = virtual tiles including the skipped ones.  Returns the number of tiles visited, -1 on bad arguments. */
ck_i64 ck_oz_tile_order(ck_i64 m, ck_i64 n, int lower, ck_i64 tb, ck_i64 row_tile0, ck_i64 row_tile_step, ck_i64 col_tile0,
                        ck_i64 col_tile_step, int* ij_out /*HOST*/, ck_i64 cap, ck_i64* nvirt_out /*HOST*/,
                        ck_i64* col_limits_out /*HOST*/);

/* The same product with the masking contract of ck_mg_update: C is the local part of a 2-D block-cyclic matrix (square
 * tiles of tb elements, local tile (li, lj) = global tile (row_tile0 + li row_tile_step, col_tile0 + lj col_tile_step));
 * tiles with J > I are skipped, tiles with J == I keep their lower triangle.  m, n whole tiles. */
int ck_oz_mg_update(const void* a_slices_dev, const double* a_scales_dev, ck_i64 m, const void* b_slices_dev,
                    const double* b_scales_dev, ck_i64 n, ck_i64 k, double* c_dev, ck_i64 ldc, ck_i64 tb, ck_i64 row_tile0,
                    ck_i64 row_tile_step, ck_i64 col_tile0, ck_i64 col_tile_step, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COKRIG_H */

// ck_matern.cu -- K1: fused distance + Matern (cross-)covariance assembly for sm_100a.
//
// One CTA produces one 64 x 64 tile of a block: the two coordinate tiles are staged once in shared
// memory (prepared per point; for haversine covariance blocks: the half-angle sines / cosines and cos(lat), so that a
// pair needs no sine evaluation at all -- ck_dist_haversine_pre), each thread evaluates 16 entries
// in registers (independent chains -> ILP on the FP64 pipe), and rows are written with 256-byte
// coalesced warp stores.  The optional mirrored copy (symmetric blocks, C01 -> C10) goes through a
// padded shared-memory transpose so that it is written coalesced as well: the matrix is written
// exactly once and never re-read (algorithmic HBM traffic = 8 N^2 bytes).
#include <stdarg.h>
#include "ck_common.cuh"

static thread_local char g_err[512] = "";
void ck_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* ck_last_error(void) { return g_err; }

#include <atomic>
static std::atomic<long long> g_launches{0};
void ck_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long ck_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int ck_version(void) { return 100; }

int ck_unpack_params(const double* v, int n_procs, CkParams* p) {
  if (!v || !p) { ck_set_error("params is NULL"); return CK_ERR_ARG; }
  memset(p, 0, sizeof(*p));
  p->n_procs = n_procs;
  if (n_procs == 1) {
    p->sigma[0] = v[0]; p->nu[0][0] = v[1]; p->len_scale[0][0] = v[2]; p->nugget[0] = v[3];
    p->sigma_prod = v[0];
    return CK_OK;
  }
  if (n_procs == 2) {
    p->sigma[0] = v[0]; p->sigma[1] = v[1];
    p->nu[0][0] = v[2]; p->nu[0][1] = p->nu[1][0] = v[3]; p->nu[1][1] = v[4];
    p->len_scale[0][0] = v[5]; p->len_scale[0][1] = p->len_scale[1][0] = v[6]; p->len_scale[1][1] = v[7];
    p->nugget[0] = v[8]; p->nugget[1] = v[9];
    p->rho01 = v[10];
    p->sigma_prod = v[0] * v[1];
    return CK_OK;
  }
  ck_set_error("n_procs=%d unsupported (1 or 2)", n_procs);
  return CK_ERR_UNSUPPORTED;
}

int ck_block_matern(const CkParams& p, int i, int j, int use_nugget, CkMatern* out) {
  if (i < 0 || j < 0 || i >= p.n_procs || j >= p.n_procs) { ck_set_error("process index out of range"); return CK_ERR_ARG; }
  int rc;
  if (i == j) rc = ck_matern_setup(out, p.sigma[i] * p.sigma[i], p.nu[i][i], p.len_scale[i][i], use_nugget ? p.nugget[i] : 0.0);
  else rc = ck_matern_setup(out, p.rho01 * p.sigma_prod, p.nu[i][j], p.len_scale[i][j], 0.0);
  if (rc) { ck_set_error("invalid Matern parameters for block (%d,%d): nu=%g len_scale=%g", i, j, p.nu[i][j], p.len_scale[i][j]); return CK_ERR_ARG; }
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
constexpr int TILE = 64;
constexpr int K1_THREADS = 256;

// per-point record and pair distance of one kernel variant: covariance blocks on the sphere use the precomputed
// half-angle form, everything else (Euclid; distance output in the reference's operation order) the plain point
template <int METRIC, int VALUE>
struct K1Point {
  using type = CkPoint;
  static __device__ __forceinline__ type prepare(double a, double b) { return ck_prepare_point(METRIC, a, b); }
  static __device__ __forceinline__ double dist(const type& p, const type& q) {
    return VALUE ? ck_dist_fast<METRIC>(p, q) : ck_dist<METRIC>(p, q);
  }
};
template <>
struct K1Point<CK_METRIC_HAVERSINE, 1> {
  using type = CkPointH;
  static __device__ __forceinline__ type prepare(double a, double b) { return ck_prepare_point_h(a, b); }
  static __device__ __forceinline__ double dist(const type& p, const type& q) { return ck_dist_haversine_pre(p, q); }
};


// ------------------------------------------------------------------------------------------------
// Closed-form covariance blocks: the 16 entries of a thread are evaluated in two groups of eight with WARP-UNIFORM fast
// paths.  Round-2 ncu (profiles/r02_k1_ncu_full_summary.txt) showed the branch-free scalar path issue-bound: 186
// instructions per entry of which only 71 on the FP64 pipe -- the rest FSEL pairs of the branch-free selects, register
// moves and one uniform-datapath coefficient load per polynomial term per entry.  Here (a) every polynomial is evaluated
// coefficient-major over the eight entries, so a coefficient is fetched once per eight Horner steps, and (b) when a vote
// shows that every lane's eight entries are in the common range (haversine a <= 1/4, i.e. pairs closer than ~6 670 km;
// 0 < x < the kv underflow cut-off) the selects disappear; any lane outside it sends the whole warp through the general
// scalar functions of ck_math.cuh (diagonal tiles, coincident points, antipodal pairs).  Same arithmetic per entry in
// both paths: ck_fast_sqrt, the same polynomials, the same exponent insertion.
// ------------------------------------------------------------------------------------------------
constexpr unsigned K1_FULL = 0xffffffffu;

// d[e] = 2 R asin(sqrt(a[e])) for eight values of the haversine argument
__device__ __forceinline__ void k1_asin8(const double (&a)[8], double (&d)[8]) {
  bool small = true;
#pragma unroll
  for (int e = 0; e < 8; ++e) small &= a[e] <= 0.25;
  if (__all_sync(K1_FULL, small)) {  // asin(sqrt(a)) = sqrt(a) (1 + a R(a))
    double r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = CK_COEF(asin_r, 12);
#pragma unroll
    for (int i = 11; i >= 0; --i) {
      const double c = CK_COEF(asin_r, i);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = fma(r[e], a[e], c);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = (ck_fast_sqrt(a[e]) * (2.0 * CK_EARTH_RADIUS)) * fma(a[e], r[e], 1.0);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = ck_fast_asin_sqrt(a[e]) * (2.0 * CK_EARTH_RADIUS);
  }
}

// out[e] = sigma^2 rho(h[e]) (+ nugget where h == 0) for eight distances, closed-form orders
template <int MODE>
__device__ __forceinline__ void k1_matern8(const CkMatern& P, const double (&h)[8], double (&out)[8]) {
  const double xlim = P.x_cut < 705.0 ? P.x_cut : 705.0;
  double x[8];
  bool plain = true;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    x[e] = fabs(h[e]) * P.xscale;
    plain &= (x[e] > 0.0) & (x[e] < xlim);
  }
  if (__all_sync(K1_FULL, plain)) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52, see ck_fast_exp_neg
    double t[8], r[8], q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      t[e] = fma(x[e], -1.4426950408889634074, magic);
      const double nf = t[e] - magic;
      r[e] = fma(nf, -6.93147180369123816490e-01, -x[e]);
      r[e] = fma(nf, -1.90821492927058770002e-10, r[e]);
      q[e] = CK_COEF(exp_e, 9);
    }
#pragma unroll
    for (int i = 8; i >= 0; --i) {
      const double c = CK_COEF(exp_e, i);
#pragma unroll
      for (int e = 0; e < 8; ++e) q[e] = fma(q[e], r[e], c);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const double p = fma(r[e] * r[e], q[e], r[e]) + 1.0;
      const double ex = __hiloint2double(__double2hiint(p) + (__double2loint(t[e]) << 20), __double2loint(p));
      double poly;
      if (MODE == CK_NU_HALF) poly = 1.0;
      else if (MODE == CK_NU_3HALF) poly = 1.0 + x[e];
      else if (MODE == CK_NU_5HALF) poly = fma(x[e], fma(x[e], 1.0 / 3.0, 1.0), 1.0);
      else poly = fma(x[e], fma(x[e], fma(x[e], 1.0 / 15.0, 0.4), 1.0), 1.0);
      out[e] = P.scale * (poly * ex);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = ck_matern_cov_fast<MODE>(P, h[e]);
  }
}

// eight pair distances of one column point q against the thread's eight row points
template <int METRIC>
struct K1Dist8;
template <>
struct K1Dist8<CK_METRIC_HAVERSINE> {
  static __device__ __forceinline__ void run(const CkPointH* pr, int ty, const CkPointH& q, double (&h)[8]) {
    double a[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const CkPointH p = pr[ty + 8 * r];
      const double s0 = __dsub_rn(__dmul_rn(p.sa, q.ca), __dmul_rn(p.ca, q.sa));
      const double s1 = __dsub_rn(__dmul_rn(p.sb, q.cb), __dmul_rn(p.cb, q.sb));
      a[r] = fma(p.c * q.c, s1 * s1, s0 * s0);
    }
    k1_asin8(a, h);
  }
};
template <>
struct K1Dist8<CK_METRIC_EUCLID> {
  static __device__ __forceinline__ void run(const CkPoint* pr, int ty, const CkPoint& q, double (&h)[8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) h[r] = ck_dist_euclid_fast(pr[ty + 8 * r], q);
  }
};

// kernel argument that carries the generic-nu correlation table only for the instantiations that use it
template <bool TAB>
struct K1TabArg {
  CkMaternTable T;
};
template <>
struct K1TabArg<false> {};

// VALUE: 0 = distance only, 1 = covariance
template <int METRIC, int MODE, int VALUE>
__global__ void __launch_bounds__(K1_THREADS, (METRIC == CK_METRIC_HAVERSINE && VALUE) ? 3 : 4)
    ck_block_kernel(const double* __restrict__ xy1, long long n1,
                                                              const double* __restrict__ xy2, long long n2,
                                                              CkMatern P, double* __restrict__ out, long long ld,
                                                              double* __restrict__ out_t, long long ld_t,
                                                              int symmetric,
                                                              const __grid_constant__ K1TabArg<(MODE == CK_NU_GENERIC && VALUE != 0)> tab) {
  const long long bi = blockIdx.y, bj = blockIdx.x;
  if (symmetric && bj < bi) return;  // mirrored from the upper tile
  using PT = K1Point<METRIC, VALUE>;
  __shared__ typename PT::type pr[TILE], pc[TILE];
  __shared__ double tr[TILE][TILE + 1];
  const int t = threadIdx.x;
  const long long r0 = bi * TILE, c0 = bj * TILE;
  if (t < TILE) {  // points past the end of the block are staged as copies of its last point: never stored, but every
    const long long r = (r0 + t < n1) ? r0 + t : n1 - 1;  // lane computes on valid data and takes part in the warp votes
    pr[t] = PT::prepare(xy1[2 * r], xy1[2 * r + 1]);
  } else if (t < 2 * TILE) {
    const long long c = (c0 + (t - TILE) < n2) ? c0 + (t - TILE) : n2 - 1;
    pc[t - TILE] = PT::prepare(xy2[2 * c], xy2[2 * c + 1]);
  }
  constexpr bool TAB = (MODE == CK_NU_GENERIC && VALUE != 0);
  __shared__ double stab[TAB ? CK_TAB_NC * CK_TAB_NSEG : 1];  // generic nu: the block's correlation table (8.4 KB)
  if constexpr (TAB) {
    const double* src = &tab.T.c[0][0];
    for (int i = t; i < CK_TAB_NC * CK_TAB_NSEG; i += K1_THREADS) stab[i] = src[i];
  }
  __syncthreads();
  const int tx = t & 31, ty = t >> 5;  // 8 warps; warp `ty` owns rows ty, ty+8, ...
  const bool mirror = (out_t != nullptr) && !(symmetric && bi == bj);
  if constexpr (VALUE != 0 && MODE != CK_NU_GENERIC) {
    // closed-form orders: two groups of eight entries (one column point each) through the warp-uniform fast paths;
    // each group is stored as soon as it is finished (eight results live, not sixteen)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int lc = tx + 32 * c;
      double h[8], o[8];
      K1Dist8<METRIC>::run(pr, ty, pc[lc], h);
      k1_matern8<MODE>(P, h, o);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int lr = ty + 8 * r;
        if (r0 + lr < n1 && c0 + lc < n2) out[(r0 + lr) * ld + (c0 + lc)] = o[r];
        if (mirror) tr[lr][lc] = o[r];
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int lr = ty + 8 * r;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int lc = tx + 32 * c;
        const double h = PT::dist(pr[lr], pc[lc]);
        double val = h;  // distance output: reference operation order
        if constexpr (TAB) {  // generic order: piecewise Chebyshev table of rho(x) e^x (ck_math.cuh)
          val = P.scale * ck_matern_corr_tab(P, stab, h);
          if (h == 0.0) val += P.nugget;
        }
        if (r0 + lr < n1 && c0 + lc < n2) out[(r0 + lr) * ld + (c0 + lc)] = val;
        if (mirror) tr[lr][lc] = val;
      }
    }
  }
  if (mirror) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int lc = ty + 8 * r;  // row of the transposed tile = column of this tile
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int lr = tx + 32 * c;
        if (c0 + lc < n2 && r0 + lr < n1) out_t[(c0 + lc) * ld_t + (r0 + lr)] = tr[lr][lc];
      }
    }
  }
}

template <int METRIC, int VALUE>
static void launch_mode(const CkMatern& P, dim3 grid, cudaStream_t st, const double* xy1, long long n1, const double* xy2,
                        long long n2, double* out, long long ld, double* out_t, long long ld_t, int symmetric) {
#define CK_K1(MODE) \
  ck_block_kernel<METRIC, MODE, VALUE><<<grid, K1_THREADS, 0, st>>>(xy1, n1, xy2, n2, P, out, ld, out_t, ld_t, symmetric, K1TabArg<false>())
  if (!VALUE) { CK_K1(CK_NU_HALF); return; }
  switch (P.mode) {
    case CK_NU_HALF: CK_K1(CK_NU_HALF); break;
    case CK_NU_3HALF: CK_K1(CK_NU_3HALF); break;
    case CK_NU_5HALF: CK_K1(CK_NU_5HALF); break;
    case CK_NU_7HALF: CK_K1(CK_NU_7HALF); break;
    default: {
      if constexpr (VALUE != 0) {
        static thread_local K1TabArg<true> targ;  // 8.4 KB, copied into the launch's parameter buffer by the <<<>>> call
        ck_matern_table_setup(P, &targ.T);
        ck_block_kernel<METRIC, CK_NU_GENERIC, VALUE><<<grid, K1_THREADS, 0, st>>>(xy1, n1, xy2, n2, P, out, ld, out_t, ld_t,
                                                                                 symmetric, targ);
      }
    } break;
  }
#undef CK_K1
}

int ck_block_launch(const double* xy1, ck_i64 n1, const double* xy2, ck_i64 n2, int metric, const CkMatern& P, int value,
                        double* out, ck_i64 ld, double* out_t, ck_i64 ld_t, int symmetric, cudaStream_t st) {
  CK_REQUIRE(n1 >= 0 && n2 >= 0, "negative size");
  if (n1 == 0 || n2 == 0) return CK_OK;
  CK_REQUIRE(xy1 && xy2 && out, "null pointer");
  CK_REQUIRE(ld >= n2, "ld (%lld) < n2 (%lld)", (long long)ld, (long long)n2);
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  if (symmetric) {
    CK_REQUIRE(n1 == n2, "symmetric block needs n1 == n2");
    out_t = out;
    ld_t = ld;
  } else if (out_t) {
    CK_REQUIRE(ld_t >= n1, "ld_t (%lld) < n1 (%lld)", (long long)ld_t, (long long)n1);
  }
  const long long gx = (n2 + TILE - 1) / TILE, gy = (n1 + TILE - 1) / TILE;
  CK_REQUIRE(gy <= 65535, "n1 too large for one launch (%lld rows)", (long long)n1);
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (metric == CK_METRIC_HAVERSINE) {
    if (value) launch_mode<CK_METRIC_HAVERSINE, 1>(P, grid, st, xy1, n1, xy2, n2, out, ld, out_t, ld_t, symmetric);
    else launch_mode<CK_METRIC_HAVERSINE, 0>(P, grid, st, xy1, n1, xy2, n2, out, ld, out_t, ld_t, symmetric);
  } else {
    if (value) launch_mode<CK_METRIC_EUCLID, 1>(P, grid, st, xy1, n1, xy2, n2, out, ld, out_t, ld_t, symmetric);
    else launch_mode<CK_METRIC_EUCLID, 0>(P, grid, st, xy1, n1, xy2, n2, out, ld, out_t, ld_t, symmetric);
  }
  CK_LAUNCH_CHECK();
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) ck_eval_kernel(const double* __restrict__ h, long long n, CkMatern P,
                                                      double* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    out[k] = ck_matern_cov<MODE>(P, h[k]);
}

extern "C" int ck_matern_eval(const double* h, ck_i64 n, double scale, double nu, double len_scale, double nugget,
                              double* out, void* stream) {
  CK_REQUIRE(n >= 0, "negative size");
  if (n == 0) return CK_OK;
  CK_REQUIRE(h && out, "null pointer");
  CkMatern P;
  CK_REQUIRE(ck_matern_setup(&P, scale, nu, len_scale, nugget) == 0, "invalid Matern parameters nu=%g len_scale=%g", nu, len_scale);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = ck_stream(stream);
  switch (P.mode) {
    case CK_NU_HALF: ck_eval_kernel<CK_NU_HALF><<<(unsigned)blocks, 256, 0, st>>>(h, n, P, out); break;
    case CK_NU_3HALF: ck_eval_kernel<CK_NU_3HALF><<<(unsigned)blocks, 256, 0, st>>>(h, n, P, out); break;
    case CK_NU_5HALF: ck_eval_kernel<CK_NU_5HALF><<<(unsigned)blocks, 256, 0, st>>>(h, n, P, out); break;
    case CK_NU_7HALF: ck_eval_kernel<CK_NU_7HALF><<<(unsigned)blocks, 256, 0, st>>>(h, n, P, out); break;
    default: ck_eval_kernel<CK_NU_GENERIC><<<(unsigned)blocks, 256, 0, st>>>(h, n, P, out); break;
  }
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_distance_block(const double* xy1, ck_i64 n1, const double* xy2, ck_i64 n2, int metric, double* out,
                                 ck_i64 ld, void* stream) {
  CkMatern P;
  ck_matern_setup(&P, 1.0, 0.5, 1.0, 0.0);
  return ck_block_launch(xy1, n1, xy2, n2, metric, P, 0, out, ld, nullptr, 0, 0, ck_stream(stream));
}

extern "C" int ck_matern_block(const double* xy1, ck_i64 n1, const double* xy2, ck_i64 n2, int metric, double scale,
                               double nu, double len_scale, double nugget, double* out, ck_i64 ld, double* out_t,
                               ck_i64 ld_t, int symmetric, void* stream) {
  CkMatern P;
  CK_REQUIRE(ck_matern_setup(&P, scale, nu, len_scale, nugget) == 0, "invalid Matern parameters nu=%g len_scale=%g", nu, len_scale);
  return ck_block_launch(xy1, n1, xy2, n2, metric, P, 1, out, ld, out_t, ld_t, symmetric, ck_stream(stream));
}

extern "C" int ck_joint_cov(const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* params, int n_procs,
                            int metric, double* sigma, ck_i64 ld, void* stream) {
  CkParams p;
  int rc = ck_unpack_params(params, n_procs, &p);
  if (rc) return rc;
  if (n_procs == 1) n1 = 0;
  CK_REQUIRE(n0 >= 0 && n1 >= 0, "negative size");
  CK_REQUIRE(ld >= n0 + n1, "ld (%lld) < N (%lld)", (long long)ld, (long long)(n0 + n1));
  cudaStream_t st = ck_stream(stream);
  CkMatern P;
  if ((rc = ck_block_matern(p, 0, 0, 1, &P))) return rc;
  if ((rc = ck_block_launch(xy0, n0, xy0, n0, metric, P, 1, sigma, ld, nullptr, 0, 1, st))) return rc;
  if (n_procs == 2) {
    if ((rc = ck_block_matern(p, 1, 1, 1, &P))) return rc;
    if ((rc = ck_block_launch(xy1, n1, xy1, n1, metric, P, 1, sigma + n0 * ld + n0, ld, nullptr, 0, 1, st))) return rc;
    if ((rc = ck_block_matern(p, 0, 1, 0, &P))) return rc;
    if ((rc = ck_block_launch(xy0, n0, xy1, n1, metric, P, 1, sigma + n0, ld, sigma + n0 * ld, ld, 0, st))) return rc;
  }
  return CK_OK;
}

extern "C" int ck_cross_cov(const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* xyp, ck_i64 m,
                            const double* params, int n_procs, int i_pred, int metric, double* cpd, ck_i64 ld,
                            void* stream) {
  CkParams p;
  int rc = ck_unpack_params(params, n_procs, &p);
  if (rc) return rc;
  if (n_procs == 1) n1 = 0;
  CK_REQUIRE(i_pred >= 0 && i_pred < n_procs, "i_pred out of range");
  CK_REQUIRE(ld >= n0 + n1, "ld (%lld) < N (%lld)", (long long)ld, (long long)(n0 + n1));
  cudaStream_t st = ck_stream(stream);
  const double* xy[2] = {xy0, xy1};
  const ck_i64 nn[2] = {n0, n1};
  ck_i64 off = 0;
  for (int j = 0; j < n_procs; ++j) {
    CkMatern P;
    if ((rc = ck_block_matern(p, i_pred, j, 1, &P))) return rc;
    if ((rc = ck_block_launch(xyp, m, xy[j], nn[j], metric, P, 1, cpd + off, ld, nullptr, 0, 0, st))) return rc;
    off += nn[j];
  }
  return CK_OK;
}

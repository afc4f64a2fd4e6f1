// ck_chol.cu -- K3: blocked FP64 Cholesky, blocked triangular solves and the simple-cokriging
// prediction epilogue for sm_100a.
//
// Structure (right-looking, panel width CK_NB = 128, matrix row-major, lower triangle):
//   for each diagonal block k:
//     ck_potf2_inv_kernel   one CTA, 32-column panels (warp-level elimination + DMMA updates): L_kk in place and
//                           X_k = L_kk^{-1} into the workspace (kept for the solves)
//     ck_gemm_nt_kernel     panel  A[k+1:, k] <- A[k+1:, k] X_k^T                (TRSM as GEMM, in place)
//     ck_gemm_nt_kernel     trailing A[k+1:, k+1:] -= P P^T on lower tiles only (DSYRK)
// All O(N^3) work is in ck_gemm_nt_kernel: C (-)= A B^T with both operands K-contiguous, computed by
// FP64 tensor-core DMMA (mma.sync m8n8k4 f64 -- tcgen05 has no f64 kind), operands staged through a
// 3-stage cp.async shared-memory pipeline with a conflict-free padded layout, accumulators in
// registers (initialised from C so the update needs no separate epilogue read).
// The triangular solve with many right-hand sides keeps the RHS TARGET-MAJOR (one right-hand side
// per row) so that it is the same NT GEMM:  V[:, k] = R[:, k] X_k^T ;  R[:, k+1:] -= V[:, k] L[k+1:, k]^T.
#include <stdlib.h>
#include "ck_common.cuh"

// ------------------------------------------------------------------------------------------------
// DMMA GEMM   C = A B^T  (mode 0)   or   C = C - A B^T  (mode 1)
// ------------------------------------------------------------------------------------------------
constexpr int G_BK = 16;            // k-depth per pipeline stage
constexpr int G_LDS = G_BK + 4;     // padded smem row stride (doubles): stride % 16 == 4 -> conflict-free fragments
constexpr int G_STAGES = 3;
constexpr int G_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct GemmArgs {
  const double* A; long long lda;  // M x K
  const double* B; long long ldb;  // N x K
  double* C; long long ldc;        // M x N
  long long M, N, K;
  int mode;        // 0: C = A B^T, 1: C = C - A B^T
  int lower_only;  // 1: C square, triangular grid over the tiles intersecting the lower triangle;
                   // 2: C rectangular (M >= N, diagonal at the top-left), 2-D grid, tiles above the diagonal exit;
                   // 3: C is the local part of a 2-D block-cyclic matrix (square tiles of `tb` elements): local
                   //    tile (li, lj) is global tile (gi0 + li*gis, gj0 + lj*gjs); tiles with J > I exit, tiles
                   //    with J == I keep the lower triangle of the tile;
                   // in all cases entries above the (global) diagonal are not stored
  int tb = 0;
  long long gi0 = 0, gis = 1, gj0 = 0, gjs = 1;
};

// CTA tile (32*WM) x (32*WN); each of the 8 warps owns a 32 x 32 sub-tile = 4 x 4 DMMA tiles.
// VEC16: operands 16-byte aligned with even leading dimensions (vector copies / stores); KFULL: additionally
// K % G_BK == 0, so every k-tile is full and the branch-free fast loader is used throughout.
template <int WM, int WN, bool VEC16, bool KFULL>
__global__ void __launch_bounds__(G_THREADS, 2) ck_gemm_nt_kernel(GemmArgs g) {
  static_assert(VEC16 || !KFULL, "KFULL needs VEC16");
  constexpr int BM = 32 * WM, BN = 32 * WN;
  constexpr int STAGE_ELEMS = (BM + BN) * G_LDS;
  extern __shared__ __align__(16) double smem[];
  long long tm, tn;
  if (g.lower_only == 1) {
    // linear tile id -> (row tile tm, col tile tn) over the lower block-triangle; row tm has RB*(tm+1) tiles
    constexpr int RB = (BM >= BN) ? BM / BN : 1;
    const long long t = blockIdx.x;
    long long r = (long long)((sqrt(8.0 * (double)t / RB + 1.0) - 1.0) * 0.5);
    while (RB * (r + 1) * (r + 2) / 2 <= t) ++r;
    while (RB * r * (r + 1) / 2 > t) --r;
    tm = r;
    tn = t - RB * r * (r + 1) / 2;
  } else {
    tm = blockIdx.y;
    tn = blockIdx.x;
    if (g.lower_only == 2 && tn * BN > tm * BM + BM - 1) return;  // rectangular C: tile entirely above the diagonal
  }
  const long long m0 = tm * BM, n0 = tn * BN;
  // store mask: entry (r, c) of C is kept iff !masked or (c - coff) <= (r - roff)
  bool masked = g.lower_only != 0;
  long long roff = 0, coff = 0;
  if (g.lower_only == 3) {
    const long long li = m0 / g.tb, lj = n0 / g.tb;
    const long long I = g.gi0 + li * g.gis, J = g.gj0 + lj * g.gjs;
    if (J > I) return;
    roff = li * g.tb;
    coff = lj * g.tb;
    masked = (J == I);
    if (masked && (n0 - coff) > (m0 - roff) + BM - 1) return;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int g4 = lane >> 2, t4 = lane & 3;

  // Operand staging.  Fast path (16-byte aligned operands, full k-tile): every thread owns the same 16-byte
  // column chunk `lch` of rows lrow + 32 i of the A tile (i < BM/32) and of the B tile (i < BN/32).  All index
  // arithmetic is done ONCE here; per stage a thread bumps two running pointers and issues its copies, so the
  // non-DMMA phase between the barrier and the first DMMA of a stage is a few dozen instructions.  (The two
  // co-resident CTAs share the DMMA pipe fairly and therefore phase-lock: the pipe idles for the length of
  // that phase every stage -- 83 % pipe-active with per-stage index math, profiles/r01b_gemm_*.)
  // Rows past M / N are simply not copied: whatever the buffer holds there only reaches accumulators of C
  // rows / columns that are never stored.  A partial last k-tile (K % 16 != 0) and unaligned operands take
  // the element-wise zero-filling path.
  constexpr int CH = G_BK / 2;                    // 16-byte chunks per row
  constexpr int RPI = G_THREADS / CH;             // rows covered per pass (32)
  constexpr int NA = BM / RPI, NB = BN / RPI;     // copies per thread per stage (A, B)
  static_assert(G_THREADS % CH == 0 && BM % RPI == 0 && BN % RPI == 0, "loader shape");
  const int lrow = tid / CH, lch = tid % CH;
  const double* pa = g.A + (m0 + lrow) * g.lda + lch * 2;  // running pointers: advanced by G_BK per issued stage
  const double* pb = g.B + (n0 + lrow) * g.ldb + lch * 2;
  const long long sa = (long long)RPI * g.lda, sb = (long long)RPI * g.ldb;
  unsigned vmask = 0;  // bit i: A row valid (i < NA); bit NA + i: B row valid
#pragma unroll
  for (int i = 0; i < NA; ++i) vmask |= (m0 + lrow + i * RPI < g.M) ? (1u << i) : 0u;
#pragma unroll
  for (int i = 0; i < NB; ++i) vmask |= (n0 + lrow + i * RPI < g.N) ? (1u << (NA + i)) : 0u;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned sdst0 = sbase + (unsigned)((lrow * G_LDS + lch * 2) * (int)sizeof(double));
  constexpr unsigned STAGE_BYTES = (unsigned)(STAGE_ELEMS * sizeof(double));

  // fast loader: `on` = 0 turns every copy off (past the last k-tile); no branches
  auto load_fast = [&](int stage, unsigned on) {
    const unsigned dst = sdst0 + (unsigned)stage * STAGE_BYTES;
    const unsigned vm = on ? vmask : 0u;
    const double* qa = pa;
    const double* qb = pb;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.cg.shared.global [%0], [%1], 16;\n}\n" ::"r"(
                       dst + (unsigned)(i * RPI * G_LDS * (int)sizeof(double))),
                   "l"(qa), "r"(vm & (1u << i)));
      qa += sa;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.cg.shared.global [%0], [%1], 16;\n}\n" ::"r"(
                       dst + (unsigned)((BM + i * RPI) * G_LDS * (int)sizeof(double))),
                   "l"(qb), "r"(vm & (1u << (NA + i))));
      qb += sb;
    }
    pa += G_BK;
    pb += G_BK;
  };
  // generic loader: element-wise, zero-fills rows past M / N and the k-tail
  auto load_slow = [&](int stage, long long k0) {
    double* base = smem + stage * STAGE_ELEMS;
#pragma unroll 1
    for (int idx = tid; idx < (BM + BN) * G_BK; idx += G_THREADS) {
      const int row = idx / G_BK, ch = idx % G_BK;
      const long long gk = k0 + ch;
      const bool isA = row < BM;
      const long long gr = isA ? m0 + row : n0 + (row - BM);
      const long long lim = isA ? g.M : g.N;
      const int bytes = (gr < lim && gk < g.K) ? 8 : 0;
      const double* src = (isA ? g.A + (bytes ? gr * g.lda : 0) : g.B + (bytes ? gr * g.ldb : 0)) + (bytes ? gk : 0);
      cp_async8(base + row * G_LDS + ch, src, bytes);
    }
  };

  const long long KT = (g.K + G_BK - 1) / G_BK;
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (KFULL) load_fast(s, s < KT ? 1u : 0u);
    else if (s < KT) load_slow(s, (long long)s * G_BK);
    cp_async_commit();
  }

  // accumulators: acc[mi][ni][0..1] <-> C[m0 + wm*32 + mi*8 + g4][n0 + wn*32 + ni*8 + 2*t4 + {0,1}]
  // mode 1 accumulates  acc = -C + A B^T  and stores -acc: bitwise equal to C - A B^T (negation is exact and
  // round-to-nearest is sign-symmetric) without touching the operand fragments in the main loop.
  double acc[4][4][2];
  const long long crow0 = m0 + wm * 32 + g4, ccol0 = n0 + wn * 32 + 2 * t4;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      acc[mi][ni][0] = 0.0;
      acc[mi][ni][1] = 0.0;
      if (g.mode == 1) {
        const long long r = crow0 + mi * 8, c = ccol0 + ni * 8;
        if (r < g.M) {
          const double* p = g.C + r * g.ldc + c;
          if (VEC16 && c + 1 < g.N) {
            const double2 v = *reinterpret_cast<const double2*>(p);
            acc[mi][ni][0] = -v.x;
            acc[mi][ni][1] = -v.y;
          } else {
            if (c < g.N) acc[mi][ni][0] = -p[0];
            if (c + 1 < g.N) acc[mi][ni][1] = -p[1];
          }
        }
      }
    }

  const unsigned a_off = (unsigned)(((wm * 32 + g4) * G_LDS + t4) * (int)sizeof(double));
  const unsigned b_off = (unsigned)(((BM + wn * 32 + g4) * G_LDS + t4) * (int)sizeof(double));
  const char* sm_c = reinterpret_cast<const char*>(smem);
  auto mma_step = [&](const double* As, const double* Bs, int kk) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = As[mi * 8 * G_LDS + kk];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * G_LDS + kk];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  };
  int rd = 0, wr = G_STAGES - 1;
  for (long long kt = 0; kt < KT; ++kt) {
    cp_async_wait<G_STAGES - 2>();
    __syncthreads();  // stage `rd` has landed for everybody; everybody is done with stage `wr` (read at kt - 1)
    const double* As = reinterpret_cast<const double*>(sm_c + (unsigned)rd * STAGE_BYTES + a_off);
    const double* Bs = reinterpret_cast<const double*>(sm_c + (unsigned)rd * STAGE_BYTES + b_off);
    const long long nk = kt + G_STAGES - 1;
    if (KFULL) {
      // the refill of stage `wr` is issued in the shadow of the first DMMA group, not between the barrier and it
      mma_step(As, Bs, 0);
      load_fast(wr, nk < KT ? 1u : 0u);
      cp_async_commit();
#pragma unroll
      for (int kk = 4; kk < G_BK; kk += 4) mma_step(As, Bs, kk);
    } else {
      if (nk < KT) load_slow(wr, nk * G_BK);
      cp_async_commit();
#pragma unroll
      for (int kk = 0; kk < G_BK; kk += 4) mma_step(As, Bs, kk);
    }
    rd = (rd + 1 == G_STAGES) ? 0 : rd + 1;
    wr = (wr + 1 == G_STAGES) ? 0 : wr + 1;
  }
  cp_async_wait<0>();
  __syncthreads();  // all operand reads of this CTA are complete before any store (in-place TRSM safety)

  const double sgn = (g.mode == 1) ? -1.0 : 1.0;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const long long r = crow0 + mi * 8, c = ccol0 + ni * 8;
      if (r >= g.M) continue;
      double* p = g.C + r * g.ldc + c;
      const bool ok0 = c < g.N && (!masked || c - coff <= r - roff);
      const bool ok1 = c + 1 < g.N && (!masked || c + 1 - coff <= r - roff);
      const double v0 = sgn * acc[mi][ni][0], v1 = sgn * acc[mi][ni][1];
      if (VEC16 && ok0 && ok1) {
        *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
      } else {
        if (ok0) p[0] = v0;
        if (ok1) p[1] = v1;
      }
    }
}

template <int WM, int WN>
static int gemm_launch(const GemmArgs& g, cudaStream_t st) {
  constexpr int BM = 32 * WM, BN = 32 * WN;
  constexpr size_t SMEM = (size_t)G_STAGES * (BM + BN) * G_LDS * sizeof(double);
  if (g.M <= 0 || g.N <= 0) return CK_OK;
  const bool vec = ((((uintptr_t)g.A | (uintptr_t)g.B | (uintptr_t)g.C) & 15) == 0) && !((g.lda | g.ldb | g.ldc) & 1);
  const bool kfull = vec && (g.K % G_BK == 0);
  static CkPerDevice attr_a, attr_b, attr_c;  // one set per <WM, WN> instantiation
  CK_SET_SMEM_ONCE(attr_a, (ck_gemm_nt_kernel<WM, WN, true, true>), SMEM);
  CK_SET_SMEM_ONCE(attr_b, (ck_gemm_nt_kernel<WM, WN, true, false>), SMEM);
  CK_SET_SMEM_ONCE(attr_c, (ck_gemm_nt_kernel<WM, WN, false, false>), SMEM);
  const long long tm = (g.M + BM - 1) / BM, tn = (g.N + BN - 1) / BN;
  dim3 grid;
  if (g.lower_only == 1) {
    constexpr int RB = (BM >= BN) ? BM / BN : 1;
    CK_REQUIRE(BM >= BN, "lower_only needs BM >= BN");
    long long tiles = RB * tm * (tm + 1) / 2;
    // the last row tile may extend past N in columns: tiles with n0 >= N are launched but exit at the store guards
    grid = dim3((unsigned)tiles, 1, 1);
  } else {
    CK_REQUIRE(tm <= 65535, "too many row tiles");
    grid = dim3((unsigned)tn, (unsigned)tm, 1);
  }
  if (kfull) ck_gemm_nt_kernel<WM, WN, true, true><<<grid, G_THREADS, SMEM, st>>>(g);
  else if (vec) ck_gemm_nt_kernel<WM, WN, true, false><<<grid, G_THREADS, SMEM, st>>>(g);
  else ck_gemm_nt_kernel<WM, WN, false, false><<<grid, G_THREADS, SMEM, st>>>(g);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
// Diagonal block: L_kk = chol(A_kk) and X_k = L_kk^{-1} (128 x 128), one CTA of 256 threads, the block in
// shared memory.  The elimination is blocked in 32-column panels so that the latency-bound part (one
// dependent pivot per column) runs inside ONE warp with register-resident rows and warp shuffles --
// no CTA barrier per column -- and everything else is FP64 DMMA on shared-memory operands:
//   per 32-panel jb:  warp 0   D = L_d L_d^T by LDL^T-style elimination (lane = row; only 1/d on the critical
//                              path, sqrt off it), then X_d = L_d^{-1} by forward substitution (lane = column)
//                     all      panel  P = A[below, jb] X_d^T          (DMMA, row-tile per warp, in place)
//                     all      trailing A[below, below] -= P P^T       (DMMA, lower 8x8 tiles)
//   finally the off-diagonal blocks of the inverse, X_ij = -X_d,i (sum_k L_ik X_kj), by DMMA, held in the
//   (otherwise unused) upper-triangle blocks of the shared-memory matrix.
// Rows/cols >= nb are padded with the identity so partial blocks need no special cases.
// (The previous register-resident variant with one CTA barrier per column took 108 us per block; ncu:
// profiles/r01b_potf2_ncu_full_summary.txt.)
// ------------------------------------------------------------------------------------------------
constexpr int PB = 32;               // panel width inside the diagonal block
constexpr int P_LDA = CK_NB + 4;     // smem row stride of the block: % 16 == 4 -> conflict-free DMMA fragment loads
constexpr int P_LDX = PB + 4;        // row stride of the four diagonal inverse blocks
constexpr int P_THREADS = 256;  // 8 warps: the 32-column elimination keeps 32 + 32 doubles per lane in registers
constexpr int P_SMEM_DOUBLES = CK_NB * P_LDA + (CK_NB / PB) * PB * P_LDX + CK_NB;

__global__ void __launch_bounds__(P_THREADS, 1)
    ck_potf2_inv_kernel(double* __restrict__ A, long long ld, int nb, double* __restrict__ X, int* info, int k0,
                        long long* __restrict__ dbg) {
  extern __shared__ __align__(16) double sm[];
  double* As = sm;                                   // As[i * P_LDA + k]
  double* Xd = As + CK_NB * P_LDA;                   // Xd[jb][r * P_LDX + c] = (L_d^{-1})[r][c]
  double* rsd = Xd + (CK_NB / PB) * PB * P_LDX;      // 1 / L[k][k]
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  constexpr int NW = P_THREADS / 32;
  if (tid == 0) bad = 0;
  int dbg_n = 0;
#define CK_P_STAMP() do { if (dbg && tid == 0) dbg[dbg_n++] = clock64(); } while (0)
  CK_P_STAMP();
  const bool vec = (nb == CK_NB) && ((((uintptr_t)A) & 15) == 0) && ((ld & 1) == 0);
  if (vec) {
    // full aligned block: every thread issues its 16 independent 16-byte loads (lower triangle only) before the
    // first shared-memory store, so one round trip to L2 / HBM covers the whole block
    constexpr int NV = CK_NB * CK_NB / 2 / P_THREADS;  // double2 per thread
    double2 v[NV];
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int idx = tid + it * P_THREADS, i = idx / (CK_NB / 2), k = 2 * (idx % (CK_NB / 2));
      v[it] = make_double2(0.0, 0.0);
      if (k <= i) v[it] = *reinterpret_cast<const double2*>(A + (long long)i * ld + k);
    }
#pragma unroll
    for (int it = 0; it < NV; ++it) {
      const int idx = tid + it * P_THREADS, i = idx / (CK_NB / 2), k = 2 * (idx % (CK_NB / 2));
      if (k + 1 > i) v[it].y = 0.0;  // (i, i + 1) is above the diagonal
      *reinterpret_cast<double2*>(As + i * P_LDA + k) = v[it];
    }
  } else {
    for (int idx = tid; idx < CK_NB * CK_NB; idx += P_THREADS) {
      const int i = idx / CK_NB, k = idx % CK_NB;
      double v = (i == k) ? 1.0 : 0.0;
      if (i < nb && k < nb) v = (k <= i) ? A[(long long)i * ld + k] : 0.0;
      As[i * P_LDA + k] = v;
    }
  }
  __syncthreads();
  CK_P_STAMP();  // 1: block loaded

  for (int jb = 0; jb < CK_NB / PB; ++jb) {
    const int c0 = jb * PB;
    double* Xdj = Xd + jb * PB * P_LDX;
    if (warp == 0) {
      // ---- D = L_d L_d^T, lane = row.  a[k] holds the UNSCALED column entry u_ik until column k is finished.
      double a[PB];
#pragma unroll
      for (int k = 0; k < PB; ++k) a[k] = As[(c0 + lane) * P_LDA + c0 + k];
      int first_bad = 0;
      double dk = 1.0;  // pivot of this lane's column
#pragma unroll
      for (int k = 0; k < PB; ++k) {
        const double d = __shfl_sync(0xffffffffu, a[k], k);
        if (!(d > 0.0) && first_bad == 0) first_bad = c0 + k + 1;
        if (lane == k) dk = d;
        // 1 / d on the critical path: approximate reciprocal + two Newton steps (<= 1 ulp); the square roots are
        // taken after the loop, for all 32 pivots at once, off the dependency chain
        double invd;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(invd) : "d"(d));
        invd = fma(fma(-d, invd, 1.0), invd, invd);
        invd = fma(fma(-d, invd, 1.0), invd, invd);
        const double u = a[k];
        const double w = u * invd;
#pragma unroll
        for (int j = k + 1; j < PB; ++j) {
          const double ujk = __shfl_sync(0xffffffffu, u, j);
          a[j] = fma(-w, ujk, a[j]);
        }
      }
      // a[k] holds the unscaled column entries u_ik (lane = row i); L_ik = u_ik / sqrt(d_k)
      const double sqv = sqrt(dk);
      const double rsv = 1.0 / sqv;
      rsd[c0 + lane] = rsv;
      if (lane == 0 && first_bad != 0 && bad == 0) bad = first_bad;
#pragma unroll
      for (int k = 0; k < PB; ++k) {
        const double rs = __shfl_sync(0xffffffffu, rsv, k);
        if (k <= lane) As[(c0 + lane) * P_LDA + c0 + k] = (lane == k) ? sqv : a[k] * rs;
      }
      __syncwarp();
      // ---- X_d = L_d^{-1}, lane = column c: forward substitution on e_c; L read by broadcast
      double x[PB];
#pragma unroll
      for (int i = 0; i < PB; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < PB; ++k) {
        x[k] *= rsd[c0 + k];
#pragma unroll
        for (int i = k + 1; i < PB; ++i) x[i] = fma(-As[(c0 + i) * P_LDA + c0 + k], x[k], x[i]);
      }
#pragma unroll
      for (int i = 0; i < PB; ++i) Xdj[i * P_LDX + lane] = x[i];
    }
    __syncthreads();
    CK_P_STAMP();  // 2 + 3 jb: warp-level elimination + inverse of the 32 x 32 diagonal block
    const int r_begin = c0 + PB;            // first row below the panel
    const int nt = (CK_NB - r_begin) / 8;   // 8-row tiles below
    // ---- panel: P = A[below, c0:c0+32] X_d^T, one 8-row tile (4 output tiles) per warp pass, in place
    for (int rt = warp; rt < nt; rt += NW) {
      const int r0 = r_begin + 8 * rt;
      double af[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) af[ks] = As[(r0 + g4) * P_LDA + c0 + 4 * ks + t4];
      double acc[4][2];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        acc[n][0] = 0.0;
        acc[n][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) dmma884(acc[n][0], acc[n][1], af[ks], Xdj[(8 * n + g4) * P_LDX + 4 * ks + t4]);
      }
      __syncwarp();  // every lane has its operand fragments before the tile row is overwritten
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        As[(r0 + g4) * P_LDA + c0 + 8 * n + 2 * t4] = acc[n][0];
        As[(r0 + g4) * P_LDA + c0 + 8 * n + 2 * t4 + 1] = acc[n][1];
      }
    }
    __syncthreads();
    CK_P_STAMP();  // 3 + 3 jb: panel
    // ---- trailing: A[below, below] -= P P^T on the lower 8x8 tiles
    const int ntile = nt * (nt + 1) / 2;
    for (int t = warp; t < ntile; t += NW) {
      int ti = 0;
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      const int tj = t - ti * (ti + 1) / 2;
      const int ri = r_begin + 8 * ti, rj = r_begin + 8 * tj;
      double* c = As + (ri + g4) * P_LDA + rj + 2 * t4;
      double acc0 = -c[0], acc1 = -c[1];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        dmma884(acc0, acc1, As[(ri + g4) * P_LDA + c0 + 4 * ks + t4], As[(rj + g4) * P_LDA + c0 + 4 * ks + t4]);
      const bool diag = (ti == tj);  // keep the strict upper triangle of the block untouched (scratch / zeros)
      if (!diag || 2 * t4 <= g4) c[0] = -acc0;
      if (!diag || 2 * t4 + 1 <= g4) c[1] = -acc1;
    }
    __syncthreads();
    CK_P_STAMP();  // 4 + 3 jb: trailing update
  }

  // ---- off-diagonal blocks of X = L^{-1}: X_ij = -X_d,i * S_ij, S_ij = sum_{k=j}^{i-1} L_ik X_kj (X_jj = X_d,j).
  // X_ij (i > j) lives in the unused upper block (j, i) of As: element [r][c] at As[(32 j + r) * P_LDA + 32 i + c].
  constexpr int NBK = CK_NB / PB;
  for (int dist = 1; dist < NBK; ++dist) {
    const int npair = NBK - dist;
    // stage 1: S_ij -> scratch block (j, i)
    for (int t = warp; t < npair * 16; t += NW) {
      const int j = t / 16, i = j + dist, mt = (t % 16) / 4, nt_ = t % 4;
      double acc0 = 0.0, acc1 = 0.0;
      for (int k = j; k < i; ++k) {
        const double* Lik = As + (PB * i + 8 * mt + g4) * P_LDA + PB * k + t4;        // A(m, kk) = L_ik[m][kk]
        const double* Xkj = (k == j) ? (Xd + j * PB * P_LDX + t4 * P_LDX + 8 * nt_ + g4)  // B(kk, n) = X_kj[kk][n]
                                     : (As + (PB * j + t4) * P_LDA + PB * k + 8 * nt_ + g4);
        const int ldb = (k == j) ? P_LDX : P_LDA;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) dmma884(acc0, acc1, Lik[4 * ks], Xkj[4 * ks * ldb]);
      }
      double* s = As + (PB * j + 8 * mt + g4) * P_LDA + PB * i + 8 * nt_ + 2 * t4;
      s[0] = acc0;
      s[1] = acc1;
    }
    __syncthreads();
    // stage 2: X_ij = -X_d,i * S_ij (in place: all tiles are computed into registers before any is written)
    double out[(48 + NW - 1) / NW][2];  // at most 48 tiles over NW warps
    int slot = 0;
    for (int t = warp; t < npair * 16; t += NW, ++slot) {
      const int j = t / 16, i = j + dist, mt = (t % 16) / 4, nt_ = t % 4;
      double acc0 = 0.0, acc1 = 0.0;
      const double* Xi = Xd + i * PB * P_LDX + (8 * mt + g4) * P_LDX + t4;            // A(m, kk) = X_d,i[m][kk]
      const double* S = As + (PB * j + t4) * P_LDA + PB * i + 8 * nt_ + g4;           // B(kk, n) = S[kk][n]
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) dmma884(acc0, acc1, Xi[4 * ks], S[4 * ks * P_LDA]);
      out[slot][0] = -acc0;
      out[slot][1] = -acc1;
    }
    __syncthreads();
    slot = 0;
    for (int t = warp; t < npair * 16; t += NW, ++slot) {
      const int j = t / 16, i = j + dist, mt = (t % 16) / 4, nt_ = t % 4;
      double* s = As + (PB * j + 8 * mt + g4) * P_LDA + PB * i + 8 * nt_ + 2 * t4;
      s[0] = out[slot][0];
      s[1] = out[slot][1];
    }
    __syncthreads();
  }

  CK_P_STAMP();  // 14: inverse assembled
  // ---- results: L in place (lower triangle of the valid part), X dense with an explicit zero upper triangle
  // address of X[i][k] (k even addresses a 16-byte aligned pair inside one 32-block): diagonal blocks live in Xd, the
  // blocks below the diagonal in the (otherwise unused) upper blocks of As; selected without divergence
  auto xsrc = [&](int i, int k) -> const double* {
    const int bi = i / PB, bj = k / PB, r = i % PB, c = k % PB;
    return (bi == bj) ? (Xd + bi * PB * P_LDX + r * P_LDX + c) : (As + (PB * bj + r) * P_LDA + PB * bi + c);
  };
  auto xval = [&](int i, int k) { return (i / PB >= k / PB) ? *xsrc(i, k) : 0.0; };
  if (vec) {
#pragma unroll 4
    for (int idx = tid; idx < CK_NB * CK_NB / 2; idx += P_THREADS) {
      const int i = idx / (CK_NB / 2), k = 2 * (idx % (CK_NB / 2));
      if (k + 1 <= i) *reinterpret_cast<double2*>(A + (long long)i * ld + k) = *reinterpret_cast<const double2*>(As + i * P_LDA + k);
      else if (k <= i) A[(long long)i * ld + k] = As[i * P_LDA + k];
      double2 xv = make_double2(0.0, 0.0);
      if (i / PB >= k / PB) xv = *reinterpret_cast<const double2*>(xsrc(i, k));
      *reinterpret_cast<double2*>(X + i * CK_NB + k) = xv;
    }
  } else {
    for (int idx = tid; idx < CK_NB * CK_NB; idx += P_THREADS) {
      const int i = idx / CK_NB, k = idx % CK_NB;
      if (i < nb && k <= i) A[(long long)i * ld + k] = As[i * P_LDA + k];
      X[i * CK_NB + k] = xval(i, k);
    }
  }
  if (tid == 0 && bad != 0 && *info == 0) *info = k0 + bad;
  CK_P_STAMP();  // 15: stored (this thread's part)
#undef CK_P_STAMP
}

// ------------------------------------------------------------------------------------------------
// small reductions (fixed summation order: strided per-thread partials, then a shared-memory tree)
// ------------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ double block_sum(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int s = T / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// one CTA per target row c:  pred[c] = V_c . y_c ;  var[c] = c0 + sgn |V_c|^2   (y_c = y + c * ldy; ldy = 0: one shared y)
// sgn = -1: simple-cokriging variance; c0 = 0, sgn = +1: partial row sums of the block-cyclic solve (ck_row_dots).
__global__ void __launch_bounds__(256) ck_predict_rows_kernel(const double* __restrict__ V, long long ldv, long long n,
                                                              const double* __restrict__ ybase, long long ldy, double c0,
                                                              double sgn, double* __restrict__ pred,
                                                              double* __restrict__ var) {
  __shared__ double red[256];
  const long long c = blockIdx.x;
  const double* v = V + c * ldv;
  const double* y = ybase + c * ldy;
  double s2 = 0.0, sy = 0.0;
  for (long long k = threadIdx.x; k < n; k += 256) {
    const double x = v[k];
    s2 += x * x;
    sy += x * y[k];
  }
  s2 = block_sum<256>(s2, red);
  sy = block_sum<256>(sy, red);
  if (threadIdx.x == 0) {
    pred[c] = sy;
    var[c] = c0 + sgn * s2;
  }
}

// out[0] = 2 sum log L_kk   (what = 0)      out[0] = sum y_k^2   (what = 1, `ld` ignored)
__global__ void __launch_bounds__(1024) ck_diag_reduce_kernel(const double* __restrict__ L, long long ld, long long n,
                                                              int what, double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (long long k = threadIdx.x; k < n; k += 1024) {
    if (what == 0) s += log(L[k * ld + k]);
    else { const double x = L[k]; s += x * x; }
  }
  s = block_sum<1024>(s, red);
  if (threadIdx.x == 0) out[0] = (what == 0) ? 2.0 * s : s;
}

__global__ void ck_nll_combine_kernel(double* out, long long n) {
  // out[1] = quadratic form, out[2] = logdet  ->  out[0] = 0.5 (quad + logdet + n log 2pi)
  out[0] = 0.5 * (out[1] + out[2] + (double)n * 1.8378770664093453);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// INT8 tensor-core path for the big trailing updates (ck_ozaki.cu).  CK_OZAKI=0 selects the FP64 DMMA kernel
// everywhere; CK_OZ_MIN_ROWS is the smallest trailing dimension handed to the INT8 kernel (below it the
// persistent kernel cannot fill the machine and the DMMA kernel is used).
// ------------------------------------------------------------------------------------------------
static int agg_blocks();
static int g_oz_enabled = -1;
static ck_i64 g_oz_min_rows = -1;
static int oz_enabled() {
  if (g_oz_enabled < 0) {
    const char* e = getenv("CK_OZAKI");
    g_oz_enabled = e ? (atoi(e) != 0) : 1;
  }
  return g_oz_enabled;
}
static ck_i64 oz_min_rows() {
  if (g_oz_min_rows < 0) {
    const char* e = getenv("CK_OZ_MIN_ROWS");
    g_oz_min_rows = e ? atoll(e) : 1024;
    if (g_oz_min_rows < 128) g_oz_min_rows = 128;
  }
  return g_oz_min_rows;
}
extern "C" int ck_oz_configure(int enabled, ck_i64 min_rows) {
  if (enabled >= 0) g_oz_enabled = enabled ? 1 : 0;
  if (min_rows >= 0) g_oz_min_rows = min_rows < 128 ? 128 : min_rows;
  return CK_OK;
}
// Look-ahead beside the persistent INT8 kernel: the update of aggregate j is split into (a) the columns of aggregate
// j + 1 (all SMs) and (b) the rest, launched on all but CK_OZ_LA_SMS SMs, while the panel chain of aggregate j + 1 (FP64
// DMMA + one-CTA diagonal blocks, mostly latency-bound) runs on a high-priority side stream on the SMs left free.
// Used while (b) is long enough to hide the chain (trailing dimension >= CK_OZ_LA_MIN_ROWS).  CK_OZ_LA_SMS=0: off.
static int oz_la_sms() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CK_OZ_LA_SMS");
    v = e ? atoi(e) : 10;
    if (v < 0 || v > 100) v = 10;
  }
  return v;
}
// CK_OZ_PRESPLIT (default 1): split the next panel on the side stream, one step ahead (two scratch sets)
static int oz_presplit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CK_OZ_PRESPLIT");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v;
}
static ck_i64 oz_la_min_rows() {
  static ck_i64 v = -1;
  if (v < 0) {
    const char* e = getenv("CK_OZ_LA_MIN_ROWS");
    v = e ? atoll(e) : 8192;
    if (v < 1024) v = 1024;
  }
  return v;
}
constexpr ck_i64 OZ_KMAX = 1024;  // deepest update one split covers (ck_oz_split)
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t xinv_bytes(ck_i64 n) {
  const size_t nblk = (size_t)((n + CK_NB - 1) / CK_NB);
  return align256(nblk * CK_NB * CK_NB * sizeof(double));
}
// scratch behind the diagonal-block inverses: A-format and B-format slices of one n x 1024 panel + two scale vectors
struct OzScratch {
  void* fa; void* fb; double* sa; double* sb;
};
// The workspace carries the slice scratch for every n >= OZ_WS_MIN_N whatever the switches say, so a workspace
// sized under one configuration stays valid under any other.
constexpr ck_i64 OZ_WS_MIN_N = 2048;
static bool oz_wanted(ck_i64 n) { return oz_enabled() && n >= OZ_WS_MIN_N && n >= 2 * oz_min_rows(); }
static size_t oz_scratch_bytes(ck_i64 n) {
  return align256(ck_oz_slices_bytes(n, OZ_KMAX, 0)) + align256(ck_oz_slices_bytes(n, OZ_KMAX, 1)) +
         2 * align256((size_t)ck_oz_scales_len(n) * sizeof(double));
}
// two sets (which = 0 / 1): while the update of aggregate j reads one, the side stream splits the finished panel of
// aggregate j + 1 into the other (the split is then off the main stream's critical path)
static OzScratch oz_scratch(void* ws, ck_i64 n, int which = 0) {
  char* p = static_cast<char*>(ws) + xinv_bytes(n) + (size_t)which * oz_scratch_bytes(n);
  OzScratch s;
  s.fa = p; p += align256(ck_oz_slices_bytes(n, OZ_KMAX, 0));
  s.fb = p; p += align256(ck_oz_slices_bytes(n, OZ_KMAX, 1));
  s.sa = reinterpret_cast<double*>(p); p += align256((size_t)ck_oz_scales_len(n) * sizeof(double));
  s.sb = reinterpret_cast<double*>(p);
  return s;
}

extern "C" int ck_oz_active(ck_i64 n) { return (oz_wanted(n) && (ck_i64)agg_blocks() * CK_NB <= OZ_KMAX) ? 1 : 0; }

extern "C" size_t ck_potrf_workspace_bytes(ck_i64 n) {
  if (n <= 0) return 0;
  size_t b = xinv_bytes(n);
  if (n >= OZ_WS_MIN_N) b += 2 * oz_scratch_bytes(n);
  return b;
}

// profiling aid (tools): 16 clock64 stamps of the last diagonal-block kernel launched
static long long* g_potf2_dbg = nullptr;
extern "C" int ck_potf2_debug_buffer(void* dev_stamps) {
  g_potf2_dbg = static_cast<long long*>(dev_stamps);
  return CK_OK;
}

static int potf2_launch(double* a, long long ld, int nb, double* x, int* info, int k0, cudaStream_t st) {
  constexpr size_t SMEM = P_SMEM_DOUBLES * sizeof(double);
  static CkPerDevice attr;
  CK_SET_SMEM_ONCE(attr, ck_potf2_inv_kernel, SMEM);
  ck_potf2_inv_kernel<<<1, P_THREADS, SMEM, st>>>(a, ld, nb, x, info, k0, g_potf2_dbg);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

// Aggregation width (in 128-blocks) of the trailing updates.  The factorisation is recursive inside
// an aggregate: panels are factored 128 columns at a time, but the O(N^3) trailing update is applied
// once per aggregate with K = 128 * CK_AGG, which halves (quarters, ...) the number of passes over the
// trailing matrix and amortises each tile's prologue/epilogue over a longer DMMA main loop.
static int agg_blocks() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("CK_AGG_BLOCKS");
    v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
    if (v > 64) v = 64;
  }
  return v;
}

struct CholCtx {
  double* a; ck_i64 n, ld; double* xinv; int* info; cudaStream_t st;
  ck_i64 s(ck_i64 b) const { return b * CK_NB < n ? b * CK_NB : n; }  // first row/col of block b (clamped)
};

// factor the column blocks [b0, b1) for all rows below, assuming every update from blocks < b0 was applied
static int factor_range(const CholCtx& c, ck_i64 b0, ck_i64 b1) {
  int rc;
  if (b1 - b0 == 1) {
    const ck_i64 k0 = c.s(b0), k1 = c.s(b1);
    const int nb = (int)(k1 - k0);
    double* xk = c.xinv + b0 * CK_NB * CK_NB;
    if ((rc = potf2_launch(c.a + k0 * c.ld + k0, c.ld, nb, xk, c.info, (int)k0, c.st))) return rc;
    if (k1 < c.n) {
      GemmArgs t;  // panel: A[k1:, k0:k1] <- A[k1:, k0:k1] X_k^T   (in place, one column tile)
      t.A = c.a + k1 * c.ld + k0; t.lda = c.ld;
      t.B = xk; t.ldb = CK_NB;
      t.C = c.a + k1 * c.ld + k0; t.ldc = c.ld;
      t.M = c.n - k1; t.N = nb; t.K = nb; t.mode = 0; t.lower_only = 0;
      if ((rc = gemm_launch<2, 4>(t, c.st))) return rc;
    }
    return CK_OK;
  }
  const ck_i64 mid = b0 + (b1 - b0 + 1) / 2;
  if ((rc = factor_range(c, b0, mid))) return rc;
  const ck_i64 r0 = c.s(mid), cl = c.s(b0), ce = c.s(b1);
  if (r0 < c.n) {
    GemmArgs u;  // A[r0:, r0:ce] -= A[r0:, cl:r0] A[r0:ce, cl:r0]^T   (thin update inside the aggregate)
    u.A = c.a + r0 * c.ld + cl; u.lda = c.ld;
    u.B = c.a + r0 * c.ld + cl; u.ldb = c.ld;
    u.C = c.a + r0 * c.ld + r0; u.ldc = c.ld;
    u.M = c.n - r0; u.N = ce - r0; u.K = r0 - cl; u.mode = 1; u.lower_only = 2;
    if ((rc = gemm_launch<4, 2>(u, c.st))) return rc;
    if ((rc = factor_range(c, mid, b1))) return rc;
  }
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
// Look-ahead.  The panel chain of an aggregate (4 x [potf2 -> panel TRSM] + thin updates) is a serial
// sequence of small, latency-bound launches (~0.6 ms per aggregate) during which most SMs idle.  With
// look-ahead the trailing update of aggregate j is split into (a) the columns of aggregate j+1 and
// (b) the rest; the panel chain of aggregate j+1 then runs on an internal high-priority stream
// concurrently with (b).  Every entry still receives the same K-ordered sum, so results are bitwise
// identical with and without look-ahead.  CK_LOOKAHEAD=0 disables it.
// ------------------------------------------------------------------------------------------------
#include <mutex>
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev_a = nullptr, ev_c = nullptr;
  std::mutex enqueue;  // host threads enqueueing on the same device take turns (the events are shared)
};
static int lookahead_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CK_LOOKAHEAD");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v;
}
// One side stream per (device, caller stream), so that independent factorisations enqueued on different
// caller streams (batched windows) do not serialise their panel chains on a shared side stream.
static int side_stream(cudaStream_t caller, SideStream** out) {
  constexpr int CAP = 64;
  struct Slot { int dev; cudaStream_t caller; SideStream* s; };
  static Slot slots[CAP];
  static int used = 0;
  static std::mutex mu;
  int dev = 0;
  CK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int fallback = -1;
  for (int i = 0; i < used; ++i) {
    if (slots[i].dev != dev) continue;
    if (slots[i].caller == caller) { *out = slots[i].s; return CK_OK; }
    if (fallback < 0) fallback = i;
  }
  if (used == CAP) {  // table full: share the first side stream of this device (correct, merely less concurrent)
    CK_REQUIRE(fallback >= 0, "side-stream table full");
    *out = slots[fallback].s;
    return CK_OK;
  }
  SideStream* t = new SideStream();
  int lo = 0, hi = 0;
  CK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK_CUDA(cudaStreamCreateWithPriority(&t->s, cudaStreamNonBlocking, hi));
  CK_CUDA(cudaEventCreateWithFlags(&t->ev_a, cudaEventDisableTiming));
  CK_CUDA(cudaEventCreateWithFlags(&t->ev_c, cudaEventDisableTiming));
  slots[used++] = Slot{dev, caller, t};
  *out = t;
  return CK_OK;
}

extern "C" int ck_potrf(double* a, ck_i64 n, ck_i64 ld, void* ws, int* info, void* stream) {
  CK_REQUIRE(n >= 0, "negative size");
  CK_REQUIRE(info, "info is NULL");
  cudaStream_t st = ck_stream(stream);
  CK_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (n == 0) return CK_OK;
  CK_REQUIRE(a && ws, "null pointer");
  CK_REQUIRE(ld >= n, "ld (%lld) < n (%lld)", (long long)ld, (long long)n);
  CholCtx c{a, n, ld, static_cast<double*>(ws), info, st};
  const ck_i64 nblk = (n + CK_NB - 1) / CK_NB;
  const int agg = agg_blocks();
  int rc;
  if (oz_wanted(n) && (ck_i64)agg * CK_NB <= OZ_KMAX) {
    // Big trailing updates on the INT8 tensor cores: per aggregate, the panel chain (FP64 DMMA, recursive), one split
    // of the finished panel into digit slices, one persistent tcgen05 kernel over the lower tiles of the trailing matrix.
    const OzScratch zs[2] = {oz_scratch(ws, n, 0), oz_scratch(ws, n, 1)};
    int cur = 0;
    bool have = false;  // zs[cur] already holds the slices of the current panel (split on the side stream, one step ahead)
    const int la = oz_la_sms();
    SideStream* oside = nullptr;
    if (la > 0 && lookahead_enabled() && n >= 2 * oz_la_min_rows()) {
      if ((rc = side_stream(st, &oside))) return rc;
    }
    std::unique_lock<std::mutex> oturn;
    if (oside) oturn = std::unique_lock<std::mutex>(oside->enqueue);
    CholCtx ocs = c;
    if (oside) ocs.st = oside->s;
    if ((rc = factor_range(c, 0, agg < nblk ? agg : nblk))) return rc;
    for (ck_i64 b0 = 0; b0 < nblk; b0 += agg) {
      const ck_i64 b1 = b0 + agg < nblk ? b0 + agg : nblk;
      const ck_i64 k0 = c.s(b0), k1 = c.s(b1);
      if (k1 >= n) break;
      const ck_i64 b2 = b1 + agg < nblk ? b1 + agg : nblk;
      const ck_i64 k2 = c.s(b2);
      const ck_i64 rows = n - k1, kk = k1 - k0, off = k2 - k1;
      const bool use_oz = rows >= oz_min_rows() && kk % 32 == 0;
      const bool la_now = oside && use_oz && (n - k2) >= oz_la_min_rows() && off % 128 == 0;
      const OzScratch z = zs[cur];
      if (use_oz && !have) {
        if ((rc = ck_oz_split(a + k1 * ld + k0, ld, rows, kk, z.fa, z.fb, z.sa, stream))) return rc;
      }
      have = false;
      if (la_now) {
        // (a) columns of the next aggregate: A[k1:, k1:k2] -= P P[k1:k2]^T  (every SM)
        if ((rc = ck_oz_gemm(z.fa, z.sa, rows, z.fb, z.sa, off, kk, a + k1 * ld + k1, ld, 1, 0, stream))) return rc;
        CK_CUDA(cudaEventRecord(oside->ev_a, st));
        CK_CUDA(cudaStreamWaitEvent(oside->s, oside->ev_a, 0));
        if ((rc = factor_range(ocs, b1, b2))) return rc;  // panel chain of the next aggregate, beside (b)
        if (oz_presplit() && k2 < n && n - k2 >= oz_min_rows() && off % 32 == 0) {
          // ... and the split of that panel into the OTHER scratch set (its last reader, update (b) of the previous step,
          // finished before (a) of this one): the next step starts with its slices ready
          const OzScratch zn = zs[cur ^ 1];
          if ((rc = ck_oz_split(a + k2 * ld + k1, ld, n - k2, off, zn.fa, zn.fb, zn.sa, oside->s))) return rc;
          have = true;
        }
        CK_CUDA(cudaEventRecord(oside->ev_c, oside->s));
        // (b) the rest: A[k2:, k2:] -= P[k2:] P[k2:]^T on all but `la` SMs
        const char* fa2 = static_cast<const char*>(z.fa) + (size_t)(off / 128) * ck_oz_slices_bytes(128, kk, 0);
        const char* fb2 = static_cast<const char*>(z.fb) + (size_t)(off / 64) * ck_oz_slices_bytes(64, kk, 1);
        if ((rc = ck_oz_gemm(fa2, z.sa + off, rows - off, fb2, z.sa + off, rows - off, kk, a + k2 * ld + k2, ld, 1,
                             ck_oz_num_sms() - la, stream)))
          return rc;
        CK_CUDA(cudaStreamWaitEvent(st, oside->ev_c, 0));
      } else {
        if (use_oz) {
          if ((rc = ck_oz_gemm(z.fa, z.sa, rows, z.fb, z.sa, rows, kk, a + k1 * ld + k1, ld, 1, 0, stream))) return rc;
        } else {
          GemmArgs u;  // A[k1:, k1:] -= P P^T, P = A[k1:, k0:k1], lower tiles
          u.A = a + k1 * ld + k0; u.lda = ld;
          u.B = a + k1 * ld + k0; u.ldb = ld;
          u.C = a + k1 * ld + k1; u.ldc = ld;
          u.M = rows; u.N = rows; u.K = kk; u.mode = 1; u.lower_only = 1;
          if ((rc = gemm_launch<4, 2>(u, st))) return rc;
        }
        if ((rc = factor_range(c, b1, b2))) return rc;
      }
      if (have) cur ^= 1;
    }
    return CK_OK;
  }
  SideStream* side = nullptr;
  if (lookahead_enabled() && nblk > 2 * agg) {
    if ((rc = side_stream(st, &side))) return rc;
  }
  std::unique_lock<std::mutex> turn;
  if (side) turn = std::unique_lock<std::mutex>(side->enqueue);
  CholCtx cs = c;
  if (side) cs.st = side->s;
  if ((rc = factor_range(c, 0, agg < nblk ? agg : nblk))) return rc;
  for (ck_i64 b0 = 0; b0 < nblk; b0 += agg) {
    const ck_i64 b1 = b0 + agg < nblk ? b0 + agg : nblk;
    const ck_i64 k0 = c.s(b0), k1 = c.s(b1);
    if (k1 >= n) break;
    const ck_i64 b2 = b1 + agg < nblk ? b1 + agg : nblk;
    const ck_i64 k2 = c.s(b2);
    {
      GemmArgs u;  // (a) columns of the next aggregate: A[k1:, k1:k2] -= P P[k1:k2]^T, P = A[k1:, k0:k1]
      u.A = a + k1 * ld + k0; u.lda = ld;
      u.B = a + k1 * ld + k0; u.ldb = ld;
      u.C = a + k1 * ld + k1; u.ldc = ld;
      u.M = n - k1; u.N = k2 - k1; u.K = k1 - k0; u.mode = 1; u.lower_only = 2;
      if ((rc = gemm_launch<4, 2>(u, st))) return rc;
    }
    GemmArgs s;  // (b) the rest: A[k2:, k2:] -= P P^T with P = A[k2:, k0:k1], lower tiles, K = 128 * agg
    s.A = a + k2 * ld + k0; s.lda = ld;
    s.B = a + k2 * ld + k0; s.ldb = ld;
    s.C = a + k2 * ld + k2; s.ldc = ld;
    s.M = n - k2; s.N = n - k2; s.K = k1 - k0; s.mode = 1; s.lower_only = 1;
    if (side && k2 < n) {
      CK_CUDA(cudaEventRecord(side->ev_a, st));
      CK_CUDA(cudaStreamWaitEvent(side->s, side->ev_a, 0));
      if ((rc = factor_range(cs, b1, b2))) return rc;  // panel chain of the next aggregate, concurrently with (b)
      CK_CUDA(cudaEventRecord(side->ev_c, side->s));
      if ((rc = gemm_launch<4, 2>(s, st))) return rc;
      CK_CUDA(cudaStreamWaitEvent(st, side->ev_c, 0));
    } else {
      if (k2 < n && (rc = gemm_launch<4, 2>(s, st))) return rc;
      if ((rc = factor_range(c, b1, b2))) return rc;
    }
  }
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
// Few right-hand sides (the likelihood objective solves with ONE): the update R[:, k1:] -= V[:, k0:k1] L[k1:, k0:k1]^T is a
// matrix-vector product per right-hand side -- HBM-bound (every entry of L is read once), so it is a streaming kernel, not a
// DMMA GEMM on a 128 x 64 tile with one valid row.  One warp per row j of L: lane-strided partial sums over the K <= 1024
// columns (16-byte loads), fixed xor-shuffle tree => deterministic.  V[:, k0:k1] sits in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int TRSV_MAX_RHS = 8;
template <int NR>
__global__ void __launch_bounds__(256) ck_trsv_update_kernel(const double* __restrict__ lpan, long long ld, long long rows, int kk,
                                                             double* __restrict__ rhs, long long ld_rhs, long long k0,
                                                             long long k1, int vec) {
  extern __shared__ double vs[];  // [NR][kk]
  for (int i = threadIdx.x; i < NR * kk; i += 256) vs[i] = rhs[(long long)(i / kk) * ld_rhs + k0 + (i % kk)];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long j = (long long)blockIdx.x * 8 + warp; j < rows; j += (long long)gridDim.x * 8) {
    const double* lrow = lpan + j * ld;
    double acc[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) acc[c] = 0.0;
    if (vec) {
      for (int k = 2 * lane; k < kk; k += 64) {
        const double2 lv = *reinterpret_cast<const double2*>(lrow + k);
#pragma unroll
        for (int c = 0; c < NR; ++c) acc[c] = fma(lv.y, vs[c * kk + k + 1], fma(lv.x, vs[c * kk + k], acc[c]));
      }
    } else {
      for (int k = lane; k < kk; k += 32) {
        const double lv = lrow[k];
#pragma unroll
        for (int c = 0; c < NR; ++c) acc[c] = fma(lv, vs[c * kk + k], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      double a = acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) rhs[(long long)c * ld_rhs + k1 + j] -= a;
    }
  }
}

static int trsv_update_launch(const double* lpan, ck_i64 ld, ck_i64 rows, ck_i64 kk, double* rhs, ck_i64 ld_rhs, ck_i64 nrhs,
                              ck_i64 k0, ck_i64 k1, cudaStream_t st) {
  const int vec = ((((uintptr_t)lpan) & 15) == 0 && (ld & 1) == 0 && (kk & 1) == 0) ? 1 : 0;
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const size_t smem = (size_t)nrhs * kk * sizeof(double);  // <= 8 x 1024 x 8 B = 64 KB
#define CK_TRSV(NR)                                                                                                    \
  case NR: {                                                                                                           \
    static CkPerDevice attr;                                                                                           \
    CK_SET_SMEM_ONCE(attr, ck_trsv_update_kernel<NR>, NR * 1024 * sizeof(double));                                     \
    ck_trsv_update_kernel<NR><<<(unsigned)blocks, 256, smem, st>>>(lpan, ld, rows, (int)kk, rhs, ld_rhs, k0, k1, vec); \
  } break;
  switch ((int)nrhs) {
    CK_TRSV(1) CK_TRSV(2) CK_TRSV(3) CK_TRSV(4) CK_TRSV(5) CK_TRSV(6) CK_TRSV(7) CK_TRSV(8)
    default: CK_REQUIRE(false, "trsv path needs 1..8 right-hand sides");
  }
#undef CK_TRSV
  CK_LAUNCH_CHECK();
  return CK_OK;
}

struct TrsmCtx {
  const double* l; ck_i64 n, ld; const double* xinv; double* rhs; ck_i64 nrhs, ldr; cudaStream_t st;
  ck_i64 s(ck_i64 b) const { return b * CK_NB < n ? b * CK_NB : n; }
};

// forward substitution restricted to the column blocks [b0, b1) of the target-major right-hand sides
static int solve_range(const TrsmCtx& c, ck_i64 b0, ck_i64 b1) {
  int rc;
  if (b1 - b0 == 1) {
    const ck_i64 k0 = c.s(b0);
    const int nb = (int)(c.s(b1) - k0);
    GemmArgs t;  // V[:, k] = R[:, k] X_k^T  (in place)
    t.A = c.rhs + k0; t.lda = c.ldr;
    t.B = c.xinv + b0 * CK_NB * CK_NB; t.ldb = CK_NB;
    t.C = c.rhs + k0; t.ldc = c.ldr;
    t.M = c.nrhs; t.N = nb; t.K = nb; t.mode = 0; t.lower_only = 0;
    return gemm_launch<2, 4>(t, c.st);
  }
  const ck_i64 mid = b0 + (b1 - b0 + 1) / 2;
  if ((rc = solve_range(c, b0, mid))) return rc;
  const ck_i64 r0 = c.s(mid), cl = c.s(b0), ce = c.s(b1);
  if (r0 < c.n) {
    GemmArgs u;  // R[:, r0:ce] -= V[:, cl:r0] L[r0:ce, cl:r0]^T
    u.A = c.rhs + cl; u.lda = c.ldr;
    u.B = c.l + r0 * c.ld + cl; u.ldb = c.ld;
    u.C = c.rhs + r0; u.ldc = c.ldr;
    u.M = c.nrhs; u.N = ce - r0; u.K = r0 - cl; u.mode = 1; u.lower_only = 0;
    if ((rc = gemm_launch<4, 2>(u, c.st))) return rc;
    if ((rc = solve_range(c, mid, b1))) return rc;
  }
  return CK_OK;
}

extern "C" int ck_trsm_lower(const double* l, ck_i64 n, ck_i64 ld, const void* ws, double* rhs, ck_i64 nrhs, ck_i64 ld_rhs,
                             void* stream) {
  CK_REQUIRE(n >= 0 && nrhs >= 0, "negative size");
  if (n == 0 || nrhs == 0) return CK_OK;
  CK_REQUIRE(l && ws && rhs, "null pointer");
  CK_REQUIRE(ld >= n && ld_rhs >= n, "leading dimension too small");
  cudaStream_t st = ck_stream(stream);
  TrsmCtx c{l, n, ld, static_cast<const double*>(ws), rhs, nrhs, ld_rhs, st};
  const ck_i64 nblk = (n + CK_NB - 1) / CK_NB;
  const int agg = agg_blocks();
  int rc;
  if (nrhs <= TRSV_MAX_RHS && (ck_i64)agg * CK_NB <= 1024) {
    // few right-hand sides: per aggregate the small solves on <= 1024 columns (a single-CTA substitution kernel was tried
    // and is slower: one CTA cannot pull the 4 MB tile fast enough), then one streaming matrix-vector update to the right
    for (ck_i64 b0 = 0; b0 < nblk; b0 += agg) {
      const ck_i64 b1 = b0 + agg < nblk ? b0 + agg : nblk;
      if ((rc = solve_range(c, b0, b1))) return rc;
      const ck_i64 k0 = c.s(b0), k1 = c.s(b1);
      if (k1 >= n) break;
      if ((rc = trsv_update_launch(l + k1 * ld + k0, ld, n - k1, k1 - k0, rhs, ld_rhs, nrhs, k0, k1, st))) return rc;
    }
    return CK_OK;
  }
  if (oz_wanted(n) && (ck_i64)agg * CK_NB <= OZ_KMAX && nrhs >= 1024 && nrhs <= n) {
    // INT8 tensor-core updates (see ck_potrf).  The slice scratch lives behind the block inverses in the
    // factorisation workspace: solves that share one factor must not run concurrently.
    const OzScratch zs[2] = {oz_scratch(const_cast<void*>(ws), n, 0), oz_scratch(const_cast<void*>(ws), n, 1)};
    int cur = 0;
    bool have = false;  // zs[cur] already holds the slices of V[:, k0:k1] and L[k1:, k0:k1] (split on the side stream)
    const int la = oz_la_sms();
    SideStream* oside = nullptr;
    if (la > 0 && lookahead_enabled() && n >= 2 * oz_la_min_rows()) {
      if ((rc = side_stream(st, &oside))) return rc;
    }
    std::unique_lock<std::mutex> oturn;
    if (oside) oturn = std::unique_lock<std::mutex>(oside->enqueue);
    TrsmCtx ocs = c;
    if (oside) ocs.st = oside->s;
    if ((rc = solve_range(c, 0, agg < nblk ? agg : nblk))) return rc;
    for (ck_i64 b0 = 0; b0 < nblk; b0 += agg) {
      const ck_i64 b1 = b0 + agg < nblk ? b0 + agg : nblk;
      const ck_i64 k0 = c.s(b0), k1 = c.s(b1);
      if (k1 >= n) break;
      const ck_i64 b2 = b1 + agg < nblk ? b1 + agg : nblk;
      const ck_i64 k2 = c.s(b2);
      const ck_i64 cols = n - k1, kk = k1 - k0, off = k2 - k1;
      const bool use_oz = cols >= oz_min_rows() && kk % 32 == 0;
      const bool la_now = oside && use_oz && (n - k2) >= oz_la_min_rows() && off % 64 == 0;
      const OzScratch z = zs[cur];
      if (use_oz && !have) {
        if ((rc = ck_oz_split(rhs + k0, ld_rhs, nrhs, kk, z.fa, nullptr, z.sa, stream))) return rc;
        if ((rc = ck_oz_split(l + k1 * ld + k0, ld, cols, kk, nullptr, z.fb, z.sb, stream))) return rc;
      }
      have = false;
      if (la_now) {
        // (a) columns of the next aggregate: R[:, k1:k2] -= V L[k1:k2]^T  (every SM)
        if ((rc = ck_oz_gemm(z.fa, z.sa, nrhs, z.fb, z.sb, off, kk, rhs + k1, ld_rhs, 0, 0, stream))) return rc;
        CK_CUDA(cudaEventRecord(oside->ev_a, st));
        CK_CUDA(cudaStreamWaitEvent(oside->s, oside->ev_a, 0));
        const bool presplit = oz_presplit() && k2 < n && n - k2 >= oz_min_rows() && off % 32 == 0;
        const OzScratch zn = zs[cur ^ 1];
        if (presplit) {  // the next L panel is final already: split it first, then the small solves, then their result
          if ((rc = ck_oz_split(l + k2 * ld + k1, ld, n - k2, off, nullptr, zn.fb, zn.sb, oside->s))) return rc;
        }
        if ((rc = solve_range(ocs, b1, b2))) return rc;  // small solves of the next aggregate, beside (b)
        if (presplit) {
          if ((rc = ck_oz_split(rhs + k1, ld_rhs, nrhs, off, zn.fa, nullptr, zn.sa, oside->s))) return rc;
          have = true;
        }
        CK_CUDA(cudaEventRecord(oside->ev_c, oside->s));
        // (b) the rest: R[:, k2:] -= V L[k2:]^T on all but `la` SMs
        const char* fb2 = static_cast<const char*>(z.fb) + (size_t)(off / 64) * ck_oz_slices_bytes(64, kk, 1);
        if ((rc = ck_oz_gemm(z.fa, z.sa, nrhs, fb2, z.sb + off, cols - off, kk, rhs + k2, ld_rhs, 0, ck_oz_num_sms() - la,
                             stream)))
          return rc;
        CK_CUDA(cudaStreamWaitEvent(st, oside->ev_c, 0));
      } else {
        if (use_oz) {
          if ((rc = ck_oz_gemm(z.fa, z.sa, nrhs, z.fb, z.sb, cols, kk, rhs + k1, ld_rhs, 0, 0, stream))) return rc;
        } else {
          GemmArgs r;  // R[:, k1:] -= V[:, k0:k1] L[k1:, k0:k1]^T
          r.A = rhs + k0; r.lda = ld_rhs;
          r.B = l + k1 * ld + k0; r.ldb = ld;
          r.C = rhs + k1; r.ldc = ld_rhs;
          r.M = nrhs; r.N = cols; r.K = kk; r.mode = 1; r.lower_only = 0;
          if ((rc = gemm_launch<4, 2>(r, st))) return rc;
        }
        if ((rc = solve_range(c, b1, b2))) return rc;
      }
      if (have) cur ^= 1;
    }
    return CK_OK;
  }
  SideStream* side = nullptr;
  if (lookahead_enabled() && nblk > 2 * agg && nrhs >= 1024) {
    if ((rc = side_stream(st, &side))) return rc;
  }
  std::unique_lock<std::mutex> turn;
  if (side) turn = std::unique_lock<std::mutex>(side->enqueue);
  TrsmCtx cs = c;
  if (side) cs.st = side->s;
  if ((rc = solve_range(c, 0, agg < nblk ? agg : nblk))) return rc;
  for (ck_i64 b0 = 0; b0 < nblk; b0 += agg) {
    const ck_i64 b1 = b0 + agg < nblk ? b0 + agg : nblk;
    const ck_i64 k0 = c.s(b0), k1 = c.s(b1);
    if (k1 >= n) break;
    const ck_i64 b2 = b1 + agg < nblk ? b1 + agg : nblk;
    const ck_i64 k2 = c.s(b2);
    {
      GemmArgs u;  // (a) columns of the next aggregate: R[:, k1:k2] -= V[:, k0:k1] L[k1:k2, k0:k1]^T
      u.A = rhs + k0; u.lda = ld_rhs;
      u.B = l + k1 * ld + k0; u.ldb = ld;
      u.C = rhs + k1; u.ldc = ld_rhs;
      u.M = nrhs; u.N = k2 - k1; u.K = k1 - k0; u.mode = 1; u.lower_only = 0;
      if ((rc = gemm_launch<4, 2>(u, st))) return rc;
    }
    GemmArgs r;  // (b) the rest: R[:, k2:] -= V[:, k0:k1] L[k2:, k0:k1]^T, K = 128 * agg
    r.A = rhs + k0; r.lda = ld_rhs;
    r.B = l + k2 * ld + k0; r.ldb = ld;
    r.C = rhs + k2; r.ldc = ld_rhs;
    r.M = nrhs; r.N = n - k2; r.K = k1 - k0; r.mode = 1; r.lower_only = 0;
    if (side && k2 < n) {
      CK_CUDA(cudaEventRecord(side->ev_a, st));
      CK_CUDA(cudaStreamWaitEvent(side->s, side->ev_a, 0));
      if ((rc = solve_range(cs, b1, b2))) return rc;  // small solves of the next aggregate, concurrently with (b)
      CK_CUDA(cudaEventRecord(side->ev_c, side->s));
      if ((rc = gemm_launch<4, 2>(r, st))) return rc;
      CK_CUDA(cudaStreamWaitEvent(st, side->ev_c, 0));
    } else {
      if (k2 < n && (rc = gemm_launch<4, 2>(r, st))) return rc;
      if ((rc = solve_range(c, b1, b2))) return rc;
    }
  }
  return CK_OK;
}

extern "C" int ck_potrs_predict(const double* l, ck_i64 n, ck_i64 ld, const void* ws, double* cpd, ck_i64 m, ck_i64 ld_c,
                                const double* z, double c0, double* pred, double* var, void* stream) {
  CK_REQUIRE(n >= 0 && m >= 0, "negative size");
  if (n == 0) return CK_OK;
  CK_REQUIRE(l && ws && cpd && z, "null pointer");
  CK_REQUIRE(m == 0 || (pred && var), "null output");
  cudaStream_t st = ck_stream(stream);
  CK_CUDA(cudaMemcpyAsync(cpd + m * ld_c, z, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  int rc = ck_trsm_lower(l, n, ld, ws, cpd, m + 1, ld_c, stream);
  if (rc) return rc;
  if (m > 0) {
    ck_predict_rows_kernel<<<(unsigned)m, 256, 0, st>>>(cpd, ld_c, n, cpd + m * ld_c, 0, c0, -1.0, pred, var);
    CK_LAUNCH_CHECK();
  }
  return CK_OK;
}

extern "C" int ck_logdet(const double* l, ck_i64 n, ck_i64 ld, double* out, void* stream) {
  CK_REQUIRE(n >= 0 && l && out, "bad argument");
  ck_diag_reduce_kernel<<<1, 1024, 0, ck_stream(stream)>>>(l, ld, n, 0, out);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_nll(const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* params, int n_procs,
                      int metric, const double* z, double* sigma, ck_i64 ld, void* ws, double* scratch, double* out,
                      int* info, void* stream) {
  if (n_procs == 1) n1 = 0;
  const ck_i64 n = n0 + n1;
  CK_REQUIRE(z && sigma && ws && scratch && out && info, "null pointer");
  cudaStream_t st = ck_stream(stream);
  int rc;
  if ((rc = ck_joint_cov(xy0, n0, xy1, n1, params, n_procs, metric, sigma, ld, stream))) return rc;
  if ((rc = ck_potrf(sigma, n, ld, ws, info, stream))) return rc;
  CK_CUDA(cudaMemcpyAsync(scratch, z, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if ((rc = ck_trsm_lower(sigma, n, ld, ws, scratch, 1, n, stream))) return rc;
  ck_diag_reduce_kernel<<<1, 1024, 0, st>>>(scratch, 0, n, 1, out + 1);
  ck_diag_reduce_kernel<<<1, 1024, 0, st>>>(sigma, ld, n, 0, out + 2);
  ck_nll_combine_kernel<<<1, 1, 0, st>>>(out, n);
  CK_LAUNCH_CHECK_N(3);
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
// building blocks of the multi-GPU (2-D block-cyclic) factorisation -- orchestrated per rank by
// cokrig_b200/parallel.py over torch.distributed / NCCL
// ------------------------------------------------------------------------------------------------
extern "C" int ck_gemm_nt(const double* a, ck_i64 lda, const double* b, ck_i64 ldb, double* c, ck_i64 ldc, ck_i64 m, ck_i64 n,
                          ck_i64 k, int subtract, void* stream) {
  CK_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative size");
  if (m == 0 || n == 0) return CK_OK;
  CK_REQUIRE(a && b && c, "null pointer");
  CK_REQUIRE(lda >= k && ldb >= k && ldc >= n, "leading dimension too small");
  GemmArgs g;
  g.A = a; g.lda = lda; g.B = b; g.ldb = ldb; g.C = c; g.ldc = ldc;
  g.M = m; g.N = n; g.K = k; g.mode = subtract ? 1 : 0; g.lower_only = 0;
  return gemm_launch<4, 2>(g, ck_stream(stream));
}

extern "C" int ck_mg_update(const double* a, ck_i64 lda, const double* b, ck_i64 ldb, double* c, ck_i64 ldc, ck_i64 m,
                            ck_i64 n, ck_i64 k, ck_i64 tb, ck_i64 row_tile0, ck_i64 row_tile_step, ck_i64 col_tile0,
                            ck_i64 col_tile_step, void* stream) {
  CK_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative size");
  if (m == 0 || n == 0) return CK_OK;
  CK_REQUIRE(a && b && c, "null pointer");
  CK_REQUIRE(lda >= k && ldb >= k && ldc >= n, "leading dimension too small");
  CK_REQUIRE(tb > 0 && tb % 128 == 0, "tile size must be a multiple of 128 (got %lld)", (long long)tb);
  CK_REQUIRE(m % tb == 0 && n % tb == 0, "m and n must be whole tiles");
  CK_REQUIRE(row_tile_step >= 1 && col_tile_step >= 1 && row_tile0 >= 0 && col_tile0 >= 0, "bad tile map");
  GemmArgs g;
  g.A = a; g.lda = lda; g.B = b; g.ldb = ldb; g.C = c; g.ldc = ldc;
  g.M = m; g.N = n; g.K = k; g.mode = 1; g.lower_only = 3;
  g.tb = (int)tb; g.gi0 = row_tile0; g.gis = row_tile_step; g.gj0 = col_tile0; g.gjs = col_tile_step;
  return gemm_launch<4, 2>(g, ck_stream(stream));
}

extern "C" int ck_row_dots(const double* v, ck_i64 ldv, ck_i64 nrows, ck_i64 ncols, const double* y, ck_i64 ldy,
                           double* out_vy, double* out_vv, void* stream) {
  CK_REQUIRE(nrows >= 0 && ncols >= 0, "negative size");
  if (nrows == 0) return CK_OK;
  CK_REQUIRE(v && y && out_vy && out_vv, "null pointer");
  ck_predict_rows_kernel<<<(unsigned)nrows, 256, 0, ck_stream(stream)>>>(v, ldv, ncols, y, ldy, 0.0, 1.0, out_vy, out_vv);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

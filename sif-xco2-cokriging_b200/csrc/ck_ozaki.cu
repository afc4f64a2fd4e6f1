// ck_ozaki.cu -- FP64-equivalent rank-K updates  C -= A B^T  on the INT8 tensor cores of sm_100a
// (tcgen05.mma kind::i8, accumulators in TMEM), used for the O(N^3) trailing updates of the blocked
// Cholesky and of the multi-right-hand-side triangular solve (ck_chol.cu).
//
// Scheme (Ozaki-style error-free splitting with a fixed number of slices):
//   every row of an operand panel (K <= 1024 columns) is scaled by a power of two so that |a| <= 0.98 and
//   rounded ONCE to a 55-bit fixed-point integer Q (the only rounding of the operands, 2^-56 relative to the
//   row maximum -- below FP64's own 2^-53), then recoded EXACTLY into S = 7 balanced base-256 digits
//   d_0..d_6 in [-128, 127]:   a = 2^(e-7) sum_p d_p 2^(-8p).
//   The product panel is   A B^T = s_a s_b sum_g 2^(-8g) G_g ,  G_g = sum_{p+q=g} A_p B_q^T   (int8 x int8 -> int32,
//   exact: |G_g| <= 7 * 1024 * 2^14 < 2^31).  Groups g = 0..6 are kept (28 slice products); the dropped groups
//   g >= 7 are zero-mean terms below 2^-56 of (row max) x (column max) per k -- the same order as the FP64
//   rounding of a DMMA update.  tests/ozaki_model.py reproduces the scheme in numpy (Cholesky + solve errors
//   equal to LAPACK's to the last digit on the C1 system).
//
// Kernel (ck_oz_gemm_kernel, persistent, one CTA per SM, 320 threads, warp-specialised):
//   CTA tile 128 (rows of A) x 64 (rows of B); ALL seven group accumulators live in TMEM at once
//   (7 x 64 = 448 of 512 columns), so every operand byte is fetched once per tile and the FP64 epilogue runs
//   once per tile (K = 1024 deep), not once per group.
//   warp 0  producer : per 32-deep k chunk one cp.async.bulk of the 7 A slices (28 KB) and one of the 7 B slices
//                      (14 KB) into a 4-stage shared-memory ring (mbarrier complete_tx).  The split kernel writes
//                      the slices to global memory already in the UMMA canonical K-major (no-swizzle) core-matrix
//                      order, so a stage is two contiguous copies -- no tensor maps.
//   warp 1  MMA      : one thread issues 10 tcgen05.mma per chunk instead of 28: the B slices q..q+3 are adjacent
//                      in shared memory with the row-group stride of one slice, so A_p x [B_q..B_q+3] is ONE
//                      M=128, N=256 instruction whose 256 accumulator columns are exactly the group blocks
//                      p+q..p+q+3 (A is read 10 times per chunk instead of 28).
//   warps 2-9 epilogue: tcgen05.ld the 7 int32 group blocks, exact 64-bit integer recombination three groups at a
//                      time, 3 conversions + 2 FMAs in FP64, scale by s_a s_b (powers of two: exact) and C -= v.
#include <stdint.h>
#include <stdlib.h>
#include <mutex>
#include "ck_common.cuh"

namespace {

constexpr int OZ_S = 7;                        // slices per operand
constexpr int OZ_QBITS = 7 + 8 * (OZ_S - 1);   // 55-bit fixed point
constexpr int OZ_KC = 32;                      // int8 k per chunk == K of one tcgen05.mma kind::i8
constexpr int OZ_TM = 128, OZ_TN = 64;
constexpr int OZ_A_SLICE = OZ_TM * OZ_KC;      // 4096 B: [ku 2][row group 16][row 8][16 B]
constexpr int OZ_A_STAGE = OZ_S * OZ_A_SLICE;  // 28672 B: [p 7][slice]
constexpr int OZ_B_KU = OZ_S * OZ_TN * 16;     // 7168 B:  [q 7][row group 8][row 8][16 B]
constexpr int OZ_B_STAGE = 2 * OZ_B_KU;        // 14336 B: [ku 2][...]
constexpr int OZ_STAGE = OZ_A_STAGE + OZ_B_STAGE;
constexpr int OZ_NSTAGE = 4;
constexpr int OZ_NSCHED = 4;                   // depth of the tile-id ring of the dynamic scheduler
constexpr int OZ_EPI_WARPS = 8;                // 2 per TMEM lane quadrant, 32 columns each
constexpr int OZ_THREADS = 64 + 32 * OZ_EPI_WARPS;
constexpr int OZ_TMEM_COLS = 512;
constexpr size_t OZ_SMEM = (size_t)OZ_NSTAGE * OZ_STAGE + 1024 /* alignment slack */ + 256 /* barriers */;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol error must end in a trap (an error the host sees), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
// L2 eviction priorities (OzGemmArgs::l2_hints): the operand slices are re-read by every tile of a super-row while the C
// tiles stream through exactly once (one read, one write), so the slices are loaded evict_last and C evict_first
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 0 normal, 1 evict_last, 2 evict_first
  uint64_t p;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_c2(const double* p, uint64_t policy) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy) : "memory");
  return v;
}
__device__ __forceinline__ void st_c2(double* p, double2 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 B contiguous (128 B);
// LBO = byte distance between the two 16-byte k units of one K=32 instruction, SBO = byte distance between
// consecutive 8-row groups along M / N (cute::UMMA::SmemDescriptor; layout ((8,n),2):((1,SBO),LBO) in 16-byte units).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = S32 (bits 4-5 = 2), A, B = signed int8 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t umma_idesc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_TM >> 4) << 24);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Split: FP64 panel rows -> 7 int8 digit slices in the two operand formats + per-row scale 2^(e-7)
//   A format (128-row blocks):  [rb][kc][p][ku][row group 16][row 8][16 B]
//   B format ( 64-row blocks):  [hb][kc][ku][q][row group 8][row 8][16 B]
// One CTA = 8 rows (one row group), one warp per row, a lane owns the 16-element k units lane and lane + 32; the digit
// bytes go through a shared-memory staging area so that global memory only sees full 128-byte lines.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ck_oz_split_kernel(const double* __restrict__ src, long long ld, long long rows, int K,
                                                          uint8_t* __restrict__ fa, uint8_t* __restrict__ fb,
                                                          long long rows_b_pad, double* __restrict__ scales, int vec) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + w;
  const int units = K >> 4, kcn = K >> 5;
  const bool live = row < rows;
  double x[2][16];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int u = lane + 32 * h;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[h][i] = 0.0;
    if (live && u < units) {
      const double* p = src + row * ld + 16 * u;
      if (vec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(p + 2 * i);
          x[h][2 * i] = v.x;
          x[h][2 * i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[h][i] = p[i];
      }
    }
  }
  double amax = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 16; ++i) amax = fmax(amax, fabs(x[h][i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  int ex = 0;
  if (amax > 0.0 && amax <= 1.7976931348623157e308) {
    const double f = frexp(amax, &ex);  // amax = f 2^ex, f in [0.5, 1)
    if (f > 0.98) ++ex;                 // |a| 2^-ex <= 0.98: the leading digit stays inside [-126, 126]
    if (ex < -900) ex = -900;
  }
  const double up = __hiloint2double((1023 + OZ_QBITS - ex) << 20, 0);  // 2^(55 - ex), exact scaling
  if (lane == 0) scales[row] = (live && amax > 0.0) ? __hiloint2double((1023 + ex - 7) << 20, 0) : 0.0;

  extern __shared__ __align__(16) uint8_t stage[];  // units x 7 lines of 128 B
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int u = lane + 32 * h;
    if (u >= units) continue;
    uint32_t wd[OZ_S][4];
#pragma unroll
    for (int p = 0; p < OZ_S; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) wd[p][j] = 0u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      long long q = __double2ll_rn(x[h][i] * up);
#pragma unroll
      for (int p = OZ_S - 1; p >= 0; --p) {
        const int d = (int)((q + 128) & 255) - 128;  // balanced digit in [-128, 127]
        q = (q - d) >> 8;
        wd[p][i >> 2] |= (uint32_t)(d & 255) << (8 * (i & 3));
      }
    }
    // stage the 7 x 16 digit bytes of this (row, unit) in shared memory in output order: line (u, p) = 8 rows x 16 B,
    // the 16-byte slot of row w XOR-swizzled with u so that the warp's 32 stores (same w, 32 units) spread over all banks
#pragma unroll
    for (int p = 0; p < OZ_S; ++p)
      *reinterpret_cast<uint4*>(stage + ((size_t)(u * OZ_S + p) * 8 + (w ^ (u & 7))) * 16) =
          make_uint4(wd[p][0], wd[p][1], wd[p][2], wd[p][3]);
  }
  __syncthreads();
  // copy out: every 128-byte line (8 rows of one 16-byte k unit of one slice) is one contiguous piece of both operand
  // formats; 8 threads move one line, a warp writes four full lines per instruction
  const int lines = units * OZ_S;
  const long long row0 = (long long)blockIdx.x * 8;
  const long long rb = row0 >> 7, hb = row0 >> 6;
  const int rg16 = (int)(row0 & 127) >> 3, rg8 = (int)(row0 & 63) >> 3;
  const int sub = threadIdx.x & 7;
  for (int ln = threadIdx.x >> 3; ln < lines; ln += 32) {
    const int u = ln / OZ_S, p = ln - u * OZ_S, kc = u >> 1, ku = u & 1;
    const uint4 v = *reinterpret_cast<const uint4*>(stage + ((size_t)ln * 8 + (sub ^ (u & 7))) * 16);
    if (fa)
      *reinterpret_cast<uint4*>(fa + ((size_t)(rb * kcn + kc)) * OZ_A_STAGE + (size_t)p * OZ_A_SLICE + ku * 2048 + rg16 * 128 +
                                sub * 16) = v;
    if (fb && row0 < rows_b_pad)
      *reinterpret_cast<uint4*>(fb + ((size_t)(hb * kcn + kc)) * OZ_B_STAGE + ku * OZ_B_KU + p * (OZ_TN * 16) + rg8 * 128 +
                                sub * 16) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------------
struct OzGemmArgs {
  const uint8_t* a;   // A-format slices of the M-side panel
  const uint8_t* b;   // B-format slices of the N-side panel
  const double* sa;   // row scales of A (length >= 128 ni)
  const double* sb;   // row scales of B (length >= 64 nj)
  double* c;
  long long ldc, m, n;
  int kcn;            // K / 32
  int ni, nj;         // 128-row blocks of A, 64-row blocks of B
  int lower;          // 1: only entries with col <= row are updated (tiles above the diagonal are skipped)
  int tri;            // 1: triangular virtual tile space (square lower update), 0: rectangular
  int sr;             // row blocks per super-row of the tile order
  long long nvirt;    // virtual tiles
  // block-cyclic mode (tb > 0): C is the local part of a 2-D block-cyclic matrix with square tiles of tb elements; local
  // tile (li, lj) is global tile (gi0 + li gis, gj0 + lj gjs); tiles with J > I are skipped, tiles with J == I keep
  // their lower triangle (same contract as ck_mg_update)
  int tb;
  long long gi0, gis, gj0, gjs;
  int vec;            // C is 16-byte aligned with an even leading dimension
  int l2_hints;       // bit 0: operand slices evict_last, bit 1: C loads / stores evict_first (CK_OZ_L2_HINTS)
  // dynamic tile scheduler (CK_OZ_DYNAMIC, default on): sched[0] = next virtual tile (atomic counter), sched[1] = CTAs done;
  // the last CTA to finish resets both, so the stream's slot is 0 again when the launch ends.  NULL: static grid-stride assignment.
  unsigned long long* sched;
  long long* dbg;     // optional per-CTA cycle counters (8 per CTA), see ck_oz_debug_buffer
};

// virtual tile t -> (I, j).  Tiles are ordered in super-rows of sr (default 16, CK_OZ_SUPER_ROWS) row blocks, inside a
// super-row column-major (j outer, I inner), so that the ~148 tiles in flight share 16 A blocks and ~9 B blocks (L2 reuse).
__host__ __device__ __forceinline__ bool oz_decode(const OzGemmArgs& g, long long t, int& I, int& j) {
  long long s, u;
  const long long sr = g.sr;  // row blocks per super-row
  if (g.tri) {
    // super-row s holds sr row blocks x 2 sr (s + 1) column blocks; prefix sum sr^2 s (s + 1)
    const long long q = sr * sr;
    s = (long long)((sqrt(1.0 + 4.0 * (double)t / (double)q) - 1.0) * 0.5);
    while (q * (s + 1) * (s + 2) <= t) ++s;
    while (q * s * (s + 1) > t) --s;
    u = t - q * s * (s + 1);
  } else {
    const long long w = sr * g.nj;
    s = t / w;
    u = t - s * w;
  }
  j = (int)(u / sr);
  I = (int)(sr * s + (u % sr));
  if (I >= g.ni || j >= g.nj) return false;
  if (g.tb) {
    const long long li = ((long long)I * OZ_TM) / g.tb, lj = ((long long)j * OZ_TN) / g.tb;
    const long long Ig = g.gi0 + li * g.gis, Jg = g.gj0 + lj * g.gjs;
    if (Jg > Ig) return false;
    return Jg < Ig || ((long long)j * OZ_TN - lj * g.tb) <= ((long long)I * OZ_TM + OZ_TM - 1 - li * g.tb);
  }
  return !g.lower || j <= 2 * I + 1;
}

// largest column of C that row `row` may update inside column block j (LLONG_MAX: no mask)
__host__ __device__ __forceinline__ long long oz_col_limit(const OzGemmArgs& g, long long row, int j) {
  if (g.tb) {
    const long long li = row / g.tb, lj = ((long long)j * OZ_TN) / g.tb;
    const long long Ig = g.gi0 + li * g.gis, Jg = g.gj0 + lj * g.gjs;
    return Jg == Ig ? lj * g.tb + (row - li * g.tb) : 0x7fffffffffffffffLL;
  }
  return g.lower ? row : 0x7fffffffffffffffLL;
}

__global__ void __launch_bounds__(OZ_THREADS, 1) ck_oz_gemm_kernel(OzGemmArgs g) {
  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)oz_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)OZ_NSTAGE * OZ_STAGE);
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full, [2S+1] tmem_empty (S = OZ_NSTAGE), then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * OZ_NSTAGE + 2);
  // tile-id ring of the dynamic scheduler: the producer fetches the next virtual tile with one atomicAdd and hands (I, j) to the
  // MMA thread and the epilogue warps; bars[2S+3 ..) = OZ_NSCHED full + OZ_NSCHED empty barriers, then the ids
  int* tile_ring = reinterpret_cast<int*>(bars + 2 * OZ_NSTAGE + 3 + 2 * OZ_NSCHED);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (OZ_NSTAGE + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * OZ_NSTAGE), tempty_bar = bar0 + 8u * (2 * OZ_NSTAGE + 1);
  auto sfull_bar = [&](int s) { return bar0 + 8u * (2 * OZ_NSTAGE + 3 + s); };
  auto sempty_bar = [&](int s) { return bar0 + 8u * (2 * OZ_NSTAGE + 3 + OZ_NSCHED + s); };
  const bool dyn = g.sched != nullptr;
  const uint32_t smem0 = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OZ_NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 32 * OZ_EPI_WARPS);
    for (int s = 0; s < OZ_NSCHED; ++s) {
      mbar_init(sfull_bar(s), 1);
      mbar_init(sempty_bar(s), 1 + OZ_EPI_WARPS);  // the MMA thread + one lane per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)OZ_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol = l2_policy((g.l2_hints & 1) ? 1 : 0);
      int sslot = 0;
      uint32_t sphase = 0;
      for (long long t = blockIdx.x;; t += gridDim.x) {
        int I = 0, j = 0;
        if (dyn) {
          // next valid tile in the global order: all CTAs stay inside one window of ~gridDim.x consecutive tiles (the
          // A blocks of one super-row and a few B blocks: L2-resident), and nobody idles at the tail
          do {
            t = (long long)atomicAdd(g.sched, 1ULL);
          } while (t < g.nvirt && !oz_decode(g, t, I, j));
          mbar_wait(sempty_bar(sslot), sphase ^ 1u);
          tile_ring[sslot] = t < g.nvirt ? ((I << 16) | j) : -1;
          mbar_arrive(sfull_bar(sslot));
          if (++sslot == OZ_NSCHED) { sslot = 0; sphase ^= 1u; }
          if (t >= g.nvirt) break;
        } else {
          if (t >= g.nvirt) break;
          if (!oz_decode(g, t, I, j)) continue;
        }
        const uint8_t* ga = g.a + (size_t)I * g.kcn * OZ_A_STAGE;
        const uint8_t* gb = g.b + (size_t)j * g.kcn * OZ_B_STAGE;
        for (int kc = 0; kc < g.kcn; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), OZ_STAGE);
          const uint32_t dst = smem0 + (uint32_t)stage * OZ_STAGE;
          bulk_g2s(dst, ga + (size_t)kc * OZ_A_STAGE, OZ_A_STAGE, full_bar(stage), pol);
          bulk_g2s(dst + OZ_A_STAGE, gb + (size_t)kc * OZ_B_STAGE, OZ_B_STAGE, full_bar(stage), pol);
          if (++stage == OZ_NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      long long t_full = 0, t_tempty = 0;
      const long long t_start = clock64();
      const uint32_t a_lbo = 2048u, a_sbo = 128u;                // A slice: [ku][row group][8 rows][16 B]
      const uint32_t b_lbo = (uint32_t)OZ_B_KU, b_sbo = 128u;   // B k-unit: [q][row group][8 rows][16 B]
      const uint32_t id256 = umma_idesc_i8(256), id192 = umma_idesc_i8(192), id128 = umma_idesc_i8(128), id64 = umma_idesc_i8(64);
      int sslot = 0;
      uint32_t sphase = 0;
      for (long long t = blockIdx.x;; t += gridDim.x) {
        int I, j;
        if (dyn) {
          mbar_wait(sfull_bar(sslot), sphase);
          const int v = tile_ring[sslot];
          mbar_arrive(sempty_bar(sslot));
          if (++sslot == OZ_NSCHED) { sslot = 0; sphase ^= 1u; }
          if (v < 0) break;
        } else {
          if (t >= g.nvirt) break;
          if (!oz_decode(g, t, I, j)) continue;
        }
        long long w0 = clock64();
        mbar_wait(tempty_bar, aphase ^ 1u);
        tc_fence_after();
        t_tempty += clock64() - w0;
        for (int kc = 0; kc < g.kcn; ++kc) {
          w0 = clock64();
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          t_full += clock64() - w0;
          const uint32_t sa0 = smem0 + (uint32_t)stage * OZ_STAGE, sb0 = sa0 + OZ_A_STAGE;
          const uint32_t acc0 = kc > 0 ? 1u : 0u;
          auto A = [&](int p) { return umma_desc(sa0 + p * OZ_A_SLICE, a_lbo, a_sbo); };
          auto B = [&](int q) { return umma_desc(sb0 + q * (OZ_TN * 16), b_lbo, b_sbo); };
          auto D = [&](int gblk) { return tmem + (uint32_t)(gblk * OZ_TN); };
          // A_p x [B_q0 .. B_q0+nq-1] -> group blocks p+q0 .. (one instruction, N = 64 nq)
          tc_mma_i8(D(0), A(0), B(0), id256, acc0);  // p = 0 touches every block first: it alone (re)initialises
          tc_mma_i8(D(4), A(0), B(4), id192, acc0);
          tc_mma_i8(D(1), A(1), B(0), id256, 1u);
          tc_mma_i8(D(5), A(1), B(4), id128, 1u);
          tc_mma_i8(D(2), A(2), B(0), id256, 1u);
          tc_mma_i8(D(6), A(2), B(4), id64, 1u);
          tc_mma_i8(D(3), A(3), B(0), id256, 1u);
          tc_mma_i8(D(4), A(4), B(0), id192, 1u);
          tc_mma_i8(D(5), A(5), B(0), id128, 1u);
          tc_mma_i8(D(6), A(6), B(0), id64, 1u);
          tc_commit(empty_bar(stage));  // frees the stage when these MMAs have read it
          if (++stage == OZ_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar);
        aphase ^= 1u;
      }
      if (g.dbg) {
        g.dbg[blockIdx.x * 8 + 0] = clock64() - t_start;  // MMA-issue thread: total, waiting for operands, waiting for TMEM
        g.dbg[blockIdx.x * 8 + 1] = t_full;
        g.dbg[blockIdx.x * 8 + 2] = t_tempty;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warps 2..9; warp w reads TMEM lane quadrant w % 4, column half (w - 2) / 4; a thread owns one
    // row x 32 columns =====
    // The C row segment is loaded BEFORE waiting for the accumulators (its latency hides behind the MMAs of the
    // tile), TMEM is released as soon as the last block has been read, and the stores are issued after that, so
    // only TMEM reads + the FP64 recombination sit between two tiles of the MMA stream.
    constexpr int EC = OZ_TN / (OZ_EPI_WARPS / 4);  // columns per thread
    const int qd = warp & 3, half = (warp - 2) >> 2;
    uint32_t aphase = 0;
    const uint64_t cpol = l2_policy((g.l2_hints & 2) ? 2 : 0);
    long long t_busy = 0, t_wait = 0;
    int sslot = 0;
    uint32_t sphase = 0;
    for (long long t = blockIdx.x;; t += gridDim.x) {
      int I, j;
      if (dyn) {
        mbar_wait(sfull_bar(sslot), sphase);
        const int v = tile_ring[sslot];
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty_bar(sslot));
        if (++sslot == OZ_NSCHED) { sslot = 0; sphase ^= 1u; }
        if (v < 0) break;
        I = v >> 16;
        j = v & 0xffff;
      } else {
        if (t >= g.nvirt) break;
        if (!oz_decode(g, t, I, j)) continue;
      }
      const long long row = (long long)I * OZ_TM + 32 * qd + lane;
      const long long cb = (long long)j * OZ_TN + EC * half;
      const bool row_ok = row < g.m;
      const double srow = row_ok ? g.sa[row] * 1.52587890625e-05 /* 2^-16, see the recombination below */ : 0.0;
      double* crow = g.c + (row_ok ? row : 0) * g.ldc + cb;
      const double sb_l = (cb + lane < g.n) ? __ldg(g.sb + cb + lane) : 0.0;
      const long long clim = oz_col_limit(g, row_ok ? row : 0, j);
      const bool fast = g.vec && row_ok && (cb + EC - 1 < g.n) && (cb + EC - 1 <= clim);
      double creg[EC];
      if (fast) {
#pragma unroll
        for (int i = 0; i < EC; i += 2) {
          const double2 cv = ld_c2(crow + i, cpol);
          creg[i] = cv.x;
          creg[i + 1] = cv.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < EC; ++i)
          creg[i] = (row_ok && cb + i < g.n && cb + i <= clim) ? crow[i] : 0.0;
      }
      const long long w0 = clock64();
      mbar_wait(tfull_bar, aphase);
      tc_fence_after();
      const long long w1 = clock64();
#pragma unroll
      for (int c8 = 0; c8 < EC / 8; ++c8) {
        int acc[OZ_S][8];
        const uint32_t taddr = tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(EC * half + 8 * c8);
#pragma unroll
        for (int gb = 0; gb < OZ_S; ++gb) tc_ld8(taddr + (uint32_t)(gb * OZ_TN), acc[gb]);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // sum_g 2^(-8g) G_g: the groups are combined three at a time in exact 64-bit integer arithmetic
          // (|G_g| < 2^28, so |G_0 2^16 + G_1 2^8 + G_2| < 2^45), which leaves 3 conversions + 2 FMAs on the FP64 pipe
          // instead of 7 + 6 -- the drain of the accumulators is FP64-pipe bound
          static_assert(OZ_S == 7, "group recombination is written for 7 slices");
          const long long t0 = ((long long)acc[0][i] << 16) + ((long long)acc[1][i] << 8) + (long long)acc[2][i];
          const long long t1 = ((long long)acc[3][i] << 16) + ((long long)acc[4][i] << 8) + (long long)acc[5][i];
          double h = fma((double)t1, 5.9604644775390625e-08 /* 2^-24 */, (double)t0);
          h = fma((double)acc[6][i], 2.3283064365386963e-10 /* 2^-32 */, h);
          const int cl = 8 * c8 + i;
          const double sbv = __shfl_sync(0xffffffffu, sb_l, cl);
          creg[cl] -= h * (srow * sbv);  // srow carries the remaining 2^-16
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);  // TMEM is free: the MMAs of the next tile start while the stores below drain
      aphase ^= 1u;
      if (fast) {
#pragma unroll
        for (int i = 0; i < EC; i += 2) st_c2(crow + i, make_double2(creg[i], creg[i + 1]), cpol);
      } else if (row_ok) {
#pragma unroll
        for (int i = 0; i < EC; ++i)
          if (cb + i < g.n && cb + i <= clim) crow[i] = creg[i];
      }
      const long long w2 = clock64();
      t_wait += w1 - w0;
      t_busy += w2 - w1;
    }
    if (g.dbg && warp == 2 && lane == 0) {
      g.dbg[blockIdx.x * 8 + 4] = t_wait;
      g.dbg[blockIdx.x * 8 + 5] = t_busy;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)OZ_TMEM_COLS) : "memory");
  }
  if (dyn && threadIdx.x == 0) {
    // every CTA has fetched its last tile before it gets here: the last one to arrive leaves the slot zeroed for its next user
    if (atomicAdd(g.sched + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL) {
      g.sched[0] = 0ULL;
      g.sched[1] = 0ULL;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
extern "C" size_t ck_oz_slices_bytes(ck_i64 rows, ck_i64 k, int fmt) {
  if (rows <= 0 || k <= 0) return 0;
  const size_t kcn = (size_t)((k + OZ_KC - 1) / OZ_KC);
  if (fmt == 0) return (size_t)((rows + OZ_TM - 1) / OZ_TM) * kcn * OZ_A_STAGE;
  return (size_t)((rows + OZ_TN - 1) / OZ_TN) * kcn * OZ_B_STAGE;
}

extern "C" ck_i64 ck_oz_scales_len(ck_i64 rows) { return rows <= 0 ? 0 : ((rows + OZ_TM - 1) / OZ_TM) * OZ_TM; }

extern "C" int ck_oz_split(const double* src, ck_i64 ld, ck_i64 rows, ck_i64 k, void* fmt_a, void* fmt_b, double* scales,
                           void* stream) {
  CK_REQUIRE(rows >= 0 && k >= 0, "negative size");
  if (rows == 0 || k == 0) return CK_OK;
  CK_REQUIRE(src && scales && (fmt_a || fmt_b), "null pointer");
  CK_REQUIRE(k % OZ_KC == 0 && k <= 1024, "k (%lld) must be a multiple of 32 and <= 1024", (long long)k);
  CK_REQUIRE(ld >= k, "ld (%lld) < k (%lld)", (long long)ld, (long long)k);
  const ck_i64 rows_pad = ck_oz_scales_len(rows);
  const ck_i64 rows_b_pad = ((rows + OZ_TN - 1) / OZ_TN) * OZ_TN;
  const int vec = ((((uintptr_t)src) & 15) == 0 && (ld & 1) == 0) ? 1 : 0;
  const size_t smem = (size_t)(k / 16) * OZ_S * 128;  // <= 56 KB
  static CkPerDevice attr;
  CK_SET_SMEM_ONCE(attr, ck_oz_split_kernel, 64 * OZ_S * 128);
  ck_oz_split_kernel<<<(unsigned)(rows_pad / 8), 256, smem, ck_stream(stream)>>>(src, ld, rows, (int)k, static_cast<uint8_t*>(fmt_a),
                                                                            static_cast<uint8_t*>(fmt_b), rows_b_pad, scales, vec);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

constexpr int OZ_SCHED_SLOTS = 256;
__device__ unsigned long long g_oz_sched_slots[2 * OZ_SCHED_SLOTS];  // dynamic tile scheduler: (next tile, CTAs done) per stream

static long long* g_oz_dbg = nullptr;
extern "C" int ck_oz_debug_buffer(void* dev_counters) {
  g_oz_dbg = static_cast<long long*>(dev_counters);
  return CK_OK;
}

static int oz_num_sms() {
  static int v = 0;  // one process drives identical GPUs
  if (v == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  }
  return v;
}
int ck_oz_num_sms() { return oz_num_sms(); }

// tile space of one launch: row / column blocks, masks, and the virtual tile order (super-rows, column-major inside)
static void oz_geometry(OzGemmArgs& g, ck_i64 m, ck_i64 n, int lower, ck_i64 tb, ck_i64 gi0, ck_i64 gis, ck_i64 gj0, ck_i64 gjs) {
  g.m = m; g.n = n;
  g.ni = (int)((m + OZ_TM - 1) / OZ_TM);
  g.nj = (int)((n + OZ_TN - 1) / OZ_TN);
  g.lower = (lower && !tb) ? 1 : 0;
  g.tb = (int)tb; g.gi0 = gi0; g.gis = gis; g.gj0 = gj0; g.gjs = gjs;
  static int sr_cfg = 0;
  if (sr_cfg == 0) {
    const char* e = getenv("CK_OZ_SUPER_ROWS");
    sr_cfg = e ? atoi(e) : 16;  // measured on the largest C3 update: 8 -> 94, 16 -> 97, 32 -> 98 TFLOP/s; DRAM reads 30 -> 22 GB
    if (sr_cfg < 1 || sr_cfg > 64) sr_cfg = 16;
  }
  g.sr = sr_cfg;
  const long long srows = (g.ni + g.sr - 1) / g.sr;
  // the triangular tile space needs every super-row s to hold its 16 s + 16 column blocks; use it for (near-)square
  // lower updates, the rectangular space otherwise
  g.tri = (g.lower && (long long)g.nj >= 2LL * g.ni - 1) ? 1 : 0;
  g.nvirt = g.tri ? (long long)g.sr * g.sr * srows * (srows + 1) : srows * (long long)g.sr * g.nj;
}

// Host-only: the order in which the update kernel visits the tiles of a problem shape (what the dynamic scheduler hands
// out), for the CPU test tier -- every tile that must be updated appears exactly once, no masked tile appears.
extern "C" ck_i64 ck_oz_tile_order(ck_i64 m, ck_i64 n, int lower, ck_i64 tb, ck_i64 row_tile0, ck_i64 row_tile_step, ck_i64 col_tile0,
                                   ck_i64 col_tile_step, int* ij_out, ck_i64 cap, ck_i64* nvirt_out, ck_i64* col_limits_out) {
  if (m <= 0 || n <= 0 || (tb && (tb % 128 || m % tb || n % tb)) || row_tile_step < 1 || col_tile_step < 1) return -1;
  OzGemmArgs g = {};
  oz_geometry(g, m, n, lower, tb, row_tile0, row_tile_step, col_tile0, col_tile_step);
  if (nvirt_out) *nvirt_out = g.nvirt;
  ck_i64 count = 0;
  for (long long t = 0; t < g.nvirt; ++t) {
    int I, j;
    if (!oz_decode(g, t, I, j)) continue;
    if (ij_out && count < cap) {
      ij_out[2 * count] = I;
      ij_out[2 * count + 1] = j;
      // largest column the FIRST row of the tile may update inside column block j (-1: no mask)
      if (col_limits_out) {
        const long long lim = oz_col_limit(g, (long long)I * OZ_TM, j);
        col_limits_out[count] = lim == 0x7fffffffffffffffLL ? -1 : lim;
      }
    }
    ++count;
  }
  return count;
}

static int oz_gemm_launch(const void* a_slices, const double* sa, ck_i64 m, const void* b_slices, const double* sb, ck_i64 n,
                          ck_i64 k, double* c, ck_i64 ldc, int lower, ck_i64 tb, ck_i64 gi0, ck_i64 gis, ck_i64 gj0, ck_i64 gjs,
                          int max_ctas, void* stream) {
  CK_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative size");
  if (m == 0 || n == 0 || k == 0) return CK_OK;
  CK_REQUIRE(a_slices && b_slices && sa && sb && c, "null pointer");
  CK_REQUIRE(k % OZ_KC == 0 && k <= 1024, "k (%lld) must be a multiple of 32 and <= 1024", (long long)k);
  CK_REQUIRE(ldc >= n, "ldc (%lld) < n (%lld)", (long long)ldc, (long long)n);
  CK_REQUIRE((((uintptr_t)a_slices | (uintptr_t)b_slices) & 15) == 0, "slice buffers must be 16-byte aligned");
  static CkPerDevice attr;
  CK_SET_SMEM_ONCE(attr, ck_oz_gemm_kernel, OZ_SMEM);
  OzGemmArgs g;
  g.a = static_cast<const uint8_t*>(a_slices);
  g.b = static_cast<const uint8_t*>(b_slices);
  g.sa = sa; g.sb = sb; g.c = c; g.ldc = ldc;
  g.kcn = (int)(k / OZ_KC);
  oz_geometry(g, m, n, lower, tb, gi0, gis, gj0, gjs);
  g.vec = ((((uintptr_t)c) & 15) == 0 && (ldc & 1) == 0) ? 1 : 0;
  static int hints_cfg = -1;
  if (hints_cfg < 0) {
    const char* e = getenv("CK_OZ_L2_HINTS");
    // default 3 (both hints): measured on the C3 step 21 371 -> 21 536 predictions/s, the largest update alone 16.72 -> 15.79 ms
    // (profiles/r02x_l2hints_sweep.log); 1: 21 483, 2: 21 411
    hints_cfg = e ? (atoi(e) & 3) : 3;
  }
  g.l2_hints = hints_cfg;
  static int dyn_cfg = -1;
  if (dyn_cfg < 0) {
    const char* e = getenv("CK_OZ_DYNAMIC");
    dyn_cfg = e ? (atoi(e) != 0) : 1;
  }
  g.sched = nullptr;
  if (dyn_cfg && g.ni < 32768 && g.nj < 65536) {
    // One counter pair per (device, stream), from statically allocated device slots (no allocation): launches on one stream
    // never overlap and every launch leaves its slot zeroed, so a stream can reuse its slot for ever, while launches on
    // different streams (look-ahead beside the trailing update, batched windows) never share one.  More than OZ_SCHED_SLOTS
    // streams on a device: the extra streams fall back to the static assignment.
    struct Slot { int dev; cudaStream_t st; };
    static Slot tab[OZ_SCHED_SLOTS];
    static int used = 0;
    static unsigned long long* base[64] = {};
    static std::mutex mu;
    int dev = 0;
    CK_CUDA(cudaGetDevice(&dev));
    CK_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    cudaStream_t cst = ck_stream(stream);
    std::lock_guard<std::mutex> lock(mu);
    if (!base[dev]) {
      void* ptr = nullptr;
      CK_CUDA(cudaGetSymbolAddress(&ptr, g_oz_sched_slots));
      base[dev] = static_cast<unsigned long long*>(ptr);
    }
    int idx = -1, on_dev = 0;
    for (int i = 0; i < used; ++i) {
      if (tab[i].dev != dev) continue;
      if (tab[i].st == cst) { idx = on_dev; break; }
      ++on_dev;
    }
    if (idx < 0 && on_dev < OZ_SCHED_SLOTS && used < OZ_SCHED_SLOTS) {
      tab[used++] = Slot{dev, cst};
      idx = on_dev;
    }
    if (idx >= 0) g.sched = base[dev] + 2 * idx;  // idx = rank of this stream among the streams seen on this device
  }
  g.dbg = g_oz_dbg;
  long long grid = oz_num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;  // caller leaves SMs to a concurrent stream (look-ahead)
  if (grid > g.nvirt) grid = g.nvirt;
  ck_oz_gemm_kernel<<<(unsigned)grid, OZ_THREADS, OZ_SMEM, ck_stream(stream)>>>(g);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_oz_gemm(const void* a_slices, const double* sa, ck_i64 m, const void* b_slices, const double* sb, ck_i64 n,
                          ck_i64 k, double* c, ck_i64 ldc, int lower, int max_ctas, void* stream) {
  return oz_gemm_launch(a_slices, sa, m, b_slices, sb, n, k, c, ldc, lower, 0, 0, 1, 0, 1, max_ctas, stream);
}

extern "C" int ck_oz_mg_update(const void* a_slices, const double* sa, ck_i64 m, const void* b_slices, const double* sb, ck_i64 n,
                               ck_i64 k, double* c, ck_i64 ldc, ck_i64 tb, ck_i64 row_tile0, ck_i64 row_tile_step, ck_i64 col_tile0,
                               ck_i64 col_tile_step, int max_ctas, void* stream) {
  CK_REQUIRE(tb > 0 && tb % 128 == 0, "tile size must be a multiple of 128 (got %lld)", (long long)tb);
  CK_REQUIRE(m % tb == 0 && n % tb == 0, "m and n must be whole tiles");
  CK_REQUIRE(row_tile_step >= 1 && col_tile_step >= 1 && row_tile0 >= 0 && col_tile0 >= 0, "bad tile map");
  return oz_gemm_launch(a_slices, sa, m, b_slices, sb, n, k, c, ldc, 0, tb, row_tile0, row_tile_step, col_tile0, col_tile_step,
                        max_ctas, stream);
}

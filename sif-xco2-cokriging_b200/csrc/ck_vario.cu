// ck_vario.cu -- K2: empirical (cross-)semivariogram pair binning for sm_100a.
//
// All pairs are visited tile by tile (VA x VB points per CTA, coordinates prepared once in shared
// memory).  Histograms live in shared memory:
//   counts  one u32 array per CTA, updated with the native warp-aggregating shared-memory atomic (ATOMS.POPC.INC):
//           integer additions are exact and order-free;
//   sums    FP64, one histogram per WARP with VC = 8 private columns ([warp][bin][column]).  The four lanes that share a
//           column (lane, lane + 8, lane + 16, lane + 24) update it in four fixed phases, so every column is built by a
//           fixed sequence of additions -- no atomics, no data-dependent order -- and the eight lanes of a phase hit
//           eight different banks.  Columns, warps and tiles are then combined by fixed-shape trees.
// A quarter of the per-lane-column footprint of round 1 (which capped the kernel at 1 CTA = 8 warps per SM, FP64 pipe 21 %
// active, profiles/r01ab_k2_ncu_full_summary.txt): 50 bins now need 26 KB per CTA, so 5 CTAs (40 warps) share an SM.
// The decomposition depends only on (na, nb, n_bins) -- never on the device or the launch -- hence the FP64 sums are
// bit-reproducible and the integer counts are exact.
// Bin decision = pandas.cut(include_lowest=True) on host-supplied edges: bin k <=> e[k] < d <= e[k+1],
// d == e[0] -> bin 0, d > e[n_bins] dropped (src/fields.py:208-222).
#include "ck_common.cuh"

constexpr int VA = 128;  // rows of A per tile
constexpr int VB = 256;  // rows of B per tile
constexpr int V_MAX_WARPS = 8;
constexpr int VC = 8;                      // private FP64 sum columns per warp histogram (32 / VC update phases)
constexpr int V_SMEM_BUDGET = 200 * 1024;  // histogram bytes per CTA

struct VarioGeom {
  long long ta, tb;  // tiles along A, B
};
static inline VarioGeom vario_geom(long long na, long long nb) {
  VarioGeom g;
  g.ta = (na + VA - 1) / VA;
  g.tb = (nb + VB - 1) / VB;
  return g;
}
// tile-row range [t0, t1) of one call: (0, <0) = all rows of A (single-GPU); multi-GPU ranks pass their share
static inline bool vario_range(const VarioGeom& g, long long* t0, long long* t1) {
  if (*t1 < 0) { *t0 = 0; *t1 = g.ta; }
  return *t0 >= 0 && *t0 <= *t1 && *t1 <= g.ta;
}
static inline int vario_warps(int n_bins) {
  int w = (V_SMEM_BUDGET - n_bins * 4) / (n_bins * VC * 8);
  if (w > V_MAX_WARPS) w = V_MAX_WARPS;
  // power of two so that the warp tree has a fixed shape
  int p = 1;
  while (p * 2 <= w) p *= 2;
  return w >= 1 ? p : 0;
}

__device__ __forceinline__ unsigned long long dbits(double d) { return (unsigned long long)__double_as_longlong(d); }

// Pair distance of the binning kernels.  Euclidean: the reference-order function (bit-identical to scipy cdist, so bin
// decisions need no guard).  Haversine: the branch-free polynomial path of the assembly kernel (ck_math.cuh), <= 4e-15
// relative from the host's libm evaluation for pairs more than ~1000 km from the antipode; every decision inside the
// guard band CK_VARIO_GUARD (5x that) is left to the host, which re-evaluates those pairs with libm.  (Within ~100 km of the
// antipode asin(sqrt(a)) amplifies the last-bit rounding of a beyond any fixed relative band -- for the device AND for
// libm; variogram ranges never reach there: pairs beyond max_dist (1 + guard) are dropped before any decision.)
template <int METRIC>
__device__ __forceinline__ double vario_dist(const CkPoint& p, const CkPoint& q) {
  return METRIC == CK_METRIC_HAVERSINE ? ck_dist_haversine_fast(p, q) : ck_dist_euclid(p, q);
}

// ------------------------------------------------------------------------------------------------
// pass 1: min non-zero distance, max distance and number of pairs with d <= max_dist
// (min / max are order-free, so plain atomics on the bit patterns of non-negative doubles are exact)
template <int METRIC>
__global__ void __launch_bounds__(256) ck_vario_minmax_kernel(const double* __restrict__ xya, long long na,
                                                              const double* __restrict__ xyb, long long nb,
                                                              int same_field, double max_dist,
                                                              unsigned long long* __restrict__ out,
                                                              unsigned long long* __restrict__ tile_mm, long long ty0) {
  const long long a0 = (ty0 + blockIdx.y) * VA, b0 = (long long)blockIdx.x * VB;
  const long long tile = (ty0 + blockIdx.y) * gridDim.x + blockIdx.x;
  if (same_field && b0 + VB - 1 <= a0) {  // tile entirely on/below the diagonal
    if (threadIdx.x == 0) { tile_mm[2 * tile] = 0x7FF0000000000000ULL; tile_mm[2 * tile + 1] = 0ULL; }
    return;
  }
  __shared__ CkPoint pa[VA], pb[VB];
  __shared__ unsigned long long smin[8], smax[8], scnt[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < VA && a0 + t < na) pa[t] = ck_prepare_point(METRIC, xya[2 * (a0 + t)], xya[2 * (a0 + t) + 1]);
  if (b0 + t < nb) pb[t] = ck_prepare_point(METRIC, xyb[2 * (b0 + t)], xyb[2 * (b0 + t) + 1]);
  __syncthreads();
  unsigned long long mn = 0x7FF0000000000000ULL, mx = 0ULL, cnt = 0ULL;
  bool any = false;
  for (int ia = warp; ia < VA; ia += 8) {
    const long long ga = a0 + ia;
    if (ga >= na) break;
    const CkPoint p = pa[ia];
#pragma unroll 4
    for (int q = 0; q < VB / 32; ++q) {
      const int ib = lane + 32 * q;
      const long long gb = b0 + ib;
      if (gb < nb && (!same_field || gb > ga)) {
        const double d = vario_dist<METRIC>(p, pb[ib]);
        if (d <= max_dist) {  // max_dist already widened by the guard band for inexact metrics
          const unsigned long long u = dbits(d);
          ++cnt;
          if (d > 0.0 && u < mn) mn = u;
          if (!any || u > mx) mx = u;
          any = true;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long omn = __shfl_xor_sync(0xffffffffu, mn, o);
    const unsigned long long omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const unsigned long long oc = __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = omn < mn ? omn : mn;
    mx = omx > mx ? omx : mx;
    cnt += oc;
  }
  if (lane == 0) { smin[warp] = mn; smax[warp] = mx; scnt[warp] = cnt; }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < 8; ++w) {
      mn = smin[w] < mn ? smin[w] : mn;
      mx = smax[w] > mx ? smax[w] : mx;
      cnt += scnt[w];
    }
    tile_mm[2 * tile] = mn;
    tile_mm[2 * tile + 1] = cnt ? mx : 0ULL;
    if (cnt) {
      atomicMin(out + 0, mn);
      atomicMax(out + 1, mx);
      atomicAdd(out + 2, cnt);
    }
  }
}

// Pairs whose device distance is within the guard band of the extrema: the host re-evaluates them
// with libm so that bin centres/edges carry exactly the reference's bits.  Only tiles whose own
// extrema reach the bands do any work.
template <int METRIC>
__global__ void __launch_bounds__(256) ck_vario_candidates_kernel(const double* __restrict__ xya, long long na,
                                                                  const double* __restrict__ xyb, long long nb,
                                                                  int same_field, double dlim, double lo, double hi,
                                                                  const unsigned long long* __restrict__ tile_mm,
                                                                  ck_i64* __restrict__ pairs, long long capacity,
                                                                  unsigned long long* __restrict__ count, long long ty0) {
  const long long a0 = (ty0 + blockIdx.y) * VA, b0 = (long long)blockIdx.x * VB;
  const long long tile = (ty0 + blockIdx.y) * gridDim.x + blockIdx.x;
  const double tmn = __longlong_as_double((long long)tile_mm[2 * tile]);
  const double tmx = __longlong_as_double((long long)tile_mm[2 * tile + 1]);
  if (!(tmn <= lo) && !(tmx >= hi)) return;
  __shared__ CkPoint pa[VA], pb[VB];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t < VA && a0 + t < na) pa[t] = ck_prepare_point(METRIC, xya[2 * (a0 + t)], xya[2 * (a0 + t) + 1]);
  if (b0 + t < nb) pb[t] = ck_prepare_point(METRIC, xyb[2 * (b0 + t)], xyb[2 * (b0 + t) + 1]);
  __syncthreads();
  for (int ia = warp; ia < VA; ia += 8) {
    const long long ga = a0 + ia;
    if (ga >= na) break;
    const CkPoint p = pa[ia];
    for (int q = 0; q < VB / 32; ++q) {
      const int ib = lane + 32 * q;
      const long long gb = b0 + ib;
      if (gb < nb && (!same_field || gb > ga)) {
        const double d = vario_dist<METRIC>(p, pb[ib]);
        if (d <= dlim && ((d > 0.0 && d <= lo) || d >= hi)) {
          const unsigned long long pos = atomicAdd(count, 1ULL);
          if ((long long)pos < capacity) { pairs[2 * pos] = ga; pairs[2 * pos + 1] = gb; }
        }
      }
    }
  }
}

__global__ void ck_vario_minmax_init_kernel(unsigned long long* out) {
  out[0] = 0x7FF0000000000000ULL;  // +inf
  out[1] = 0ULL;
  out[2] = 0ULL;
}
__global__ void ck_vario_minmax_final_kernel(unsigned long long* out) {
  const unsigned long long cnt = out[2];
  double* o = reinterpret_cast<double*>(out);
  if (cnt == 0) o[1] = -__longlong_as_double(0x7FF0000000000000LL);
  o[2] = (double)cnt;
}

// relative half-width of the guard band around decision boundaries (~90 ulp, 5x the measured device-vs-libm deviation of
// the fast haversine); haversine only -- Euclidean distances are bit-identical to scipy's and need no guard
#define CK_VARIO_GUARD 2.0e-14
static inline double vario_dlim(int metric, double max_dist) {
  return metric == CK_METRIC_HAVERSINE ? max_dist * (1.0 + CK_VARIO_GUARD) : max_dist;
}

extern "C" size_t ck_vario_minmax_workspace_bytes(ck_i64 na, ck_i64 nb) {
  if (na <= 0 || nb <= 0) return 256;
  const VarioGeom g = vario_geom(na, nb);
  return (size_t)g.ta * g.tb * 16 + 256;
}

extern "C" ck_i64 ck_vario_tile_rows(ck_i64 na) { return na > 0 ? (na + VA - 1) / VA : 0; }
extern "C" ck_i64 ck_vario_tile_cols(ck_i64 nb) { return nb > 0 ? (nb + VB - 1) / VB : 0; }
extern "C" int ck_vario_tile_shape(int* va, int* vb) { if (va) *va = VA; if (vb) *vb = VB; return CK_OK; }

extern "C" int ck_vario_minmax(const double* xya, ck_i64 na, const double* xyb, ck_i64 nb, int metric, int same_field,
                               double max_dist, ck_i64 tile_row_begin, ck_i64 tile_row_end, double* out, void* ws,
                               void* stream) {
  CK_REQUIRE(na >= 0 && nb >= 0 && out, "bad argument");
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  CK_REQUIRE(!same_field || na == nb, "same_field needs na == nb");
  cudaStream_t st = ck_stream(stream);
  unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
  ck_vario_minmax_init_kernel<<<1, 1, 0, st>>>(o);
  if (na > 0 && nb > 0) {
    CK_REQUIRE(xya && xyb && ws, "null pointer");
    const VarioGeom g = vario_geom(na, nb);
    long long t0 = tile_row_begin, t1 = tile_row_end;
    CK_REQUIRE(vario_range(g, &t0, &t1), "bad tile-row range [%lld, %lld) of %lld", (long long)tile_row_begin, (long long)tile_row_end, g.ta);
    CK_REQUIRE(t1 - t0 <= 65535, "na too large");
    if (t1 > t0) {
      dim3 grid((unsigned)g.tb, (unsigned)(t1 - t0));
      unsigned long long* mm = static_cast<unsigned long long*>(ws);
      const double dlim = vario_dlim(metric, max_dist);
      if (metric == CK_METRIC_HAVERSINE) ck_vario_minmax_kernel<CK_METRIC_HAVERSINE><<<grid, 256, 0, st>>>(xya, na, xyb, nb, same_field, dlim, o, mm, t0);
      else ck_vario_minmax_kernel<CK_METRIC_EUCLID><<<grid, 256, 0, st>>>(xya, na, xyb, nb, same_field, dlim, o, mm, t0);
      CK_LAUNCH_CHECK();
    }
  }
  ck_vario_minmax_final_kernel<<<1, 1, 0, st>>>(o);
  CK_LAUNCH_CHECK_N(2);
  return CK_OK;
}

extern "C" int ck_vario_candidates(const double* xya, ck_i64 na, const double* xyb, ck_i64 nb, int metric, int same_field,
                                   double max_dist, double lo, double hi, ck_i64 tile_row_begin, ck_i64 tile_row_end,
                                   const void* ws, ck_i64* pairs, ck_i64 capacity, unsigned long long* count,
                                   void* stream) {
  CK_REQUIRE(na >= 0 && nb >= 0 && pairs && count && capacity >= 0, "bad argument");
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  cudaStream_t st = ck_stream(stream);
  CK_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
  if (na == 0 || nb == 0) return CK_OK;
  CK_REQUIRE(xya && xyb && ws, "null pointer");
  const VarioGeom g = vario_geom(na, nb);
  long long t0 = tile_row_begin, t1 = tile_row_end;
  CK_REQUIRE(vario_range(g, &t0, &t1), "bad tile-row range [%lld, %lld) of %lld", (long long)tile_row_begin, (long long)tile_row_end, g.ta);
  if (t1 == t0) return CK_OK;
  dim3 grid((unsigned)g.tb, (unsigned)(t1 - t0));
  const unsigned long long* mm = static_cast<const unsigned long long*>(ws);
  const double dlim = vario_dlim(metric, max_dist);
  if (metric == CK_METRIC_HAVERSINE) ck_vario_candidates_kernel<CK_METRIC_HAVERSINE><<<grid, 256, 0, st>>>(xya, na, xyb, nb, same_field, dlim, lo, hi, mm, pairs, capacity, count, t0);
  else ck_vario_candidates_kernel<CK_METRIC_EUCLID><<<grid, 256, 0, st>>>(xya, na, xyb, nb, same_field, dlim, lo, hi, mm, pairs, capacity, count, t0);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

// ------------------------------------------------------------------------------------------------
// pass 2: binning
struct VarioBinArgs {
  const double* xya; const double* va; long long na; double mean_a;
  const double* xyb; const double* vb; long long nb; double mean_b;
  int same_field, covariogram;
  double max_dist;
  const double* edges;  // device copy inside the workspace
  int n_bins;
  double e1, inv_w;     // uniform-width guess: k = (d - e1) * inv_w + 1
  double* tile_sums;                // [tile][bin]
  unsigned int* tile_counts;        // [tile][bin]
  // pairs within the guard band of an edge / of max_dist are not binned but appended here (host decides)
  ck_i64* flagged; long long flag_capacity; unsigned long long* flag_count;
  double guard, dlim;
};

template <int METRIC>
__global__ void __launch_bounds__(256) ck_vario_bin_kernel(VarioBinArgs g, int nwarps, long long ty0) {
  extern __shared__ __align__(16) unsigned char vsm[];
  const int nb_ = g.n_bins;
  double* hsum = reinterpret_cast<double*>(vsm);                                  // [warp][bin][column]
  double* edges = hsum + (size_t)nwarps * nb_ * VC;                               // n_bins + 1
  unsigned int* hcnt = reinterpret_cast<unsigned int*>(edges + nb_ + 1);          // [bin], one per CTA
  __shared__ CkPoint pa[VA], pb[VB];
  __shared__ double ra[VA], rb[VB];
  const long long a0 = (ty0 + blockIdx.y) * VA, b0 = (long long)blockIdx.x * VB;
  const long long tile = (ty0 + blockIdx.y) * gridDim.x + blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nthreads = nwarps * 32;
  const bool skip = g.same_field && (b0 + VB - 1 <= a0);
  if (skip) {  // still publish zeros so that the tile tree reads defined values
    for (int k = t; k < nb_; k += nthreads) {
      g.tile_sums[tile * nb_ + k] = 0.0;
      g.tile_counts[tile * nb_ + k] = 0u;
    }
    return;
  }
  for (int i = t; i < nwarps * nb_ * VC; i += nthreads) hsum[i] = 0.0;
  for (int i = t; i < nb_; i += nthreads) hcnt[i] = 0u;
  for (int i = t; i <= nb_; i += nthreads) edges[i] = g.edges[i];
  for (int i = t; i < VA; i += nthreads)
    if (a0 + i < g.na) {
      pa[i] = ck_prepare_point(METRIC, g.xya[2 * (a0 + i)], g.xya[2 * (a0 + i) + 1]);
      ra[i] = g.va[a0 + i] - g.mean_a;
    }
  for (int i = t; i < VB; i += nthreads)
    if (b0 + i < g.nb) {
      pb[i] = ck_prepare_point(METRIC, g.xyb[2 * (b0 + i)], g.xyb[2 * (b0 + i) + 1]);
      rb[i] = g.vb[b0 + i] - g.mean_b;
    }
  __syncthreads();
  double* mysum = hsum + (size_t)warp * nb_ * VC + (lane & (VC - 1));
  const int myphase = lane / VC;
  const double e_last = edges[nb_], e_first = edges[0];
  for (int ia = warp; ia < VA; ia += nwarps) {
    const long long ga = a0 + ia;
    if (ga >= g.na) break;
    const CkPoint p = pa[ia];
    const double r = ra[ia];
#pragma unroll 4
    for (int q = 0; q < VB / 32; ++q) {
      const int ib = lane + 32 * q;
      const long long gb = b0 + ib;
      int k = -1;  // bin of this lane's pair, -1: nothing to add
      double v = 0.0;
      if (gb < g.nb && (!g.same_field || gb > ga)) {
        const double d = vario_dist<METRIC>(p, pb[ib]);
        if (d <= g.dlim && d >= e_first) {
          k = (int)((d - g.e1) * g.inv_w) + 1;
          k = k < 0 ? 0 : (k > nb_ - 1 ? nb_ - 1 : k);
          while (k > 0 && d <= edges[k]) --k;
          while (k < nb_ - 1 && d > edges[k + 1]) ++k;
          bool keep = d <= g.max_dist && d <= e_last;
          if (g.flagged) {
            const double band = g.guard * d;
            const bool near = fabs(d - g.max_dist) <= g.guard * g.max_dist || (k > 0 && d - edges[k] <= band) ||
                              fabs(edges[k + 1] - d) <= band;
            if (near) {
              const unsigned long long pos = atomicAdd(g.flag_count, 1ULL);
              if ((long long)pos < g.flag_capacity) { g.flagged[2 * pos] = ga; g.flagged[2 * pos + 1] = gb; }
              keep = false;
            }
          }
          if (keep) {
            if (g.covariogram) v = r * rb[ib];
            else { const double df = r - rb[ib]; v = 0.5 * (df * df); }
          } else {
            k = -1;
          }
        }
      }
      if (k >= 0) atomicAdd(hcnt + k, 1u);  // ATOMS.POPC.INC: exact, order-free
      // the 32 / VC lanes that share a sum column take turns, in a fixed order
#pragma unroll
      for (int ph = 0; ph < 32 / VC; ++ph)
        if (myphase == ph && k >= 0) mysum[k * VC] += v;
    }
  }
  __syncthreads();
  // columns -> (fixed xor tree over VC lanes), warps -> (fixed tree), one thread writes the tile partial.
  // A warp handles 32 / VC bins per pass: lane = (bin slot, column).
  constexpr int BPW = 32 / VC;
  for (int k0 = warp * BPW; k0 < nb_; k0 += nwarps * BPW) {
    const int k = k0 + lane / VC, col = lane & (VC - 1);
    double s[V_MAX_WARPS];
    for (int w = 0; w < nwarps; ++w) {
      double sv = (k < nb_) ? hsum[((size_t)w * nb_ + k) * VC + col] : 0.0;
#pragma unroll
      for (int o = VC / 2; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      s[w] = sv;
    }
    for (int h = 1; h < nwarps; h <<= 1)
      for (int w = 0; w + h < nwarps; w += 2 * h) s[w] += s[w + h];
    if (col == 0 && k < nb_) {
      g.tile_sums[tile * nb_ + k] = s[0];
      g.tile_counts[tile * nb_ + k] = hcnt[k];
    }
  }
}

// one CTA per bin: strided per-thread partials over tiles (fixed order), shared-memory tree
__global__ void __launch_bounds__(256) ck_vario_tile_reduce_kernel(const double* __restrict__ tile_sums,
                                                                   const unsigned int* __restrict__ tile_counts,
                                                                   long long ntiles, int n_bins,
                                                                   unsigned long long* __restrict__ counts,
                                                                   double* __restrict__ sums) {
  __shared__ double rs[256];
  __shared__ unsigned long long rc[256];
  const int k = blockIdx.x, t = threadIdx.x;
  double s = 0.0;
  unsigned long long c = 0ULL;
  for (long long i = t; i < ntiles; i += 256) {
    s += tile_sums[i * n_bins + k];
    c += tile_counts[i * n_bins + k];
  }
  rs[t] = s;
  rc[t] = c;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (t < h) {
      rs[t] += rs[t + h];
      rc[t] += rc[t + h];
    }
    __syncthreads();
  }
  if (t == 0) {
    sums[k] = rs[0];
    counts[k] = rc[0];
  }
}

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t ck_vario_bin_workspace_bytes(ck_i64 na, ck_i64 nb, int n_bins) {
  if (na <= 0 || nb <= 0 || n_bins <= 0) return 256;
  const VarioGeom g = vario_geom(na, nb);
  const size_t tiles = (size_t)g.ta * g.tb;
  return align256((size_t)(n_bins + 1) * 8) + align256(tiles * n_bins * 8) + align256(tiles * n_bins * 4);
}

// workspace layout: [edges (n_bins+1 doubles)] [tile_sums: tiles x n_bins doubles] [tile_counts: tiles x n_bins u32]
struct VarioWs {
  double* edges; double* tile_sums; unsigned int* tile_counts; size_t tiles;
};
static inline VarioWs vario_ws(void* ws, ck_i64 na, ck_i64 nb, int n_bins) {
  const VarioGeom geo = vario_geom(na, nb);
  VarioWs w;
  w.tiles = (size_t)geo.ta * geo.tb;
  unsigned char* b = static_cast<unsigned char*>(ws);
  w.edges = reinterpret_cast<double*>(b);
  w.tile_sums = reinterpret_cast<double*>(b + align256((size_t)(n_bins + 1) * 8));
  w.tile_counts = reinterpret_cast<unsigned int*>(b + align256((size_t)(n_bins + 1) * 8) + align256(w.tiles * n_bins * 8));
  return w;
}

extern "C" size_t ck_vario_bin_partials_offset(int n_bins) { return align256((size_t)(n_bins + 1) * 8); }

extern "C" int ck_vario_bin_tiles(const double* xya, const double* va, ck_i64 na, double mean_a, const double* xyb,
                                  const double* vb, ck_i64 nb, double mean_b, int metric, int same_field, int covariogram,
                                  double max_dist, const double* edges, int n_bins, ck_i64 tile_row_begin,
                                  ck_i64 tile_row_end, ck_i64* flagged, ck_i64 flag_capacity,
                                  unsigned long long* flag_count, void* ws, void* stream) {
  CK_REQUIRE(na > 0 && nb > 0, "empty point set");
  CK_REQUIRE(n_bins >= 1 && edges && ws, "bad argument");
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  CK_REQUIRE(!same_field || na == nb, "same_field needs na == nb");
  for (int k = 0; k < n_bins; ++k) CK_REQUIRE(edges[k] < edges[k + 1], "edges must be strictly ascending");
  const int nwarps = vario_warps(n_bins);
  if (nwarps < 1) { ck_set_error("n_bins=%d exceeds the shared-memory histogram capacity", n_bins); return CK_ERR_UNSUPPORTED; }
  cudaStream_t st = ck_stream(stream);
  CK_REQUIRE(!flagged || (flag_count && flag_capacity >= 0), "flag_count is NULL");
  if (flag_count) CK_CUDA(cudaMemsetAsync(flag_count, 0, sizeof(unsigned long long), st));
  CK_REQUIRE(xya && va && xyb && vb, "null pointer");
  const VarioGeom geo = vario_geom(na, nb);
  long long t0 = tile_row_begin, t1 = tile_row_end;
  CK_REQUIRE(vario_range(geo, &t0, &t1), "bad tile-row range [%lld, %lld) of %lld", (long long)tile_row_begin, (long long)tile_row_end, geo.ta);
  CK_REQUIRE(t1 - t0 <= 65535, "na too large");
  const VarioWs w = vario_ws(ws, na, nb, n_bins);
  // partials of tile rows outside [t0, t1) are ZERO: other ranks own them and the cross-rank combine is a
  // sum in which every partial has exactly one non-zero contributor (exact, order-free)
  CK_CUDA(cudaMemsetAsync(w.tile_sums, 0, w.tiles * n_bins * sizeof(double), st));
  CK_CUDA(cudaMemsetAsync(w.tile_counts, 0, w.tiles * n_bins * sizeof(unsigned int), st));
  CK_CUDA(cudaMemcpyAsync(w.edges, edges, sizeof(double) * (n_bins + 1), cudaMemcpyHostToDevice, st));
  if (t1 == t0) return CK_OK;
  VarioBinArgs g;
  g.xya = xya; g.va = va; g.na = na; g.mean_a = mean_a;
  g.xyb = xyb; g.vb = vb; g.nb = nb; g.mean_b = mean_b;
  g.same_field = same_field; g.covariogram = covariogram; g.max_dist = max_dist;
  g.edges = w.edges; g.n_bins = n_bins;
  g.e1 = n_bins >= 2 ? edges[1] : edges[0];
  const double wdt = n_bins >= 2 ? (edges[n_bins] - edges[1]) / (double)(n_bins - 1) : 0.0;
  g.inv_w = wdt > 0.0 ? 1.0 / wdt : 0.0;
  g.tile_sums = w.tile_sums; g.tile_counts = w.tile_counts;
  const bool guard = flagged && metric == CK_METRIC_HAVERSINE;  // Euclidean distances are exact: no guard
  g.flagged = guard ? flagged : nullptr; g.flag_capacity = flag_capacity; g.flag_count = flag_count;
  g.guard = CK_VARIO_GUARD; g.dlim = guard ? vario_dlim(metric, max_dist) : max_dist;
  const size_t smem = (size_t)nwarps * n_bins * VC * 8 + (size_t)(n_bins + 1) * 8 + (size_t)n_bins * 4 + 16;
  dim3 grid((unsigned)geo.tb, (unsigned)(t1 - t0));
  if (metric == CK_METRIC_HAVERSINE) {
    CK_CUDA(cudaFuncSetAttribute(ck_vario_bin_kernel<CK_METRIC_HAVERSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ck_vario_bin_kernel<CK_METRIC_HAVERSINE><<<grid, nwarps * 32, smem, st>>>(g, nwarps, t0);
  } else {
    CK_CUDA(cudaFuncSetAttribute(ck_vario_bin_kernel<CK_METRIC_EUCLID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ck_vario_bin_kernel<CK_METRIC_EUCLID><<<grid, nwarps * 32, smem, st>>>(g, nwarps, t0);
  }
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_vario_bin_reduce(ck_i64 na, ck_i64 nb, int n_bins, const void* ws, unsigned long long* counts,
                                   double* sums, void* stream) {
  CK_REQUIRE(na > 0 && nb > 0 && n_bins >= 1 && ws && counts && sums, "bad argument");
  const VarioWs w = vario_ws(const_cast<void*>(ws), na, nb, n_bins);
  ck_vario_tile_reduce_kernel<<<n_bins, 256, 0, ck_stream(stream)>>>(w.tile_sums, w.tile_counts, (long long)w.tiles, n_bins,
                                                                    counts, sums);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_vario_bin(const double* xya, const double* va, ck_i64 na, double mean_a, const double* xyb,
                            const double* vb, ck_i64 nb, double mean_b, int metric, int same_field, int covariogram,
                            double max_dist, const double* edges, int n_bins, unsigned long long* counts, double* sums,
                            ck_i64* flagged, ck_i64 flag_capacity, unsigned long long* flag_count, void* ws,
                            void* stream) {
  CK_REQUIRE(na >= 0 && nb >= 0, "negative size");
  CK_REQUIRE(n_bins >= 1 && edges && counts && sums && ws, "bad argument");
  if (na == 0 || nb == 0) {
    cudaStream_t st = ck_stream(stream);
    if (flag_count) CK_CUDA(cudaMemsetAsync(flag_count, 0, sizeof(unsigned long long), st));
    CK_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * n_bins, st));
    CK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * n_bins, st));
    return CK_OK;
  }
  int rc = ck_vario_bin_tiles(xya, va, na, mean_a, xyb, vb, nb, mean_b, metric, same_field, covariogram, max_dist, edges,
                              n_bins, 0, -1, flagged, flag_capacity, flag_count, ws, stream);
  if (rc) return rc;
  return ck_vario_bin_reduce(na, nb, n_bins, ws, counts, sums, stream);
}

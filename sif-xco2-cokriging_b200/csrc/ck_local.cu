// ck_local.cu -- K4: batched local-neighbourhood (point) cokriging for sm_100a.
//
// Replaces the reference's per-target Python loop (src/point_prediction.py:127-249: boolean neighbour masks, np.ix_
// gathers from the stored Sigma blocks, two k x k Cholesky factorisations per target) by two launches for ALL targets:
//
//   ck_local_count_kernel    neighbour counts per target, split into the eight (process, scan segment) pieces that the
//                            second kernel needs to compact the neighbours IN THE REFERENCE'S ORDER (process 0 then 1,
//                            index order = the order of the boolean masks) without re-counting;
//   ck_local_predict_kernel  persistent CTAs of 160 threads (3 per SM, one target each, targets handed out by an
//                            atomic counter): FOUR UPDATE WARPS + ONE DIAGONAL-BLOCK WARP.
//     1. scan + ordered compaction (update warps).  A conservative bounding test (|dx|, |dy| for Euclid; latitude band and
//        a longitude bound for haversine, margins 1e-12 / 1e-9 >> the rounding of the distance) rejects most data with
//        a handful of instructions; survivors get the reference-order distance (bit-identical Euclid, libm-level
//        haversine), so the neighbour sets equal the reference's.  The kept points (or their joint-matrix indices), their
//        covariance with the target (c) and their data (z) go to shared memory.
//     2. LEFT-looking blocked Cholesky in 32-column block columns on the (kp + 2)-row array  [Sigma_loc ; c^T ; z^T]
//        (kp = k rounded up to 32, identity padded).  For block column J and every 64-row tile:
//            W = Sigma[tile, J] - L[tile, 0:J) L[J, 0:J)^T
//        The Sigma entries are gathered from the stored joint covariance or RE-COMPUTED from the coordinates straight
//        into the FP64 DMMA accumulator registers (no k x k matrix is ever stored: the only array in memory is L itself);
//        the product is FP64 tensor-core DMMA (mma.sync m8n8k4) fed by a cp.async-staged pipeline from the L2-resident
//        workspace slot of the CTA.  The 32 x 32 diagonal block of W goes to shared memory, where the DIAG WARP eliminates
//        it with register-resident rows and shuffles (pivot reciprocal by Newton, square roots off the chain) and forms
//        X_d = L_d^-1 by forward substitution -- ~30 k cycles of dependent FP64 work that used to idle the other three
//        warps at a barrier (a quarter of the CTA's time, profiles/r02_k4_variants_phases.log) and now overlaps the update
//        warps' main loops over the remaining row tiles, whose W is parked in place in the factor array.  Once X_d is
//        published (named barriers) every update warp turns the rows it parked into  P = W X_d^T  (DMMA, skipping the zero
//        half of X_d).  Because c and z ride along as two extra rows, the sweep leaves v = L^-1 c and y = L^-1 z.
//     3. pred = v . y,  sd = sqrt(max(c0 - v . v, 0))   (src/point_prediction.py:200-222); NaN for an empty or
//        non-positive-definite neighbourhood like the reference.
//
// Matern evaluation: when the three blocks of the joint model share one closed-form order nu in {1/2, 3/2, 5/2, 7/2}
// (every BASELINE config) the matrix entries use the branch-free polynomial path of K1 (<= 2 ulp per piece,
// ck_math.cuh); otherwise the reference-order generic path (K_nu).  The target-to-neighbour vector c always uses the
// reference-order path on the exact selection distances.
#include "ck_common.cuh"

constexpr int LP = 32;             // block-column (panel) width
#ifndef CK_LOCAL_TM
#define CK_LOCAL_TM 64
#endif
constexpr int LTM = CK_LOCAL_TM;   // rows per tile (64 or 128)
constexpr int LWR = LTM / 4;       // rows per update warp (16 or 32)
constexpr int LMI = LWR / 8;       // 8-row DMMA tiles per warp
constexpr int LK = 16;             // k-depth per pipeline stage
constexpr int LLD = LK + 4;        // staging row stride (doubles): % 16 == 4 -> conflict-free DMMA fragment loads
#ifndef CK_LOCAL_STAGES
#define CK_LOCAL_STAGES 3  // measured (profiles/r02_k4_shapes.log): 2..3 stages, 16 / 32 deep, 3 / 4 CTAs per SM all land within 4 %
#endif
constexpr int LSTG = CK_LOCAL_STAGES;  // pipeline stages
constexpr int LWLD = LP + 4;       // row stride of the W tile and of X_d
constexpr int L_UPD_THREADS = 128;  // four update warps: scan, covariance entries, DMMA main loops, panel products
constexpr int L_WARPS = L_UPD_THREADS / 32;
constexpr int L_THREADS = L_UPD_THREADS + 32;  // + one warp that only factors the 32 x 32 diagonal blocks
constexpr int L_STAGE_ELEMS = (LTM + LP) * LLD;
constexpr int L_SMEM_KMAX = 2048;  // neighbour records live in shared memory up to this k, in the workspace beyond
constexpr int L_FAST_NONE = -1;    // FAST template values: 0..3 closed-form nu (polynomial path), -1 generic, -2 gather
constexpr int L_GATHER = -2;
#ifndef CK_LOCAL_MIN_CTAS
#define CK_LOCAL_MIN_CTAS 3
#endif
constexpr int L_MIN_CTAS = CK_LOCAL_MIN_CTAS;  // CTAs per SM the register allocation is bounded for

struct LocalArgs {
  const double* xy[2]; const double* z[2]; long long n[2];
  const double* xyp; long long m;
  CkMatern S[3];    // joint-model blocks 00, 01, 11 (frozen at Predictor construction, src/point_prediction.py:42)
  CkMatern C[2];    // target-to-process-j blocks for the predicted process (own block carries the nugget)
  int n_procs, i_pred, metric, cv;
  double max_dist, c0;
  const double* sigma; long long ld_sigma;  // optional precomputed joint covariance (gather mode)
  const int* kcount; const int* kseg; long long kmax, kmaxp;
  double* pred; double* sd; int* info;
  double* ws;       // slots x [(kmaxp + 2) x kmaxp factor rows | optional neighbour records]
  long long slot_stride;
  int pts_in_ws;
  unsigned int* counter;
  long long* dbg;   // optional phase cycle counters (ck_local_debug_buffer): scan, init, main loop, diagonal block, panel, rest
};

// out-of-line so that the five Matern variants are instantiated once per kernel, not per call site
static __device__ __noinline__ double cov_dyn(const CkMatern& P, double h) { return ck_matern_cov_dyn(P, h); }

__device__ __forceinline__ void l_cp_async16(unsigned dst, const void* src, unsigned on) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.cg.shared.global [%0], [%1], 16;\n}\n" ::"r"(dst), "l"(src),
               "r"(on));
}
__device__ __forceinline__ void l_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void l_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void l_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// neighbour scan
// ------------------------------------------------------------------------------------------------
// per-target constants of the conservative pre-filter
struct LocalFilter {
  double p0, p1;     // raw target coordinates
  double lim0;       // Euclid: max_dist (1 + 1e-12);  haversine: latitude band in degrees (or inf)
  double lon_k;      // haversine: reject when |x_red| * lon_k > 1  (0: longitude test off)
};

template <int METRIC>
__device__ __forceinline__ LocalFilter local_filter(const LocalArgs& g, double p0, double p1) {
  LocalFilter f;
  f.p0 = p0; f.p1 = p1;
  const double inf = __longlong_as_double(0x7FF0000000000000LL);
  if (METRIC == CK_METRIC_EUCLID) {
    f.lim0 = g.max_dist * (1.0 + 1.0e-12);
    f.lon_k = 0.0;
  } else {
    // haversine >= R |dlat| and >= 2 R asin(sqrt(cos(lat1) cos(lat2)) |sin(dlon / 2)|); |sin x| >= (2 / pi) |x| on
    // [-pi/2, pi/2].  Margins 1e-9 relative (the computed distance is good to ~1e-15).
    const double md = g.max_dist * (1.0 + 1.0e-9);
    const double delta = md / CK_EARTH_RADIUS;  // radians
    const bool lat_ok = fabs(p0) <= 90.0 && md >= 0.0;
    f.lim0 = lat_ok ? delta / CK_DEG2RAD : inf;
    f.lon_k = 0.0;
    const double phi = fabs(p0) * CK_DEG2RAD + delta;
    if (lat_ok && 0.5 * delta < 1.5707963267948966 && phi < 1.5707963267948966) {
      const double cmin = sqrt(cos(p0 * CK_DEG2RAD) * cos(phi));
      const double smax = sin(0.5 * delta) * (1.0 + 1.0e-9);
      if (cmin > 0.0 && smax > 0.0) f.lon_k = 0.6366197723675814 * cmin / smax;  // (2 / pi) cmin / smax
    }
  }
  return f;
}

// true: the datum cannot be within max_dist (decided without the exact distance)
template <int METRIC>
__device__ __forceinline__ bool local_reject(const LocalFilter& f, double q0, double q1) {
  if (METRIC == CK_METRIC_EUCLID) return fabs(q0 - f.p0) > f.lim0 || fabs(q1 - f.p1) > f.lim0;
  if (fabs(q0 - f.p0) > f.lim0 && fabs(q0) <= 90.0) return true;
  if (f.lon_k > 0.0 && fabs(q0) <= 90.0) {
    const double x = 0.5 * CK_DEG2RAD * (q1 - f.p1);
    const double xr = x - 3.141592653589793 * rint(x * 0.3183098861837907);
    return fabs(xr) * f.lon_k > 1.0;
  }
  return false;
}

__device__ __forceinline__ bool local_keep(const LocalArgs& g, int proc, double d) {
  if (!(d <= g.max_dist)) return false;
  if (g.cv && proc == g.i_pred && !(d > 0.0)) return false;
  return true;
}

__device__ __forceinline__ long long local_seglen(long long n) { return ((n + L_WARPS - 1) / L_WARPS + 31) / 32 * 32; }

// One warp scans its segment of process `proc` in index order.  STORE: kept points are appended at `off` (warp-private,
// ordered); returns the number kept.
template <int METRIC, bool STORE, bool GATHER>
__device__ __forceinline__ int local_scan_segment(const LocalArgs& g, const LocalFilter& f, const CkPoint& p0, int proc, int warp,
                                                  int lane, int off, double* pa, double* pb, double* pc, double* pcv, double* pzv) {
  const long long n = g.n[proc], sl = local_seglen(n);
  const long long b0 = (long long)warp * sl, b1 = (b0 + sl < n) ? b0 + sl : n;
  const double* xy = g.xy[proc];
  int cnt = 0;
  for (long long base = b0; base < b1; base += 32) {
    const long long i = base + lane;
    bool keep = false;
    double d = 0.0;
    CkPoint q;
    q.a = q.b = q.c = 0.0;
    if (i < b1) {
      const double2 v = *reinterpret_cast<const double2*>(xy + 2 * i);
      if (!local_reject<METRIC>(f, v.x, v.y)) {
        q = ck_prepare_point(METRIC, v.x, v.y);
        d = ck_dist<METRIC>(p0, q);
        keep = local_keep(g, proc, d);
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (STORE && keep) {
      const int pos = off + cnt + __popc(mask & ((1u << lane) - 1u));
      if (GATHER) {
        reinterpret_cast<int*>(pa)[pos] = (int)(proc ? g.n[0] + i : i);  // row / column of the joint matrix
      } else {
        pa[pos] = q.a;
        pb[pos] = q.b;
        if (METRIC == CK_METRIC_HAVERSINE) pc[pos] = q.c;
      }
      pcv[pos] = cov_dyn(g.C[proc], d);
      pzv[pos] = g.z[proc][i];
    }
    cnt += __popc(mask);
  }
  return cnt;
}

template <int METRIC>
__global__ void __launch_bounds__(L_UPD_THREADS) ck_local_count_kernel(const __grid_constant__ LocalArgs g, int* __restrict__ kout, int* __restrict__ kseg) {
  __shared__ int seg[2 * L_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long c = blockIdx.x; c < g.m; c += gridDim.x) {
    const double t0 = g.xyp[2 * c], t1 = g.xyp[2 * c + 1];
    const CkPoint p0 = ck_prepare_point(METRIC, t0, t1);
    const LocalFilter f = local_filter<METRIC>(g, t0, t1);
    for (int proc = 0; proc < 2; ++proc) {
      int cnt = 0;
      if (proc < g.n_procs)
        cnt = local_scan_segment<METRIC, false, false>(g, f, p0, proc, warp, lane, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
      if (lane == 0) seg[proc * L_WARPS + warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 2 * L_WARPS) kseg[c * (2 * L_WARPS) + threadIdx.x] = seg[threadIdx.x];
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int s = 0; s < 2 * L_WARPS; ++s) tot += seg[s];
      kout[c] = tot;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// factor + solve
// ------------------------------------------------------------------------------------------------
// Entry (i, j) of the block [Sigma_loc ; c ; z] padded with the identity to kp (i >= j inside the matrix part).
template <int METRIC, int FAST>
__device__ __forceinline__ double local_cov_entry(const LocalArgs& g, int i, int j, int k0, const double* pa, const double* pb,
                                                  const double* pc) {
  if (FAST == L_GATHER) {  // j <= i in the local order implies column <= row in the joint matrix: the lower triangle
    const int* gi = reinterpret_cast<const int*>(pa);
    return __ldg(g.sigma + (long long)gi[i] * g.ld_sigma + gi[j]);
  }
  CkPoint pi, pj;
  pi.a = pa[i]; pi.b = pb[i];
  pj.a = pa[j]; pj.b = pb[j];
  pi.c = pj.c = 0.0;
  if (METRIC == CK_METRIC_HAVERSINE) { pi.c = pc[i]; pj.c = pc[j]; }
  // j <= i and process 0 comes first: (proc_j, proc_i) in {(0,0), (0,1), (1,1)}; rows of the reference block come from
  // the lower process id (src/point_prediction.py:159-179)
  const int blk = (i < k0) ? 0 : (j < k0 ? 1 : 2);
  if (FAST >= 0) return ck_matern_cov_fast<(FAST >= 0 ? FAST : 0)>(g.S[blk], ck_dist_fast<METRIC>(pj, pi));
  return cov_dyn(g.S[blk], ck_dist<METRIC>(pj, pi));
}

template <int METRIC, int FAST>
__device__ __forceinline__ double local_entry(const LocalArgs& g, int i, int j, int k, int kp, int k0, const double* pa,
                                              const double* pb, const double* pc, const double* pcv, const double* pzv) {
  if (i < kp) {
    if (j > i) return 0.0;
    if (i >= k) return (i == j) ? 1.0 : 0.0;
    return local_cov_entry<METRIC, FAST>(g, i, j, k0, pa, pb, pc);
  }
  if (j >= k) return 0.0;
  return (i == kp) ? pcv[j] : ((i == kp + 1) ? pzv[j] : 0.0);
}

// 32 x 32 diagonal block (one warp; out of line so that its 64 + 64 registers of row / column state are allocated
// separately from the tile code): Wt rows 0..31 hold the lower triangle of D on entry (D = L_d L_d^T) and, on exit,
// Xd receives X_d = L_d^-1 (explicit zeros above the diagonal); Xd may be the same buffer as Wt.
// Returns 0 or 1 + the first column whose pivot is not > 0.
static __device__ __noinline__ int local_diag_block(double* Wt, double* Xd, double* rsd, int lane) {
  // ---- elimination, lane = row; a[kk] holds the UNSCALED column entry until column kk is finished: only 1 / d sits on
  // the pivot chain (approximate reciprocal + two Newton steps), the 32 square roots are taken after the loop
  double a[LP];
#pragma unroll
  for (int kk = 0; kk < LP; ++kk) a[kk] = Wt[lane * LWLD + kk];
  int first_bad = 0;
  double dk = 1.0;
#pragma unroll
  for (int kk = 0; kk < LP; ++kk) {
    const double d = __shfl_sync(0xffffffffu, a[kk], kk);
    if (!(d > 0.0) && first_bad == 0) first_bad = kk + 1;
    if (lane == kk) dk = d;
    double invd;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(invd) : "d"(d));
    invd = fma(fma(-d, invd, 1.0), invd, invd);
    invd = fma(fma(-d, invd, 1.0), invd, invd);
    const double u = a[kk];
    const double w = u * invd;
#pragma unroll
    for (int jj = kk + 1; jj < LP; ++jj) {
      const double ujk = __shfl_sync(0xffffffffu, u, jj);
      a[jj] = fma(-w, ujk, a[jj]);
    }
  }
  const double sqv = sqrt(dk);
  const double rsv = 1.0 / sqv;
  rsd[lane] = rsv;
#pragma unroll
  for (int kk = 0; kk < LP; ++kk) {
    const double rs = __shfl_sync(0xffffffffu, rsv, kk);
    if (kk <= lane) Wt[lane * LWLD + kk] = (lane == kk) ? sqv : a[kk] * rs;
  }
  __syncwarp();
  // ---- X_d = L_d^-1, lane = column: forward substitution on e_lane, L_d read by broadcast
  double x[LP];
#pragma unroll
  for (int i = 0; i < LP; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int kk = 0; kk < LP; ++kk) {
    x[kk] *= rsd[kk];
#pragma unroll
    for (int i = kk + 1; i < LP; ++i) x[i] = fma(-Wt[i * LWLD + kk], x[kk], x[i]);
  }
  __syncwarp();  // every read of L_d is complete before X_d overwrites the buffer
#pragma unroll
  for (int i = 0; i < LP; ++i) Xd[i * LWLD + lane] = x[i];
  return first_bad;
}

// named barriers (id 0 = __syncthreads over all five warps)
#define L_BAR_UPD 1  // the four update warps among themselves
#define L_BAR_D 2    // update warps arrive: the diagonal block of this block column is in shared memory; the diag warp waits
#define L_BAR_X 3    // the diag warp arrives: X_d is in shared memory; the update warps wait before their panel products
__device__ __forceinline__ void l_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void l_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int METRIC, int FAST>
__global__ void __launch_bounds__(L_THREADS, L_MIN_CTAS) ck_local_predict_kernel(const __grid_constant__ LocalArgs g) {
  extern __shared__ __align__(16) double lsm[];
  double* stage = lsm;                                  // LSTG x (LTM + LP) x LLD
  double* Xd = stage + LSTG * L_STAGE_ELEMS;            // LP x LWLD: the diagonal block D, then X_d = L_d^-1
  double* rsd = Xd + LP * LWLD;                         // 1 / L_d[k][k]
  double* red = rsd + LP;                               // L_UPD_THREADS
  double* pts_sm = red + L_UPD_THREADS;
  __shared__ int s_next, s_fail;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  double* Lw = g.ws + (size_t)blockIdx.x * (size_t)g.slot_stride;
  const long long ld = g.kmaxp;
  const long long kcap = (g.kmax + 7) / 8 * 8;
  double* pbase = g.pts_in_ws ? Lw + (size_t)(g.kmaxp + 2) * (size_t)g.kmaxp : pts_sm;
  // records: [a | b | c-vector | z | cos(lat)]; gather mode: [joint index (int32) | c-vector | z]
  double *pa = pbase, *pb = pa + kcap, *pcv = (FAST == L_GATHER) ? pa + kcap / 2 : pb + kcap, *pzv = pcv + kcap, *pc = pzv + kcap;
  const double qnan = __longlong_as_double(0x7FF8000000000000LL);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage);
  const bool diag_warp = warp == L_WARPS;
#ifdef CK_LOCAL_PROFILE  // phase counters (tools/k4_run.py --phases on a -DCK_LOCAL_PROFILE build); off in the product build
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;
  const bool prof = g.dbg != nullptr && tid == 0;
#define L_STAMP(i)                      \
  do {                                  \
    if (prof) {                         \
      const long long now_ = clock64(); \
      tph[i] += now_ - tlast;           \
      tlast = now_;                     \
    }                                   \
  } while (0)
  if (prof) tlast = clock64();
#else
#define L_STAMP(i) do {} while (0)
#endif

  for (;;) {
    if (tid == 0) { s_next = (int)atomicAdd(g.counter, 1u); s_fail = 0; }
    __syncthreads();
    const long long c = s_next;
    __syncthreads();
    if (c >= g.m) break;
    const int k = g.kcount[c];
    if (k <= 0) {
      if (tid == 0) { g.pred[c] = qnan; g.sd[c] = qnan; g.info[c] = 0; }
      continue;
    }
    const int kp = (k + LP - 1) / LP * LP, R = kp + 2;

    if (diag_warp) {
      // ===== the diagonal-block warp: per block column wait for D, eliminate, publish X_d.  Its ~30 k cycles of
      // dependent FP64 work per block overlap the update warps' main loops of the remaining row tiles.
      for (int c0 = 0; c0 < kp; c0 += LP) {
        l_bar_sync(L_BAR_D, L_THREADS);
        const int bad = local_diag_block(Xd, Xd, rsd, lane);
        if (lane == 0 && bad != 0 && s_fail == 0) s_fail = c0 + bad;
        l_bar_arrive(L_BAR_X, L_THREADS);
      }
      continue;  // joins the update warps at the __syncthreads of the next target
    }

    // ===== update warps
    // ---- 1. scan + ordered compaction (offsets from the eight segment counts of the first pass)
    int k0 = 0;
    {
      const int* sg = g.kseg + c * (2 * L_WARPS);
      int off0 = 0, off1 = 0;
      for (int w = 0; w < L_WARPS; ++w) {
        const int s0 = sg[w], s1 = sg[L_WARPS + w];
        if (w < warp) { off0 += s0; off1 += s1; }
        k0 += s0;
      }
      const double t0 = g.xyp[2 * c], t1 = g.xyp[2 * c + 1];
      const CkPoint p0 = ck_prepare_point(METRIC, t0, t1);
      const LocalFilter f = local_filter<METRIC>(g, t0, t1);
      local_scan_segment<METRIC, true, FAST == L_GATHER>(g, f, p0, 0, warp, lane, off0, pa, pb, pc, pcv, pzv);
      if (g.n_procs == 2) local_scan_segment<METRIC, true, FAST == L_GATHER>(g, f, p0, 1, warp, lane, k0 + off1, pa, pb, pc, pcv, pzv);
    }
    l_bar_sync(L_BAR_UPD, L_UPD_THREADS);
    L_STAMP(0);

    // ---- 2. left-looking blocked Cholesky of [Sigma_loc ; c ; z]
    for (int c0 = 0; c0 < kp; c0 += LP) {
      const int ntile = (R - c0 + LTM - 1) / LTM;
      const int KT = c0 / LK;
      // 2a. W = Sigma[tile, J] - L[tile, 0:J) L[J, 0:J)^T for every row tile of the block column.  The diagonal block goes
      // to shared memory (for the diag warp), every other row is parked IN PLACE in the factor array; the panel products
      // follow in 2b once X_d is there.
      for (int t = 0; t < ntile; ++t) {
        const int i0 = c0 + t * LTM;
        const int wrow = i0 + LWR * warp;  // first row of this warp's LWR x 32 piece
        // operand loader: A = L[i0 : i0 + 64, 0 : c0), B = L[c0 : c0 + 32, 0 : c0); per stage 96 rows x 8 chunks of 16 B
        auto load_stage = [&](int s, int kt) {
          const unsigned dst0 = sbase + (unsigned)(s * L_STAGE_ELEMS * 8);
#pragma unroll
          for (int it = 0; it < (LTM + LP) * (LK / 2) / L_UPD_THREADS; ++it) {
            const int idx = tid + it * L_UPD_THREADS, row = idx / (LK / 2), ch = idx % (LK / 2);
            const int grow = row < LTM ? i0 + row : c0 + (row - LTM);
            const unsigned on = (kt < KT && grow < R) ? 1u : 0u;
            l_cp_async16(dst0 + (unsigned)((row * LLD + ch * 2) * 8), Lw + (long long)(on ? grow : 0) * ld + kt * LK + ch * 2, on);
          }
        };
        if (KT > 0) {
#pragma unroll
          for (int s = 0; s < LSTG - 1; ++s) {
            load_stage(s, s);
            l_cp_commit();
          }
        }
        // accumulators start at -Sigma (so that W = -(acc + A B^T ...) needs no negated operand): entries gathered /
        // re-computed while the first operand stage is in flight
        double acc[LMI][4][2];
        const bool plain = (t > 0) && (i0 + LTM <= k);  // every row is a matrix row below the diagonal block
        // rows of the identity pad (k <= i < kp) stay e_i through the whole sweep and rows past R do not exist: a warp whose
        // 16 rows are all of that kind issues no DMMA (it still takes part in the operand staging and the barriers)
#ifdef CK_LOCAL_NO_SKIP
        const bool live = true;
#else
        const bool live = wrow < R && !(wrow >= k && wrow + LWR <= kp);
#endif
#pragma unroll
        for (int mi = 0; mi < LMI; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = wrow + 8 * mi + g4, j = c0 + 8 * ni + 2 * t4 + e;
              double v;
              if (plain) v = local_cov_entry<METRIC, FAST>(g, i, j, k0, pa, pb, pc);
              else v = (i < R) ? local_entry<METRIC, FAST>(g, i, j, k, kp, k0, pa, pb, pc, pcv, pzv) : 0.0;
              acc[mi][ni][e] = -v;
            }
        L_STAMP(1);
        if (KT > 0) {
          int rd = 0, wr = LSTG - 1;
          for (int kt = 0; kt < KT; ++kt) {
            l_cp_wait<LSTG - 2>();
            l_bar_sync(L_BAR_UPD, L_UPD_THREADS);
            load_stage(wr, kt + LSTG - 1);
            l_cp_commit();
            const double* As = stage + rd * L_STAGE_ELEMS + (LWR * warp + g4) * LLD + t4;
            const double* Bs = stage + rd * L_STAGE_ELEMS + (LTM + g4) * LLD + t4;
            if (live)
#pragma unroll
            for (int kk = 0; kk < LK; kk += 4) {
              double a[LMI], b[4];
#pragma unroll
              for (int mi = 0; mi < LMI; ++mi) a[mi] = As[mi * 8 * LLD + kk];
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[ni * 8 * LLD + kk];
#pragma unroll
              for (int mi = 0; mi < LMI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) l_dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            rd = (rd + 1 == LSTG) ? 0 : rd + 1;
            wr = (wr + 1 == LSTG) ? 0 : wr + 1;
          }
          l_cp_wait<0>();
          l_bar_sync(L_BAR_UPD, L_UPD_THREADS);  // every warp is done with the staging buffers: the next tile may refill them
        }
        L_STAMP(2);
        // W: diagonal block -> shared memory (read by the diag warp), everything else -> parked in place
        if (t == 0 && warp < LP / LWR) {
#pragma unroll
          for (int mi = 0; mi < LMI; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
              *reinterpret_cast<double2*>(Xd + (LWR * warp + 8 * mi + g4) * LWLD + 8 * ni + 2 * t4) =
                  make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
        } else {
#pragma unroll
          for (int mi = 0; mi < LMI; ++mi) {
            const int i = wrow + 8 * mi + g4;
            if (i < R) {
              double* dst = Lw + (long long)i * ld + c0 + 2 * t4;
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) *reinterpret_cast<double2*>(dst + 8 * ni) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
            }
          }
        }
        if (t == 0) l_bar_arrive(L_BAR_D, L_THREADS);  // D is complete once all four update warps have arrived
      }
      // 2b. panel products P = W X_d^T: every warp handles the rows it parked itself (same lanes' warp, so __syncwarp
      // orders the global round trip); column tile ni only needs k <= 8 ni + 7 because X_d is lower triangular
      l_bar_sync(L_BAR_X, L_THREADS);
      L_STAMP(3);
      __syncwarp();
      for (int ts = 0; ts < ntile * (LWR / 16); ++ts) {
        const int t = ts / (LWR / 16);
        if (t == 0 && warp < LP / LWR) continue;
        const int wrow = c0 + t * LTM + LWR * warp + 16 * (ts % (LWR / 16));  // 16-row strips of the rows this warp parked
        if (wrow >= R) continue;
        double out[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            out[mi][ni][0] = 0.0;
            out[mi][ni][1] = 0.0;
          }
        const int r0 = wrow + g4, r1 = wrow + 8 + g4;
        const double* w0 = Lw + (long long)(r0 < R ? r0 : wrow) * ld + c0 + t4;
        const double* w1 = Lw + (long long)(r1 < R ? r1 : wrow) * ld + c0 + t4;
        double a0[LP / 4], a1[LP / 4];
#pragma unroll
        for (int ks = 0; ks < LP / 4; ++ks) {
          a0[ks] = __ldcg(w0 + 4 * ks);
          a1[ks] = __ldcg(w1 + 4 * ks);
        }
        __syncwarp();  // all fragments of the strip are in registers before any row of it is overwritten
#pragma unroll
        for (int ks = 0; ks < LP / 4; ++ks) {
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            if (4 * ks > 8 * ni + 7) continue;  // X_d is lower triangular
            const double b = Xd[(8 * ni + g4) * LWLD + 4 * ks + t4];
            l_dmma(out[0][ni][0], out[0][ni][1], a0[ks], b);
            l_dmma(out[1][ni][0], out[1][ni][1], a1[ks], b);
          }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const int i = wrow + 8 * mi + g4;
          if (i < R) {
            double* dst = Lw + (long long)i * ld + c0 + 2 * t4;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) *reinterpret_cast<double2*>(dst + 8 * ni) = make_double2(out[mi][ni][0], out[mi][ni][1]);
          }
        }
      }
      l_bar_sync(L_BAR_UPD, L_UPD_THREADS);  // the block column of L is complete: the next one may stage it
      L_STAMP(4);
    }
    // ---- 3. prediction: rows kp (v = L^-1 c) and kp + 1 (y = L^-1 z) of the factor array
    double sv = 0.0, sy = 0.0;
    {
      const double* vrow = Lw + (long long)kp * ld;
      const double* yrow = vrow + ld;
      for (int b = tid; b < kp; b += L_UPD_THREADS) {
        const double v = __ldcg(vrow + b), y = __ldcg(yrow + b);
        sv = fma(v, v, sv);
        sy = fma(v, y, sy);
      }
    }
    red[tid] = sv;
    l_bar_sync(L_BAR_UPD, L_UPD_THREADS);
    for (int h = L_UPD_THREADS / 2; h > 0; h >>= 1) { if (tid < h) red[tid] += red[tid + h]; l_bar_sync(L_BAR_UPD, L_UPD_THREADS); }
    sv = red[0];
    l_bar_sync(L_BAR_UPD, L_UPD_THREADS);
    red[tid] = sy;
    l_bar_sync(L_BAR_UPD, L_UPD_THREADS);
    for (int h = L_UPD_THREADS / 2; h > 0; h >>= 1) { if (tid < h) red[tid] += red[tid + h]; l_bar_sync(L_BAR_UPD, L_UPD_THREADS); }
    sy = red[0];
    if (tid == 0) {
      if (s_fail) {
        g.pred[c] = qnan; g.sd[c] = qnan; g.info[c] = s_fail;
      } else {
        const double var = g.c0 - sv;
        double sd = sqrt(var);          // NaN if var < 0
        if (!(sd > 0.0)) sd = 0.0;      // np.nanmax([sd, 0.0])
        g.pred[c] = sy; g.sd[c] = sd; g.info[c] = (var > 0.0) ? 0 : -1;  // -1: augmented matrix not PD (warning only)
      }
    }
    L_STAMP(5);
  }
#ifdef CK_LOCAL_PROFILE
  if (prof)
    for (int i = 0; i < 6; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(g.dbg) + i, (unsigned long long)tph[i]);
#endif
#undef L_STAMP
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int local_fill(LocalArgs* g, const double* xy0, const double* z0, ck_i64 n0, const double* xy1, const double* z1,
                      ck_i64 n1, const double* xyp, ck_i64 m, const double* params_sigma, const double* params_pred,
                      int n_procs, int i_pred, int metric, double max_dist, int cv) {
  CkParams ps, pp;
  int rc = ck_unpack_params(params_sigma, n_procs, &ps);
  if (rc) return rc;
  if ((rc = ck_unpack_params(params_pred ? params_pred : params_sigma, n_procs, &pp))) return rc;
  CK_REQUIRE(i_pred >= 0 && i_pred < n_procs, "i_pred out of range");
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  CK_REQUIRE(n0 >= 0 && n1 >= 0 && m >= 0, "negative size");
  memset(g, 0, sizeof(*g));
  g->xy[0] = xy0; g->z[0] = z0; g->n[0] = n0;
  g->xy[1] = xy1; g->z[1] = z1; g->n[1] = (n_procs == 2) ? n1 : 0;
  g->xyp = xyp; g->m = m;
  if ((rc = ck_block_matern(ps, 0, 0, 1, &g->S[0]))) return rc;
  if (n_procs == 2) {
    if ((rc = ck_block_matern(ps, 0, 1, 1, &g->S[1]))) return rc;
    if ((rc = ck_block_matern(ps, 1, 1, 1, &g->S[2]))) return rc;
  } else {
    g->S[1] = g->S[0];
    g->S[2] = g->S[0];
  }
  for (int j = 0; j < n_procs; ++j)
    if ((rc = ck_block_matern(pp, i_pred, j, 1, &g->C[j]))) return rc;
  g->n_procs = n_procs; g->i_pred = i_pred; g->metric = metric; g->cv = cv;
  g->max_dist = max_dist;
  return CK_OK;
}

struct LocalPlan {
  long long kmax, kmaxp, slots, slot_stride;  // slot_stride in doubles
  int pts_in_ws;
  size_t smem;
};

static LocalPlan local_plan(ck_i64 m, ck_i64 kmax, int metric, bool gather = false) {
  LocalPlan p;
  p.kmax = kmax > 0 ? kmax : 1;
  p.kmaxp = (p.kmax + LP - 1) / LP * LP;
  p.pts_in_ws = p.kmax > L_SMEM_KMAX ? 1 : 0;
  const long long kcap = (p.kmax + 7) / 8 * 8;
  const long long rec = 5 * kcap;  // a, b, c-vector, z, cos(lat)
  p.slot_stride = (p.kmaxp + 2) * p.kmaxp + (p.pts_in_ws ? rec : 0);
  p.slot_stride = (p.slot_stride + 31) / 32 * 32;
  long long slots = 148 * L_MIN_CTAS;  // one slot per resident CTA
  const long long budget = (8LL << 30) / 8;  // at most 8 GB of factor slots
  if (slots * p.slot_stride > budget) slots = budget / p.slot_stride;
  if (slots < 148) slots = 148;
  if (slots > m) slots = m > 0 ? m : 1;
  p.slots = slots;
  const long long narr = (metric == CK_METRIC_HAVERSINE) ? 5 : 4;
  const long long recs = gather ? kcap / 2 + 2 * kcap : narr * kcap;
  p.smem = (size_t)(LSTG * L_STAGE_ELEMS + LP * LWLD + LP + L_UPD_THREADS + (p.pts_in_ws ? 0 : recs)) * sizeof(double);
  return p;
}

// profiling aid (tools): six cycle counters summed over the CTAs of the following ck_local_predict launches
static long long* g_local_dbg = nullptr;
extern "C" int ck_local_debug_buffer(void* dev_counters) {
  g_local_dbg = static_cast<long long*>(dev_counters);
#ifdef CK_LOCAL_PROFILE
  return CK_OK;
#else
  return dev_counters ? CK_ERR_UNSUPPORTED : CK_OK;  // the product build carries no phase counters
#endif
}

extern "C" size_t ck_local_predict_workspace_bytes(ck_i64 m, ck_i64 kmax) {
  if (m <= 0 || kmax <= 0) return 256;
  const LocalPlan p = local_plan(m, kmax, CK_METRIC_HAVERSINE);
  return (size_t)p.slots * (size_t)p.slot_stride * sizeof(double) + 256;
}

extern "C" int ck_local_count(const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* xyp, ck_i64 m,
                              int n_procs, int i_pred, int metric, double max_dist, int cv, int* k_dev, int* seg_dev,
                              void* stream) {
  static const double dummy1[4] = {1, 1.5, 1, 0}, dummy2[11] = {1, 1, 1.5, 1.5, 1.5, 1, 1, 1, 0, 0, 0};
  LocalArgs g;
  int rc = local_fill(&g, xy0, nullptr, n0, xy1, nullptr, n1, xyp, m, n_procs == 1 ? dummy1 : dummy2, nullptr, n_procs, i_pred,
                      metric, max_dist, cv);
  if (rc) return rc;
  if (m == 0) return CK_OK;
  CK_REQUIRE(k_dev && seg_dev && xyp, "null pointer");
  CK_REQUIRE((((uintptr_t)xy0 | (uintptr_t)xy1) & 15) == 0, "coordinate arrays must be 16-byte aligned");
  const unsigned grid = (unsigned)(m < 148 * 16 ? m : 148 * 16);
  cudaStream_t st = ck_stream(stream);
  if (metric == CK_METRIC_HAVERSINE) ck_local_count_kernel<CK_METRIC_HAVERSINE><<<grid, L_UPD_THREADS, 0, st>>>(g, k_dev, seg_dev);
  else ck_local_count_kernel<CK_METRIC_EUCLID><<<grid, L_UPD_THREADS, 0, st>>>(g, k_dev, seg_dev);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

template <int METRIC, int FAST>
static int local_launch(const LocalArgs& g, const LocalPlan& p, cudaStream_t st) {
  static CkPerDevice attr;
  CK_SET_SMEM_ONCE(attr, (ck_local_predict_kernel<METRIC, FAST>), 200 * 1024);
  ck_local_predict_kernel<METRIC, FAST><<<(unsigned)p.slots, L_THREADS, p.smem, st>>>(g);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

template <int METRIC>
static int local_dispatch(const LocalArgs& g, const LocalPlan& p, cudaStream_t st) {
  if (g.sigma) return local_launch<METRIC, L_GATHER>(g, p, st);
  const int mode = g.S[0].mode;
  const bool uniform = mode != CK_NU_GENERIC && g.S[1].mode == mode && g.S[2].mode == mode;
  if (uniform) {
    switch (mode) {
      case CK_NU_HALF: return local_launch<METRIC, CK_NU_HALF>(g, p, st);
      case CK_NU_3HALF: return local_launch<METRIC, CK_NU_3HALF>(g, p, st);
      case CK_NU_5HALF: return local_launch<METRIC, CK_NU_5HALF>(g, p, st);
      default: return local_launch<METRIC, CK_NU_7HALF>(g, p, st);
    }
  }
  return local_launch<METRIC, L_FAST_NONE>(g, p, st);
}

extern "C" int ck_local_predict(const double* xy0, const double* z0, ck_i64 n0, const double* xy1, const double* z1,
                                ck_i64 n1, const double* xyp, ck_i64 m, const double* params_sigma, const double* params_pred,
                                int n_procs, int i_pred, int metric, double max_dist, int cv, double c0, const double* sigma,
                                ck_i64 ld_sigma, const int* k_dev, const int* seg_dev, ck_i64 kmax, double* pred, double* sd,
                                int* info, void* ws, void* stream) {
  LocalArgs g;
  int rc = local_fill(&g, xy0, z0, n0, xy1, z1, n1, xyp, m, params_sigma, params_pred, n_procs, i_pred, metric, max_dist, cv);
  if (rc) return rc;
  if (m == 0) return CK_OK;
  CK_REQUIRE(k_dev && seg_dev && pred && sd && info && xyp, "null pointer");
  CK_REQUIRE(kmax >= 0, "negative kmax");
  CK_REQUIRE(ws, "workspace is NULL");
  CK_REQUIRE((((uintptr_t)xy0 | (uintptr_t)xy1 | (uintptr_t)ws) & 15) == 0, "coordinate arrays / workspace must be 16-byte aligned");
  const ck_i64 ntot = n0 + ((n_procs == 2) ? n1 : 0);
  CK_REQUIRE(!sigma || (ld_sigma >= ntot && ntot < (1LL << 31)), "bad leading dimension of the joint covariance");
  const LocalPlan p = local_plan(m, kmax, metric, sigma != nullptr);
  CK_REQUIRE(p.smem <= 200 * 1024, "kmax too large for the shared-memory plan");
  g.c0 = c0;
  g.sigma = sigma; g.ld_sigma = ld_sigma;
  g.kcount = k_dev; g.kseg = seg_dev; g.kmax = p.kmax; g.kmaxp = p.kmaxp;
  g.pred = pred; g.sd = sd; g.info = info;
  // the last 256 bytes of the workspace hold the target counter of the persistent CTAs
  char* wsb = static_cast<char*>(ws);
  g.counter = reinterpret_cast<unsigned int*>(wsb + (size_t)p.slots * (size_t)p.slot_stride * sizeof(double));
  g.ws = static_cast<double*>(ws);
  g.slot_stride = p.slot_stride;
  g.pts_in_ws = p.pts_in_ws;
  g.dbg = g_local_dbg;
  cudaStream_t st = ck_stream(stream);
  CK_CUDA(cudaMemsetAsync(g.counter, 0, sizeof(unsigned int), st));
  if (metric == CK_METRIC_HAVERSINE) return local_dispatch<CK_METRIC_HAVERSINE>(g, p, st);
  return local_dispatch<CK_METRIC_EUCLID>(g, p, st);
}

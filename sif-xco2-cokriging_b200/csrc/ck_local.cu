// ck_local.cu -- K4: batched local-neighbourhood (point) cokriging for sm_100a.
//
// One CTA per prediction target (persistent over targets, one workspace slot per CTA):
//   1. ordered compaction of the data within max_dist (process 0 then 1, index order -- the order of
//      the reference's boolean masks, src/point_prediction.py:127-151);
//   2. the local covariance matrix is RE-COMPUTED from coordinates (cheaper than gathering k^2
//      entries of a stored N x N Sigma and needs no N^2 memory), lower triangle only, with two
//      extra rows appended: c (target-to-neighbour covariances) and z (neighbour data);
//   3. right-looking blocked Cholesky (panel 32) applied to the (k+2) x k array: the panel step
//      that turns rows of S into rows of L turns the two extra rows into v = L^{-1} c, y = L^{-1} z,
//      so no separate triangular solve is needed;
//   4. pred = v . y,  sd = sqrt(max(c0 - v . v, 0))   (src/point_prediction.py:200-222).
#include "ck_common.cuh"

constexpr int LW = 32;         // panel width
constexpr int LT = 64;         // trailing-update tile
constexpr int L_THREADS = 256;

struct LocalArgs {
  const double* xy[2]; const double* z[2]; long long n[2];
  const double* xyp; long long m;
  CkMatern S[2][2];  // joint-model blocks (S[1][0] == S[0][1])
  CkMatern C[2];     // target-to-process-j blocks for the predicted process (own block carries the nugget)
  int n_procs, i_pred, metric, cv;
  double max_dist, c0;
  const int* kcount; long long kmax, ldk;
  double* pred; double* sd; int* info;
  double* ws_mat;   // slots x (kmax+2) x ldk
  double* ws_pts;   // slots x kmax x 4   (CkPoint a, b, c + process id)
};

// out-of-line so that the five Matern variants are instantiated once per kernel, not per call site
static __device__ __noinline__ double cov_dyn(const CkMatern& P, double h) { return ck_matern_cov_dyn(P, h); }

template <int METRIC>
__device__ __forceinline__ bool local_keep(const LocalArgs& g, int proc, double d) {
  if (!(d <= g.max_dist)) return false;
  if (g.cv && proc == g.i_pred && !(d > 0.0)) return false;
  return true;
}

template <int METRIC>
__global__ void __launch_bounds__(L_THREADS) ck_local_count_kernel(LocalArgs g, int* __restrict__ kout) {
  __shared__ int red[L_THREADS];
  for (long long c = blockIdx.x; c < g.m; c += gridDim.x) {
    const CkPoint p0 = ck_prepare_point(METRIC, g.xyp[2 * c], g.xyp[2 * c + 1]);
    int cnt = 0;
    for (int proc = 0; proc < g.n_procs; ++proc)
      for (long long i = threadIdx.x; i < g.n[proc]; i += L_THREADS) {
        const CkPoint q = ck_prepare_point(METRIC, g.xy[proc][2 * i], g.xy[proc][2 * i + 1]);
        cnt += local_keep<METRIC>(g, proc, ck_dist<METRIC>(p0, q)) ? 1 : 0;
      }
    red[threadIdx.x] = cnt;
    __syncthreads();
    for (int h = L_THREADS / 2; h > 0; h >>= 1) {
      if ((int)threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
      __syncthreads();
    }
    if (threadIdx.x == 0) kout[c] = red[0];
    __syncthreads();
  }
}

template <int METRIC>
__global__ void __launch_bounds__(L_THREADS) ck_local_predict_kernel(LocalArgs g) {
  __shared__ double Ds[LW][LW + 1];
  __shared__ double Pa[LT][LW + 1], Pb[LT][LW + 1];
  __shared__ int wcount[L_THREADS / 32];
  __shared__ int s_total, s_fail;
  __shared__ double red[L_THREADS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* M = g.ws_mat + (size_t)blockIdx.x * (size_t)(g.kmax + 2) * (size_t)g.ldk;
  double* pts = g.ws_pts + (size_t)blockIdx.x * (size_t)g.kmax * 4;
  const long long ld = g.ldk;
  const double qnan = __longlong_as_double(0x7FF8000000000000LL);

  for (long long c = blockIdx.x; c < g.m; c += gridDim.x) {
    const int k = g.kcount[c];
    if (k <= 0) {
      if (tid == 0) { g.pred[c] = qnan; g.sd[c] = qnan; g.info[c] = 0; }
      continue;
    }
    const CkPoint p0 = ck_prepare_point(METRIC, g.xyp[2 * c], g.xyp[2 * c + 1]);
    // ---- 1. ordered compaction; also fills row k (c vector) and row k+1 (z) ----
    if (tid == 0) { s_total = 0; s_fail = 0; }
    __syncthreads();
    for (int proc = 0; proc < g.n_procs; ++proc) {
      for (long long base = 0; base < g.n[proc]; base += L_THREADS) {
        const long long i = base + tid;
        bool keep = false;
        CkPoint q;
        double d = 0.0;
        if (i < g.n[proc]) {
          q = ck_prepare_point(METRIC, g.xy[proc][2 * i], g.xy[proc][2 * i + 1]);
          d = ck_dist<METRIC>(p0, q);
          keep = local_keep<METRIC>(g, proc, d);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcount[warp] = __popc(mask);
        __syncthreads();
        int off = s_total;
        for (int w = 0; w < warp; ++w) off += wcount[w];
        if (keep) {
          const int pos = off + __popc(mask & ((1u << lane) - 1u));
          if (pos < k) {
            pts[4 * pos + 0] = q.a; pts[4 * pos + 1] = q.b; pts[4 * pos + 2] = q.c; pts[4 * pos + 3] = (double)proc;
            M[(long long)k * ld + pos] = cov_dyn(g.C[proc], d);
            M[(long long)(k + 1) * ld + pos] = g.z[proc][i];
          }
        }
        __syncthreads();
        if (tid == 0) {
          int tot = 0;
          for (int w = 0; w < L_THREADS / 32; ++w) tot += wcount[w];
          s_total += tot;
        }
        __syncthreads();
      }
    }
    // ---- 2. lower triangle of the local covariance ----
    for (int a = warp; a < k; a += L_THREADS / 32) {
      CkPoint pa; pa.a = pts[4 * a]; pa.b = pts[4 * a + 1]; pa.c = pts[4 * a + 2];
      const int proca = (int)pts[4 * a + 3];
      for (int b = lane; b <= a; b += 32) {
        CkPoint pb; pb.a = pts[4 * b]; pb.b = pts[4 * b + 1]; pb.c = pts[4 * b + 2];
        const int procb = (int)pts[4 * b + 3];
        // distance argument order follows the reference block (i <= j): rows from the lower process id
        const double d = (procb <= proca) ? ck_dist<METRIC>(pb, pa) : ck_dist<METRIC>(pa, pb);
        M[(long long)a * ld + b] = cov_dyn(g.S[procb][proca], d);
      }
    }
    __syncthreads();
    // ---- 3. blocked Cholesky on the (k+2) x k array ----
    const int R = k + 2;
    for (int j0 = 0; j0 < k; j0 += LW) {
      const int w = (k - j0 < LW) ? k - j0 : LW;
      for (int e = tid; e < LW * LW; e += L_THREADS) {
        const int i = e / LW, cc = e % LW;
        double v = (i == cc) ? 1.0 : 0.0;
        if (i < w && cc < w && cc <= i) v = M[(long long)(j0 + i) * ld + j0 + cc];
        else if (i < w && cc < w) v = 0.0;
        Ds[i][cc] = v;
      }
      __syncthreads();
      if (warp == 0) {  // unblocked factorisation of the diagonal block, lane = row
        for (int jj = 0; jj < LW; ++jj) {
          const double dj = Ds[jj][jj];
          if (!(dj > 0.0) && lane == 0 && s_fail == 0) s_fail = j0 + jj + 1;
          const double sq = sqrt(dj);
          __syncwarp();
          if (lane == jj) Ds[jj][jj] = sq;
          if (lane > jj) Ds[lane][jj] = Ds[lane][jj] / sq;
          __syncwarp();
          if (lane > jj) {
            const double l = Ds[lane][jj];
            for (int cc = jj + 1; cc <= lane; ++cc) Ds[lane][cc] -= l * Ds[cc][jj];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      for (int e = tid; e < w * w; e += L_THREADS) {
        const int i = e / w, cc = e % w;
        if (cc <= i) M[(long long)(j0 + i) * ld + j0 + cc] = Ds[i][cc];
      }
      // panel: rows below the diagonal block (incl. the two extra rows): row <- row L_d^{-T}
      for (int i = j0 + w + tid; i < R; i += L_THREADS) {
        double r[LW];
        double* row = M + (long long)i * ld + j0;
#pragma unroll
        for (int cc = 0; cc < LW; ++cc) r[cc] = (cc < w) ? row[cc] : 0.0;
#pragma unroll
        for (int cc = 0; cc < LW; ++cc) {
          double s = r[cc];
#pragma unroll
          for (int t = 0; t < cc; ++t) s -= r[t] * Ds[cc][t];
          r[cc] = s / Ds[cc][cc];
        }
#pragma unroll
        for (int cc = 0; cc < LW; ++cc)
          if (cc < w) row[cc] = r[cc];
      }
      __syncthreads();
      // trailing update: M[i][cc] -= sum_t P[i][t] P[cc][t], i >= j0+w, j0+w <= cc < k, cc <= i for matrix rows
      const int t0 = j0 + w;
      if (t0 < k) {
        const int tx = tid & 15, ty = tid >> 4;
        for (int ti = t0; ti < R; ti += LT) {
          for (int tc = t0; tc < k && tc <= ti + LT - 1; tc += LT) {
            for (int e = tid; e < LT * LW; e += L_THREADS) {
              const int rr = e / LW, cc = e % LW;
              Pa[rr][cc] = (ti + rr < R && cc < w) ? M[(long long)(ti + rr) * ld + j0 + cc] : 0.0;
              Pb[rr][cc] = (tc + rr < k && cc < w) ? M[(long long)(tc + rr) * ld + j0 + cc] : 0.0;
            }
            __syncthreads();
            double acc[4][4];
#pragma unroll
            for (int r_ = 0; r_ < 4; ++r_)
#pragma unroll
              for (int q_ = 0; q_ < 4; ++q_) acc[r_][q_] = 0.0;
#pragma unroll 8
            for (int t = 0; t < LW; ++t) {
              double av[4], bv[4];
#pragma unroll
              for (int r_ = 0; r_ < 4; ++r_) av[r_] = Pa[ty + 16 * r_][t];
#pragma unroll
              for (int q_ = 0; q_ < 4; ++q_) bv[q_] = Pb[tx + 16 * q_][t];
#pragma unroll
              for (int r_ = 0; r_ < 4; ++r_)
#pragma unroll
                for (int q_ = 0; q_ < 4; ++q_) acc[r_][q_] += av[r_] * bv[q_];
            }
#pragma unroll
            for (int r_ = 0; r_ < 4; ++r_)
#pragma unroll
              for (int q_ = 0; q_ < 4; ++q_) {
                const int i = ti + ty + 16 * r_, cc = tc + tx + 16 * q_;
                if (i < R && cc < k && (cc <= i)) M[(long long)i * ld + cc] -= acc[r_][q_];
              }
            __syncthreads();
          }
        }
      }
    }
    // ---- 4. prediction ----
    double sv = 0.0, sy = 0.0;
    for (int b = tid; b < k; b += L_THREADS) {
      const double v = M[(long long)k * ld + b], y = M[(long long)(k + 1) * ld + b];
      sv += v * v;
      sy += v * y;
    }
    red[tid] = sv;
    __syncthreads();
    for (int h = L_THREADS / 2; h > 0; h >>= 1) { if (tid < h) red[tid] += red[tid + h]; __syncthreads(); }
    sv = red[0];
    __syncthreads();
    red[tid] = sy;
    __syncthreads();
    for (int h = L_THREADS / 2; h > 0; h >>= 1) { if (tid < h) red[tid] += red[tid + h]; __syncthreads(); }
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
    __syncthreads();
  }
}

static int local_fill(LocalArgs* g, const double* xy0, const double* z0, ck_i64 n0, const double* xy1, const double* z1,
                      ck_i64 n1, const double* xyp, ck_i64 m, const double* params, int n_procs, int i_pred, int metric,
                      double max_dist, int cv) {
  CkParams p;
  int rc = ck_unpack_params(params, n_procs, &p);
  if (rc) return rc;
  CK_REQUIRE(i_pred >= 0 && i_pred < n_procs, "i_pred out of range");
  CK_REQUIRE(metric == CK_METRIC_EUCLID || metric == CK_METRIC_HAVERSINE, "bad metric %d", metric);
  CK_REQUIRE(n0 >= 0 && n1 >= 0 && m >= 0, "negative size");
  memset(g, 0, sizeof(*g));
  g->xy[0] = xy0; g->z[0] = z0; g->n[0] = n0;
  g->xy[1] = xy1; g->z[1] = z1; g->n[1] = (n_procs == 2) ? n1 : 0;
  g->xyp = xyp; g->m = m;
  for (int i = 0; i < n_procs; ++i)
    for (int j = 0; j < n_procs; ++j)
      if ((rc = ck_block_matern(p, i, j, 1, &g->S[i][j]))) return rc;
  for (int j = 0; j < n_procs; ++j)
    if ((rc = ck_block_matern(p, i_pred, j, 1, &g->C[j]))) return rc;
  g->n_procs = n_procs; g->i_pred = i_pred; g->metric = metric; g->cv = cv;
  g->max_dist = max_dist;
  g->c0 = p.sigma[i_pred] * p.sigma[i_pred] + p.nugget[i_pred];  // covariance(i, 0, use_nugget=True), src/point_prediction.py:66
  return CK_OK;
}

static inline long long local_slots(ck_i64 m) { return m < 148 * 2 ? (m > 0 ? m : 1) : 148 * 2; }
static inline long long local_ldk(ck_i64 kmax) { return (kmax + 15) / 16 * 16; }

extern "C" size_t ck_local_predict_workspace_bytes(ck_i64 m, ck_i64 kmax) {
  if (m <= 0 || kmax <= 0) return 256;
  const size_t slots = (size_t)local_slots(m);
  return slots * ((size_t)(kmax + 2) * (size_t)local_ldk(kmax) + (size_t)kmax * 4) * sizeof(double) + 256;
}

extern "C" int ck_local_count(const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* xyp, ck_i64 m,
                              int n_procs, int i_pred, int metric, double max_dist, int cv, int* k_dev, void* stream) {
  static const double dummy1[4] = {1, 1.5, 1, 0}, dummy2[11] = {1, 1, 1.5, 1.5, 1.5, 1, 1, 1, 0, 0, 0};
  LocalArgs g;
  int rc = local_fill(&g, xy0, nullptr, n0, xy1, nullptr, n1, xyp, m, n_procs == 1 ? dummy1 : dummy2, n_procs, i_pred,
                      metric, max_dist, cv);
  if (rc) return rc;
  if (m == 0) return CK_OK;
  CK_REQUIRE(k_dev && xyp, "null pointer");
  const unsigned grid = (unsigned)(m < 148 * 8 ? m : 148 * 8);
  cudaStream_t st = ck_stream(stream);
  if (metric == CK_METRIC_HAVERSINE) ck_local_count_kernel<CK_METRIC_HAVERSINE><<<grid, L_THREADS, 0, st>>>(g, k_dev);
  else ck_local_count_kernel<CK_METRIC_EUCLID><<<grid, L_THREADS, 0, st>>>(g, k_dev);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_local_predict(const double* xy0, const double* z0, ck_i64 n0, const double* xy1, const double* z1,
                                ck_i64 n1, const double* xyp, ck_i64 m, const double* params, int n_procs, int i_pred,
                                int metric, double max_dist, int cv, const int* k_dev, ck_i64 kmax, double* pred,
                                double* sd, int* info, void* ws, void* stream) {
  LocalArgs g;
  int rc = local_fill(&g, xy0, z0, n0, xy1, z1, n1, xyp, m, params, n_procs, i_pred, metric, max_dist, cv);
  if (rc) return rc;
  if (m == 0) return CK_OK;
  CK_REQUIRE(k_dev && pred && sd && info && xyp, "null pointer");
  CK_REQUIRE(kmax >= 0, "negative kmax");
  CK_REQUIRE(kmax == 0 || ws, "workspace is NULL");
  const long long slots = local_slots(m);
  g.kcount = k_dev; g.kmax = kmax > 0 ? kmax : 1; g.ldk = local_ldk(g.kmax);
  g.pred = pred; g.sd = sd; g.info = info;
  g.ws_mat = static_cast<double*>(ws);
  g.ws_pts = g.ws_mat + (size_t)slots * (size_t)(g.kmax + 2) * (size_t)g.ldk;
  cudaStream_t st = ck_stream(stream);
  if (metric == CK_METRIC_HAVERSINE) ck_local_predict_kernel<CK_METRIC_HAVERSINE><<<(unsigned)slots, L_THREADS, 0, st>>>(g);
  else ck_local_predict_kernel<CK_METRIC_EUCLID><<<(unsigned)slots, L_THREADS, 0, st>>>(g);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

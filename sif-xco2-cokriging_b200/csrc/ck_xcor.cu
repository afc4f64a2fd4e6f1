// ck_xcor.cu -- batched temporal cross-correlation for sm_100a.
//
// The reference computes, per (lon, lat) cell, the masked Pearson cross-correlation of two time series at an integer lag
// (src/stat_tools.py:128-160 compute_xcor_nd), after a per-cell linear detrend (apply_detrend, :56-75), once per lag,
// and takes the lag with the largest |xcor| (optim_lag_nd, :181-233): nlag passes of masked-array numpy over the cube.
// Here ONE launch streams both cubes once: one warp per cell keeps the two series in shared memory, optionally removes
// the least-squares index trend, subtracts the masked means and evaluates every lag, the arg-max included.
// HBM-bound by construction (2 x 8 T bytes read per cell, 8 nlag written); all sums use fixed xor-shuffle trees.
//
// Semantics follow numpy.ma exactly: X = Z1 - mean(valid Z1), Y likewise; for lag != 0 the windows are
// Python slices X[lag:], Y[:-lag] (a NEGATIVE lag pairs the last |lag| entries of X with the first |lag| of Y);
// sum(X Y) runs over jointly valid positions, sum(X X) / sum(Y Y) over each series' own valid positions inside its
// window; fewer than tau joint positions (tau > 0), an empty joint support or a zero denominator give NaN.
#include "ck_common.cuh"

constexpr int X_THREADS = 256;
constexpr int X_WARPS = X_THREADS / 32;
constexpr int X_MAX_LAGS = 64;

struct XcorArgs {
  const double* z1; const double* z2; long long ncell; int T;
  int lags[X_MAX_LAGS]; int nlag; int tau; int detrend;
  double* xcor;      // [nlag][ncell]
  int* best_idx;     // [ncell] or NULL: index (into lags) of the largest |xcor|, 0 if every lag is NaN
  double* best_xcor; // [ncell] or NULL
};

__device__ __forceinline__ double x_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int x_warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// in place: s <- s - mean(valid s), after an optional removal of the least-squares line through (index, value)
__device__ __forceinline__ void x_center(double* s, int T, int lane, bool detrend) {
  double sy = 0.0, st = 0.0;
  int n = 0;
  for (int t = lane; t < T; t += 32) {
    const double v = s[t];
    if (v == v) { sy += v; st += (double)t; ++n; }
  }
  sy = x_warp_sum(sy); st = x_warp_sum(st); n = x_warp_sum(n);
  if (n == 0) return;  // all missing: stays NaN
  double ybar = sy / n;
  if (detrend) {
    const double tbar = st / n;
    double stt = 0.0, sty = 0.0;
    for (int t = lane; t < T; t += 32) {
      const double v = s[t];
      if (v == v) { const double tc = (double)t - tbar; stt += tc * tc; sty += tc * (v - ybar); }
    }
    stt = x_warp_sum(stt); sty = x_warp_sum(sty);
    const double slope = stt > 0.0 ? sty / stt : 0.0;
    double sr = 0.0;
    for (int t = lane; t < T; t += 32) {
      const double v = s[t];
      if (v == v) { const double r = v - (ybar + slope * ((double)t - tbar)); s[t] = r; sr += r; }
    }
    __syncwarp();
    ybar = x_warp_sum(sr) / n;  // the residuals' own mean (~1e-17): compute_xcor_nd subtracts it again
  }
  for (int t = lane; t < T; t += 32) s[t] -= ybar;  // NaN stays NaN
  __syncwarp();
}

__global__ void __launch_bounds__(X_THREADS) ck_xcor_kernel(const __grid_constant__ XcorArgs g) {
  extern __shared__ double xsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* x = xsm + (size_t)warp * 2 * g.T;
  double* y = x + g.T;
  const double qnan = __longlong_as_double(0x7FF8000000000000LL);
  for (long long c = (long long)blockIdx.x * X_WARPS + warp; c < g.ncell; c += (long long)gridDim.x * X_WARPS) {
    for (int t = lane; t < g.T; t += 32) {
      x[t] = g.z1[c * g.T + t];
      y[t] = g.z2[c * g.T + t];
    }
    __syncwarp();
    x_center(x, g.T, lane, g.detrend != 0);
    x_center(y, g.T, lane, g.detrend != 0);
    int best = 0;
    double best_abs = -1.0, best_val = qnan;
    for (int l = 0; l < g.nlag; ++l) {
      const int lag = g.lags[l];
      // Python slices X[lag:], Y[:-lag] on T entries (lag == 0: everything)
      int sx = 0, len = g.T;
      if (lag > 0) { sx = lag < g.T ? lag : g.T; len = g.T - sx; }
      else if (lag < 0) { len = -lag < g.T ? -lag : g.T; sx = g.T - len; }
      double sxy = 0.0, sxx = 0.0, syy = 0.0;
      int cnt = 0, nx = 0, ny = 0;
      for (int t = lane; t < len; t += 32) {
        const double a = x[sx + t], b = y[t];
        const bool va = a == a, vb = b == b;
        if (va) { sxx = fma(a, a, sxx); ++nx; }
        if (vb) { syy = fma(b, b, syy); ++ny; }
        if (va && vb) { sxy = fma(a, b, sxy); ++cnt; }
      }
      sxy = x_warp_sum(sxy); sxx = x_warp_sum(sxx); syy = x_warp_sum(syy);
      cnt = x_warp_sum(cnt); nx = x_warp_sum(nx); ny = x_warp_sum(ny);
      double r = qnan;
      const double den = sqrt(sxx) * sqrt(syy);
      if (cnt > 0 && nx > 0 && ny > 0 && den > 0.0 && !(g.tau > 0 && cnt < g.tau)) r = sxy / den;
      if (lane == 0) g.xcor[(long long)l * g.ncell + c] = r;
      if (r == r && fabs(r) > best_abs) { best_abs = fabs(r); best = l; best_val = r; }  // first maximum, like argmax
    }
    if (lane == 0) {
      if (g.best_idx) g.best_idx[c] = best;
      if (g.best_xcor) g.best_xcor[c] = best_val;
    }
    __syncwarp();
  }
}

extern "C" int ck_xcor_lags(const double* z1, const double* z2, ck_i64 ncell, int T, const int* lags, int nlag, int tau,
                            int detrend, double* xcor, int* best_idx, double* best_xcor, void* stream) {
  CK_REQUIRE(ncell >= 0 && T >= 0 && nlag >= 0, "negative size");
  if (ncell == 0 || nlag == 0) return CK_OK;
  CK_REQUIRE(z1 && z2 && lags && xcor, "null pointer");
  CK_REQUIRE(nlag <= X_MAX_LAGS, "at most %d lags per call (got %d)", X_MAX_LAGS, nlag);
  CK_REQUIRE(T >= 1 && T <= 12000, "time axis length %d out of range (1..12000)", T);
  XcorArgs g;
  g.z1 = z1; g.z2 = z2; g.ncell = ncell; g.T = T; g.nlag = nlag; g.tau = tau; g.detrend = detrend;
  for (int l = 0; l < nlag; ++l) g.lags[l] = lags[l];
  g.xcor = xcor; g.best_idx = best_idx; g.best_xcor = best_xcor;
  const size_t smem = (size_t)X_WARPS * 2 * T * sizeof(double);
  static CkPerDevice attr;
  CK_SET_SMEM_ONCE(attr, ck_xcor_kernel, 200 * 1024);
  long long blocks = (ncell + X_WARPS - 1) / X_WARPS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ck_xcor_kernel<<<(unsigned)blocks, X_THREADS, smem, ck_stream(stream)>>>(g);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

// ck_math.cuh -- scalar FP64 math of the cokriging hot path (distances, Matern correlation, K_nu).
//
// Every function is `CK_HD` (host+device) so that the SAME source that the sm_100a kernels inline
// can be instantiated by g++ in tests/hostmath (numerical validation of the math against
// scipy/mpmath in the CPU-only test tier).  The product library never calls the host instantiation.
//
// Semantics follow the reference (paths relative to /root/reference):
//   distances   src/fields.py:318-342   (sklearn haversine formula x 6371, scipy cdist Euclidean)
//   Matern      src/model.py:354-385    (Rasmussen-Williams parametrisation, h==0 -> 1, non-finite -> 0, clamp >= 0)
//   scaling     src/model.py:193-207    (sigma^2 * rho (+ nugget where h == 0);  rho_ij * prod(sigma) * rho)
// K_nu is evaluated with Temme's series (x <= 2) / Steed's CF2 continued fraction (x > 2) and upward
// recurrence; the reference gets it from scipy.special.kv (AMOS), which also flushes to exactly 0
// once ln K_nu(x) < -700.9217936944459 (probed), reproduced here.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CK_HD __host__ __device__ __forceinline__
#else
#define CK_HD inline
#endif

// Round-to-nearest primitives that the compiler may NOT contract into FMAs: the reference's
// distance arithmetic (numpy / scipy C / Cython on x86-64) is unfused, and bin decisions and the
// h == 0 nugget test depend on the exact bits.
#if defined(__CUDA_ARCH__)
#define CK_MUL(a, b) __dmul_rn((a), (b))
#define CK_ADD(a, b) __dadd_rn((a), (b))
#define CK_SUB(a, b) __dsub_rn((a), (b))
#define CK_DIV(a, b) __ddiv_rn((a), (b))
#define CK_SQRT(a) __dsqrt_rn((a))
#else  // host instantiation is compiled with -ffp-contract=off
#define CK_MUL(a, b) ((a) * (b))
#define CK_ADD(a, b) ((a) + (b))
#define CK_SUB(a, b) ((a) - (b))
#define CK_DIV(a, b) ((a) / (b))
#define CK_SQRT(a) sqrt((a))
#endif

#define CK_EARTH_RADIUS 6371.0            /* src/fields.py:17 */
#define CK_DEG2RAD 0.017453292519943295   /* numpy.radians: x * (pi/180) */
#define CK_KV_UNDERFLOW 3.9222272510438e-305 /* exp(-700.9217936944459): AMOS `elim` cut-off */

#ifndef CK_METRIC_EUCLID
#define CK_METRIC_EUCLID 0
#define CK_METRIC_HAVERSINE 1
#endif
enum { CK_NU_HALF = 0, CK_NU_3HALF = 1, CK_NU_5HALF = 2, CK_NU_7HALF = 3, CK_NU_GENERIC = 4 };

// One point prepared for pair-distance evaluation.
//   Euclid:     a = x, b = y, c unused
//   Haversine:  a = lat [rad], b = lon [rad], c = cos(lat)      (input rows are [lat, lon] degrees)
struct CkPoint {
  double a, b, c;
};

CK_HD CkPoint ck_prepare_point(int metric, double p0, double p1) {
  CkPoint q;
  if (metric == CK_METRIC_HAVERSINE) {
    q.a = CK_MUL(p0, CK_DEG2RAD);
    q.b = CK_MUL(p1, CK_DEG2RAD);
    q.c = cos(q.a);
  } else {
    q.a = p0;
    q.b = p1;
    q.c = 0.0;
  }
  return q;
}

// scipy cdist 'euclidean' for 2 columns: sqrt((x1-x2)^2 + (y1-y2)^2), unfused -> bit-identical.
CK_HD double ck_dist_euclid(const CkPoint& p, const CkPoint& q) {
  const double dx = CK_SUB(p.a, q.a);
  const double dy = CK_SUB(p.b, q.b);
  return CK_SQRT(CK_ADD(CK_MUL(dx, dx), CK_MUL(dy, dy)));
}

// sklearn haversine_distances (same operation order) times the Earth radius in km.
CK_HD double ck_dist_haversine(const CkPoint& p, const CkPoint& q) {
  const double s0 = sin(CK_MUL(0.5, CK_SUB(p.a, q.a)));
  const double s1 = sin(CK_MUL(0.5, CK_SUB(p.b, q.b)));
  const double t = CK_MUL(CK_MUL(CK_MUL(p.c, q.c), s1), s1);
  const double a = CK_ADD(CK_MUL(s0, s0), t);
  return CK_MUL(CK_MUL(2.0, asin(CK_SQRT(a))), CK_EARTH_RADIUS);
}

template <int METRIC>
CK_HD double ck_dist(const CkPoint& p, const CkPoint& q) {
  return METRIC == CK_METRIC_HAVERSINE ? ck_dist_haversine(p, q) : ck_dist_euclid(p, q);
}

// ------------------------------------------------------------------------------------------------
#define CK_KNU_TT 24  /* Temme terms with tabulated reciprocals (x <= 2 converges in <= 13) */
#define CK_KNU_CT 64  /* continued-fraction steps with tabulated reciprocals */
#define CK_KNU_CN 14  /* Chebyshev coefficients per segment and order (x > 2) */
// Matern parameters for one (i, j) block; filled on the host by ck_matern_setup (ck_matern_setup.h).
struct CkMatern {
  double scale;      // sigma_i^2  or  rho_ij * prod(sigma)
  double nugget;     // added where h == 0 exactly (0 for cross blocks / use_nugget=False)
  double len_scale;  // l
  double sqrt2nu;    // sqrt(2 nu)
  double xscale;     // sqrt(2 nu) / l   (fast assembly path: x = h * xscale)
  double x_cut;      // closed-form orders: smallest x with K_nu(x) below the AMOS cut-off (rho flushes to exactly 0 beyond)
  double nu;
  double lc;         // (1 - nu) ln 2 - lgamma(nu)
  // generic-nu constants (Temme): nu = nl + mu, |mu| <= 1/2
  double mu, mu2;
  double temme_g1, temme_g2;    // (1/G(1-mu) -/+ 1/G(1+mu)) / (2mu | 2)
  double rgam_plus, rgam_minus;  // 1/Gamma(1+mu), 1/Gamma(1-mu)
  double mu_pi_ratio;          // pi mu / sin(pi mu)
  int nl;
  int mode;  // CK_NU_*
  // reciprocal tables of the K_nu iterations (depend on mu and the iteration index only; filled by ck_matern_setup):
  // they replace 4 of the 4 divisions per Temme term and 2 of the 3 per continued-fraction step by multiplications
  double t_rq[CK_KNU_TT];   // 1 / (i^2 - mu^2)
  double t_ri[CK_KNU_TT];   // 1 / i
  double t_rm[CK_KNU_TT];   // 1 / (i - mu)
  double t_rp[CK_KNU_TT];   // 1 / (i + mu)
  double c_ra[CK_KNU_CT];   // 1 / a_i,  a_i = -(1/4 - mu^2) - i (i - 1)
  double c_cc[CK_KNU_CT];   // -a_i / i
  // x > 2: Chebyshev expansions in t = 2 / x of g_o(t) = sqrt(x) e^x K_o(x), o = mu and mu + 1, on the three segments
  // t in [0, 1/4], [1/4, 1/2], [1/2, 1] (x >= 8, 4..8, 2..4); CK_KNU_CN coefficients each reach ~2e-15 (tools/cheb_knu.py).
  // Fitted per block on the host from the continued fraction (ck_matern_setup); cheb_ok = 0 selects the continued fraction.
  double cheb[2][3][CK_KNU_CN];
  int cheb_ok;
};

// reciprocal inside the continued fraction: approximate reciprocal + two Newton steps on the device (<= 1 ulp; the
// recurrence is self-correcting), IEEE division on the host
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double ck_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  return fma(fma(-x, r, 1.0), r, r);
}
#define CK_RCP(x) ck_rcp((x))
#else
#define CK_RCP(x) (1.0 / (x))
#endif
#define CK_KNU_EPS 1.0e-16
#define CK_KNU_MAXIT 400

// Modified Bessel function of the second kind K_nu(x), x > 0, nu = P.nl + P.mu, |mu| <= 1/2:
//   K_mu, K_mu+1 (k_lo, k_hi) from N. M. Temme's series (J. Comput. Phys. 19 (1975) 324, eqs. 1.3-1.9) for x <= 2,
//   from per-block Chebyshev fits of sqrt(x) e^x K(x) in 2 / x (x > 2; fallback: Steed's continued fraction CF2),
//   then the forward recurrence K_(o+1) = (2 o / x) K_o + K_(o-1), which is stable for K.
// All constants that depend on mu alone (Temme's Gamma_1, Gamma_2, the reciprocal Gamma values, the reciprocals of the
// series / continued-fraction indices, the Chebyshev coefficients) come from ck_matern_setup (host, long double).
CK_HD double ck_besselk(const CkMatern& P, double x) {
  const double mu = P.mu, mu2 = P.mu2;
  const double xi = 1.0 / x, xi2 = 2.0 * xi;
  double k_lo, k_hi;
  if (x <= 2.0) {  // Temme's series: K_mu = sum c_k f_k, K_mu+1 = (2 / x) sum c_k (p_k - k f_k), c_k = (x^2 / 4)^k / k!
    const double b = 0.5 * x;
    double d = -log(b);
    double e = mu * d;
    const double fact2 = (fabs(e) < 1.0e-6) ? 1.0 + e * e * (1.0 / 6.0) : sinh(e) / e;
    double ff = P.mu_pi_ratio * (P.temme_g1 * cosh(e) + P.temme_g2 * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / P.rgam_plus;
    double q = 0.5 / (e * P.rgam_minus);
    double c = 1.0;
    d = b * b;
    double sum1 = p;
    for (int i = 1; i <= CK_KNU_MAXIT; ++i) {
      const double fi = (double)i;
      if (i < CK_KNU_TT) {
        ff = (fi * ff + p + q) * P.t_rq[i];
        c *= d * P.t_ri[i];
        p *= P.t_rm[i];
        q *= P.t_rp[i];
      } else {
        ff = (fi * ff + p + q) / (fi * fi - mu2);
        c *= d / fi;
        p /= (fi - mu);
        q /= (fi + mu);
      }
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - fi * ff);
      if (fabs(del) < fabs(sum) * CK_KNU_EPS) break;
    }
    k_lo = sum;
    k_hi = sum1 * xi2;
  } else if (P.cheb_ok) {  // Chebyshev expansions of sqrt(x) e^x K(x) in t = 2 / x (Clenshaw, both orders at once)
    const double t = 2.0 * xi;
    const int seg = t <= 0.25 ? 0 : (t <= 0.5 ? 1 : 2);
    const double u = seg == 2 ? fma(4.0, t, -3.0) : fma(8.0, t, seg == 0 ? -1.0 : -3.0);
    const double u2 = 2.0 * u;
    double b1 = 0.0, b2 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int j = CK_KNU_CN - 1; j >= 1; --j) {
      const double nb = fma(u2, b1, P.cheb[0][seg][j] - b2);
      const double nd = fma(u2, d1, P.cheb[1][seg][j] - d2);
      b2 = b1; b1 = nb;
      d2 = d1; d1 = nd;
    }
    const double g0 = fma(u, b1, P.cheb[0][seg][0] - b2);
    const double g1 = fma(u, d1, P.cheb[1][seg][0] - d2);
    const double w = exp(-x) * sqrt(xi);  // e^-x / sqrt(x)
    k_lo = g0 * w;
    k_hi = g1 * w;
  } else {  // Steed's algorithm, continued fraction CF2
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - mu2;
    double q = a1, c = a1;
    double a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i <= CK_KNU_MAXIT; ++i) {
      a -= 2.0 * (double)(i - 1);
      double qnew;
      if (i < CK_KNU_CT) {
        c *= P.c_cc[i];
        qnew = (q1 - b * q2) * P.c_ra[i];
      } else {
        c = -a * c / (double)i;
        qnew = (q1 - b * q2) / a;
      }
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = CK_RCP(b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels) < fabs(s) * CK_KNU_EPS) break;
    }
    h = a1 * h;
    k_lo = sqrt(1.5707963267948966 * xi) * exp(-x) / s;
    k_hi = k_lo * (mu + x + 0.5 - h) * xi;
  }
  for (int i = 1; i <= P.nl; ++i) {
    const double t = (mu + (double)i) * xi2 * k_hi + k_lo;
    k_lo = k_hi;
    k_hi = t;
  }
  return k_lo;
}

// K_nu(x) < AMOS cut-off for the closed-form orders (mode = CK_NU_HALF .. CK_NU_7HALF): scipy's kv returns exactly 0 there
CK_HD bool ck_knu_half_underflows(int mode, double x) {
  const double xi = 1.0 / x;
  double pk;
  if (mode == CK_NU_HALF) pk = 1.0;
  else if (mode == CK_NU_3HALF) pk = 1.0 + xi;
  else if (mode == CK_NU_5HALF) pk = 1.0 + xi * (3.0 + 3.0 * xi);
  else pk = 1.0 + xi * (6.0 + xi * (15.0 + 15.0 * xi));
  return sqrt(1.5707963267948966 * xi) * exp(-x) * pk < CK_KV_UNDERFLOW;
}

// Matern correlation rho(h) with the reference's conventions (src/model.py:354-385).
template <int MODE>
CK_HD double ck_matern_corr(const CkMatern& P, double h) {
  h = fabs(h);
  if (!(h > 0.0)) return 1.0;  // h == 0 (and NaN: the reference's `h > 0` mask is False) -> 1
  const double x = CK_MUL(P.sqrt2nu, CK_DIV(h, P.len_scale));  // reference order: (h / l) then * sqrt(2 nu)
  double rho;
  if (MODE == CK_NU_GENERIC) {
    const double k = ck_besselk(P, x);
    rho = (k < CK_KV_UNDERFLOW) ? 0.0 : exp(P.lc + P.nu * log(x)) * k;
  } else {
    double poly;
    if (MODE == CK_NU_HALF) poly = 1.0;
    else if (MODE == CK_NU_3HALF) poly = 1.0 + x;
    else if (MODE == CK_NU_5HALF) poly = 1.0 + x * (1.0 + x * (1.0 / 3.0));
    else poly = 1.0 + x * (1.0 + x * (0.4 + x * (1.0 / 15.0)));
    rho = poly * exp(-x);
    if (x > 697.0 && ck_knu_half_underflows(MODE, x)) rho = 0.0;  // AMOS underflow guard of scipy.special.kv
  }
  if (!(fabs(rho) <= 1.7976931348623157e308)) rho = 0.0;  // non-finite -> 0
  return rho > 0.0 ? rho : 0.0;
}

// sigma^2 * rho (+ nugget where h == 0)  /  rho_ij * prod(sigma) * rho     (src/model.py:193-207)
template <int MODE>
CK_HD double ck_matern_cov(const CkMatern& P, double h) {
  double c = P.scale * ck_matern_corr<MODE>(P, h);
  if (h == 0.0) c += P.nugget;
  return c;
}

CK_HD double ck_matern_cov_dyn(const CkMatern& P, double h) {
  switch (P.mode) {
    case CK_NU_HALF: return ck_matern_cov<CK_NU_HALF>(P, h);
    case CK_NU_3HALF: return ck_matern_cov<CK_NU_3HALF>(P, h);
    case CK_NU_5HALF: return ck_matern_cov<CK_NU_5HALF>(P, h);
    case CK_NU_7HALF: return ck_matern_cov<CK_NU_7HALF>(P, h);
    default: return ck_matern_cov<CK_NU_GENERIC>(P, h);
  }
}

// ------------------------------------------------------------------------------------------------
// Generic nu in the ASSEMBLY kernel: piecewise Chebyshev table of g(x) = rho(x) e^x, rho(x) = 2^(1-nu)/Gamma(nu) x^nu K_nu(x).
// K_nu costs ~500 FP64 instructions per entry (series / fits of two orders + recurrence + log + two exp), data dependent
// and divergent between the x <= 2 and x > 2 lanes of a warp.  g is analytic and slowly varying for x > 0, so on quarter-
// octave segments [2^e (1 + q/4), 2^e (1 + (q+1)/4)) a degree-11 Chebyshev fit is accurate to ~1e-14 (Bernstein-ellipse
// parameter ~20); the segment index is the biased exponent and the top two mantissa bits of x -- integer work only -- and
// the local variable u in [-1, 1) comes from the remaining mantissa bits with one FMA.  Per entry: 12 shared-memory loads,
// a 12-term Clenshaw recurrence and one exp(-x): ~70 instructions.  The table (one per model block, 8.4 KB) is fitted on
// the host (ck_matern_setup.h) from the long-double continued fraction (x >= 2) / the Temme series (x < 2).
// x below 2^CK_TAB_EMIN (h of a few metres at l = 500 km) takes the series path; x >= x_cut is the kv underflow flush.
// ------------------------------------------------------------------------------------------------
#define CK_TAB_EMIN (-12)
#define CK_TAB_EMAX 10
#define CK_TAB_NSEG (4 * (CK_TAB_EMAX - CK_TAB_EMIN))
#define CK_TAB_NC 12
struct CkMaternTable {
  double c[CK_TAB_NC][CK_TAB_NSEG];  // coefficient-major: lanes in different segments read different banks
};

// segment index of x > 0 (may be < 0 or >= CK_TAB_NSEG) and the local variable u in [-1, 1)
CK_HD int ck_tab_locate(double x, double* u) {
  long long bits;
  memcpy(&bits, &x, sizeof(bits));
  const int hi = (int)(bits >> 32);
  const int idx = (hi >> 18) - ((1023 + CK_TAB_EMIN) << 2);  // 4 * (exponent - EMIN) + top two mantissa bits
  const int q = (hi >> 18) & 3;
  const long long ybits = (bits & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL;  // mantissa with exponent 0: y in [1, 2)
  double y;
  memcpy(&y, &ybits, sizeof(y));
  *u = fma(8.0, y, -9.0 - 2.0 * (double)q);  // f = 4 (y - 1) - q in [0, 1), u = 2 f - 1
  return idx;
}

// Matern correlation from the table (tab = &T.c[0][0], stride CK_TAB_NSEG between coefficients)
CK_HD double ck_matern_corr_tab(const CkMatern& P, const double* tab, double h);  // defined after ck_fast_exp_neg

// ------------------------------------------------------------------------------------------------
// Fast, branch-free variants for the covariance-ASSEMBLY kernel (K1, half-integer nu).  The reference-order
// functions above reproduce the reference's operation order (bit-identical Euclidean distances, libm-level
// haversine) and are what the distance output, the variogram binning and the local-neighbourhood kernels use.
// The assembly kernel only owes 1e-12 relative on covariance entries (north star), so it may use range-limited
// polynomials (coefficients: Chebyshev-node fits generated with mpmath at 60 digits, tools/fit_k1_polys.py),
// FMA contraction and one Newton-corrected reciprocal square root: ~90 FP64 instructions per haversine +
// Matern(3/2) entry instead of ~150 with library sin/asin/exp call sequences, and no calls or slow-path branches,
// so the 16 independent entries of a thread interleave.  Every piece is accurate to <= 2 ulp relative, distances
// to <= 6e-16 relative (tests/test_hostmath.py), h == 0 exactly for identical points.
// ------------------------------------------------------------------------------------------------
#define CK_FMA(a, b, c) fma((a), (b), (c))

// Polynomial coefficients.  On the device they live in constant memory so that every DFMA takes its coefficient
// as a constant-bank operand; as literals each 64-bit coefficient costs two extra (uniform-datapath) move
// instructions per use, which made the kernel issue-bound (profiles/r01l K1 notes).
struct CkFastCoef {
  double sin_q[8];    // sin(y) = y + y z Q(z), z = y^2, |y| <= pi/2
  double asin_r[13];  // asin(x) = x (1 + t R(t)), t = x^2 <= 1/4
  double exp_e[10];   // e^r = 1 + r + r^2 E(r), |r| <= ln2 / 2
};
#define CK_FAST_COEF_INIT                                                                                                      \
  {{-0x1.5555555555555p-3, 0x1.1111111111107p-7, -0x1.a01a01a018aaap-13, 0x1.71de3a54564dfp-19, -0x1.ae6455a1c0795p-26,       \
    0x1.612401540ed0fp-33, -0x1.ae51366b753a9p-41, 0x1.89a43bea5b135p-49},                                                    \
   {0x1.5555555555556p-3, 0x1.3333333332ec4p-4, 0x1.6db6db6e3292ep-5, 0x1.f1c71c1d4aa0dp-6, 0x1.6e8bb1d89492bp-6,              \
    0x1.1c4d343fa57f6p-6, 0x1.c9cf3759fd8ffp-7, 0x1.78246f32d075cp-7, 0x1.524ea8b3aad31p-7, 0x1.653a8f8cfd43cp-8,              \
    0x1.1d661531b222ep-6, -0x1.e7a24f548c99fp-7, 0x1.d78189767524ep-6},                                                       \
   {0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d37p-5, 0x1.11111111109a6p-7, 0x1.6c16c1789d1d7p-10,            \
    0x1.a01a01a7cebcdp-13, 0x1.a019b8ca26fcfp-16, 0x1.71de0d85293b8p-19, 0x1.2891d4ffbb0f9p-22, 0x1.af390ba7e6f47p-26}}
#if defined(__CUDACC__)
static __constant__ CkFastCoef ck_fast_coef_dev = CK_FAST_COEF_INIT;
#endif
static const CkFastCoef ck_fast_coef_host = CK_FAST_COEF_INIT;
#if defined(__CUDA_ARCH__)
#define CK_COEF(arr, i) (ck_fast_coef_dev.arr[i])
#else
#define CK_COEF(arr, i) (ck_fast_coef_host.arr[i])
#endif

// sin(y) for |y| <= pi/2:  y + y z Q(z), z = y^2, Q of degree 7 (approximation error 0.011 ulp)
CK_HD double ck_fast_sin(double y) {
  const double z = y * y;
  double q = CK_COEF(sin_q, 7);
#pragma unroll
  for (int i = 6; i >= 0; --i) q = CK_FMA(q, z, CK_COEF(sin_q, i));
  return CK_FMA(y * z, q, y);
}

// sqrt(a) for a >= 0 (a below 1e-300 counts as 0): approximate reciprocal square root + two coupled
// Goldschmidt steps + one residual correction (<= 1 ulp)
CK_HD double ck_fast_sqrt(double a) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double g = a * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  g = fma(fma(-g, g, a), h, g);
  return a > 1.0e-300 ? g : 0.0;
#else
  return a > 1.0e-300 ? sqrt(a) : 0.0;
#endif
}

// asin(sqrt(a)) for 0 <= a <= 1 without branches.  With R of degree 12, asin(x) = x (1 + x^2 R(x^2)) on |x| <= 1/2:
//   a <= 1/4        : sqrt(a) (1 + a R(a))
//   1/4 < a <= 3/4  : pi/4 + u/2 (1 + u^2 R(u^2)),  u = 2a - 1            (asin(sqrt(a)) = pi/4 + asin(2a - 1)/2)
//   a > 3/4         : pi/2 - sqrt(1 - a) (1 + (1 - a) R(1 - a))
CK_HD double ck_fast_asin_sqrt(double a) {
  a = a < 1.0 ? a : 1.0;
  const bool lo = a <= 0.25, mid = !lo && a <= 0.75;
  const double om = 1.0 - a, u = CK_FMA(2.0, a, -1.0);
  const double t = lo ? a : (mid ? u * u : om);
  const double root = ck_fast_sqrt(lo ? a : om);
  const double base = lo ? root : (mid ? 0.5 * u : -root);
  const double off = lo ? 0.0 : (mid ? 0.78539816339744830962 : 1.57079632679489661923);
  double r = CK_COEF(asin_r, 12);
#pragma unroll
  for (int i = 11; i >= 0; --i) r = CK_FMA(r, t, CK_COEF(asin_r, i));
  return CK_FMA(base, CK_FMA(t, r, 1.0), off);  // off + base (1 + t R(t))
}

// exp(-x) for 0 <= x <= 700:  2^n e^r, n = rint(-x log2 e), |r| <= ln2 / 2, e^r = 1 + r + r^2 E(r), E of degree 9
CK_HD double ck_fast_exp_neg(double x) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: the low word of (t + magic) is rint(t) in two's complement
  const double t = CK_FMA(x, -1.4426950408889634074, magic);
  const double nf = t - magic;
  double r = CK_FMA(nf, -6.93147180369123816490e-01, -x);
  r = CK_FMA(nf, -1.90821492927058770002e-10, r);
  double e = CK_COEF(exp_e, 9);
#pragma unroll
  for (int i = 8; i >= 0; --i) e = CK_FMA(e, r, CK_COEF(exp_e, i));
  const double p = CK_FMA(r * r, e, r) + 1.0;
#if defined(__CUDA_ARCH__)
  const int n = __double2loint(t);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));  // p in [0.70, 1.42], n >= -1010: normal
#else
  return ldexp(p, (int)nf);
#endif
}

CK_HD double ck_matern_corr_tab(const CkMatern& P, const double* tab, double h) {
  h = fabs(h);
  if (!(h > 0.0)) return 1.0;  // h == 0 and NaN (the reference's `h > 0` mask is False): rho = 1
  const double x = CK_MUL(P.sqrt2nu, CK_DIV(h, P.len_scale));  // reference order: (h / l) then * sqrt(2 nu)
  if (!(x < P.x_cut)) return 0.0;                              // kv underflow flush (and +inf)
  double u;
  const int idx = ck_tab_locate(x, &u);
  if (idx < 0 || idx >= CK_TAB_NSEG) return ck_matern_corr<CK_NU_GENERIC>(P, h);  // x < 2^EMIN: series
  const double u2 = 2.0 * u;
  double b1 = 0.0, b2 = 0.0;
#pragma unroll
  for (int j = CK_TAB_NC - 1; j >= 1; --j) {
    const double nb = CK_FMA(u2, b1, tab[j * CK_TAB_NSEG + idx] - b2);
    b2 = b1;
    b1 = nb;
  }
  const double g = CK_FMA(u, b1, tab[idx] - b2);
  const double rho = g * ck_fast_exp_neg(x);
  return rho > 0.0 ? rho : 0.0;
}

CK_HD double ck_dist_euclid_fast(const CkPoint& p, const CkPoint& q) {
  const double dx = p.a - q.a, dy = p.b - q.b;
  return ck_fast_sqrt(CK_FMA(dx, dx, dy * dy));
}

// haversine km, branch-free.  sin^2 is even and pi-periodic, so half the longitude difference is reduced by the
// nearest multiple of pi (two-term pi, exact product for |k| < 2^20); half the latitude difference is within
// [-pi/2, pi/2] for valid latitudes and goes through the same reduction for safety.
CK_HD double ck_reduce_pi(double y) {
  const double magic = 6755399441055744.0;
  const double k = CK_FMA(y, 0.31830988618379067154, magic) - magic;  // rint(y / pi)
  return CK_FMA(k, -1.2246467991473532e-16, CK_FMA(k, -3.14159265358979311600, y));
}
CK_HD double ck_dist_haversine_fast(const CkPoint& p, const CkPoint& q) {
  const double s0 = ck_fast_sin(ck_reduce_pi(0.5 * (p.a - q.a)));
  const double s1 = ck_fast_sin(ck_reduce_pi(0.5 * (p.b - q.b)));
  const double a = CK_FMA(p.c * q.c, s1 * s1, s0 * s0);
  return ck_fast_asin_sqrt(a) * (2.0 * CK_EARTH_RADIUS);
}

template <int METRIC>
CK_HD double ck_dist_fast(const CkPoint& p, const CkPoint& q) {
  return METRIC == CK_METRIC_HAVERSINE ? ck_dist_haversine_fast(p, q) : ck_dist_euclid_fast(p, q);
}

// Haversine with PER-POINT half-angle trigonometry: sin((x1 - x2) / 2) = sin(x1/2) cos(x2/2) - cos(x1/2) sin(x2/2),
// so a pair costs two products and a difference per angle instead of a range reduction + a degree-15 sine polynomial:
// ~70 -> ~45 FP64 instructions per distance.  The two products are rounded separately (no FMA contraction), hence the
// difference is EXACTLY 0 for identical points (the h == 0 nugget decision) -- also across the antimeridian, since
// sin^2 of the half difference is pi-periodic.  Cancellation leaves an ABSOLUTE error of ~4e-16 in each half-angle sine,
// i.e. <= ~7e-12 km in the distance (2e-12 relative at 3 km, 1e-15 at 5000 km): inside the 1e-12 relative tolerance of
// the covariance entries (rho'(h) h / rho -> 0 as h -> 0), but outside the guard band of the variogram bin decisions --
// so only the ASSEMBLY kernel (K1) uses it; K2 / K4 keep the direct forms.
struct CkPointH {
  double sa, ca;  // sin(lat / 2), cos(lat / 2)
  double sb, cb;  // sin(lon / 2), cos(lon / 2)
  double c;       // cos(lat)
};

CK_HD CkPointH ck_prepare_point_h(double lat_deg, double lon_deg) {
  CkPointH q;
  const double a = CK_MUL(lat_deg, CK_DEG2RAD), b = CK_MUL(lon_deg, CK_DEG2RAD);
  q.sa = sin(0.5 * a);
  q.ca = cos(0.5 * a);
  q.sb = sin(0.5 * b);
  q.cb = cos(0.5 * b);
  q.c = cos(a);
  return q;
}

CK_HD double ck_dist_haversine_pre(const CkPointH& p, const CkPointH& q) {
  const double s0 = CK_SUB(CK_MUL(p.sa, q.ca), CK_MUL(p.ca, q.sa));
  const double s1 = CK_SUB(CK_MUL(p.sb, q.cb), CK_MUL(p.cb, q.sb));
  const double a = CK_FMA(p.c * q.c, s1 * s1, s0 * s0);
  return ck_fast_asin_sqrt(a) * (2.0 * CK_EARTH_RADIUS);
}

// sigma^2 rho(h) (+ nugget where h == 0) for the closed-form orders nu in {1/2, 3/2, 5/2, 7/2}, branch-free:
// h == 0 -> sigma^2 + nugget; NaN -> sigma^2 (the reference's `h > 0` mask is False, rho stays 1); x >= x_cut -> 0
// (scipy's kv underflow flush; x_cut is found on the host by bisecting the reference-order predicate).
template <int MODE>
CK_HD double ck_matern_cov_fast(const CkMatern& P, double h) {
  const double x = fabs(h) * P.xscale;
  const double xe = x < 705.0 ? x : 705.0;
  double poly;
  if (MODE == CK_NU_HALF) poly = 1.0;
  else if (MODE == CK_NU_3HALF) poly = 1.0 + xe;
  else if (MODE == CK_NU_5HALF) poly = CK_FMA(xe, CK_FMA(xe, 1.0 / 3.0, 1.0), 1.0);
  else poly = CK_FMA(xe, CK_FMA(xe, CK_FMA(xe, 1.0 / 15.0, 0.4), 1.0), 1.0);
  double c = P.scale * (poly * ck_fast_exp_neg(xe));
  c = x < P.x_cut ? c : 0.0;              // also maps +inf to 0
  c = x > 0.0 ? c : P.scale;              // x == 0 and NaN: rho = 1
  return x == 0.0 ? c + P.nugget : c;
}

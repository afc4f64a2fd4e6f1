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
// Matern parameters for one (i, j) block; filled on the host by ck_matern_setup (ck_matern_setup.h).
struct CkMatern {
  double scale;      // sigma_i^2  or  rho_ij * prod(sigma)
  double nugget;     // added where h == 0 exactly (0 for cross blocks / use_nugget=False)
  double len_scale;  // l
  double sqrt2nu;    // sqrt(2 nu)
  double nu;
  double lc;         // (1 - nu) ln 2 - lgamma(nu)
  // generic-nu constants (Temme): nu = nl + mu, |mu| <= 1/2
  double mu, mu2;
  double gam1, gam2;    // (1/G(1-mu) -/+ 1/G(1+mu)) / (2mu | 2)
  double gampl, gammi;  // 1/Gamma(1+mu), 1/Gamma(1-mu)
  double pimu;          // pi mu / sin(pi mu)
  int nl;
  int mode;  // CK_NU_*
};

#define CK_KNU_EPS 1.0e-16
#define CK_KNU_MAXIT 400

// Modified Bessel function of the second kind K_nu(x), x > 0, nu = P.nl + P.mu.
CK_HD double ck_besselk(const CkMatern& P, double x) {
  const double mu = P.mu, mu2 = P.mu2;
  const double xi = 1.0 / x, xi2 = 2.0 * xi;
  double rkmu, rk1;
  if (x <= 2.0) {  // Temme's series for K_mu, K_mu+1
    const double b = 0.5 * x;
    double d = -log(b);
    double e = mu * d;
    const double fact2 = (fabs(e) < 1.0e-6) ? 1.0 + e * e * (1.0 / 6.0) : sinh(e) / e;
    double ff = P.pimu * (P.gam1 * cosh(e) + P.gam2 * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / P.gampl;
    double q = 0.5 / (e * P.gammi);
    double c = 1.0;
    d = b * b;
    double sum1 = p;
    for (int i = 1; i <= CK_KNU_MAXIT; ++i) {
      const double fi = (double)i;
      ff = (fi * ff + p + q) / (fi * fi - mu2);
      c *= d / fi;
      p /= (fi - mu);
      q /= (fi + mu);
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - fi * ff);
      if (fabs(del) < fabs(sum) * CK_KNU_EPS) break;
    }
    rkmu = sum;
    rk1 = sum1 * xi2;
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
      c = -a * c / (double)i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels) < fabs(s) * CK_KNU_EPS) break;
    }
    h = a1 * h;
    rkmu = sqrt(1.5707963267948966 * xi) * exp(-x) / s;
    rk1 = rkmu * (mu + x + 0.5 - h) * xi;
  }
  for (int i = 1; i <= P.nl; ++i) {
    const double t = (mu + (double)i) * xi2 * rk1 + rkmu;
    rkmu = rk1;
    rk1 = t;
  }
  return rkmu;
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
    if (x > 697.0) {  // AMOS underflow guard of scipy.special.kv, applied to K_nu itself
      const double xi = 1.0 / x;
      double pk;
      if (MODE == CK_NU_HALF) pk = 1.0;
      else if (MODE == CK_NU_3HALF) pk = 1.0 + xi;
      else if (MODE == CK_NU_5HALF) pk = 1.0 + xi * (3.0 + 3.0 * xi);
      else pk = 1.0 + xi * (6.0 + xi * (15.0 + 15.0 * xi));
      if (sqrt(1.5707963267948966 * xi) * exp(-x) * pk < CK_KV_UNDERFLOW) rho = 0.0;
    }
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

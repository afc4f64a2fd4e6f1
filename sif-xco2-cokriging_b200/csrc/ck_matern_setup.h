// ck_matern_setup.h -- host-side preparation of the per-block Matern constants (CkMatern).
// Pure host C++ (no CUDA); shared by the C-ABI (ck_api.cu) and the host validation build
// (tests/hostmath).  Parameter meaning follows src/model.py:188-207 of the reference.
#pragma once
#include <math.h>
#include <string.h>
#include "ck_math.cuh"

// Taylor coefficients of 1/Gamma(1+z) about z = 0 (generated with mpmath at 40 digits).
static const double CK_RGAMMA1P[27] = {
    1.0,
    0.5772156649015328606065121,
    -0.6558780715202538810770195,
    -0.04200263503409523552900393,
    0.1665386113822914895017008,
    -0.0421977345555443367482083,
    -0.009621971527876973562114922,
    0.00721894324666309954239501,
    -0.001165167591859065112113971,
    -0.00021524167411495097281573,
    0.0001280502823881161861531986,
    -0.00002013485478078823865568939,
    -0.000001250493482142670657345359,
    0.00000113302723198169588237413,
    -0.0000002056338416977607103450154,
    6.116095104481415817862499e-9,
    5.002007644469222930055665e-9,
    -1.181274570487020144588127e-9,
    1.04342671169110051049154e-10,
    7.782263439905071254049937e-12,
    -3.696805618642205708187816e-12,
    5.100370287454475979015481e-13,
    -2.05832605356650678322243e-14,
    -5.348122539423017982370017e-15,
    1.226778628238260790158894e-15,
    -1.181259301697458769513765e-16,
    1.186692254751600332579777e-18};

// g_mu(x) = sqrt(x) e^x K_mu(x) and g_(mu+1)(x) for x >= 2 by Steed's continued fraction (CF2) in long double: the
// node values of the Chebyshev fits below (no exponential involved, so nodes at x ~ 2500 are as accurate as at x = 2)
static inline void ck_knu_cf2_nodes(long double mu, long double x, long double* g0, long double* g1) {
  const long double a1 = 0.25L - mu * mu;
  long double b = 2.0L * (1.0L + x), d = 1.0L / b, h = d, delh = d, q1 = 0.0L, q2 = 1.0L;
  long double q = a1, c = a1, a = -a1, s = 1.0L + q * delh;
  for (int i = 2; i <= 5000; ++i) {
    a -= 2.0L * (long double)(i - 1);
    c = -a * c / (long double)i;
    const long double qnew = (q1 - b * q2) / a;
    q1 = q2;
    q2 = qnew;
    q += c * qnew;
    b += 2.0L;
    d = 1.0L / (b + a * d);
    delh = (b * d - 1.0L) * delh;
    h += delh;
    const long double dels = q * delh;
    s += dels;
    if (fabsl(dels) < fabsl(s) * 1.0e-19L) break;
  }
  h = a1 * h;
  *g0 = sqrtl(1.57079632679489661923132169163975144L) / s;
  *g1 = *g0 * (mu + x + 0.5L - h) / x;
}

// returns 0 on success, -1 on invalid parameters
static inline int ck_matern_setup(CkMatern* P, double scale, double nu, double len_scale, double nugget) {
  if (!(nu > 0.0) || !(len_scale > 0.0) || !(nu < 1.0e3)) return -1;
  P->scale = scale;
  P->nugget = nugget;
  P->len_scale = len_scale;
  P->nu = nu;
  P->sqrt2nu = sqrt(2.0 * nu);
  P->xscale = P->sqrt2nu / len_scale;
  P->lc = (1.0 - nu) * log(2.0) - lgamma(nu);
  if (nu == 0.5) P->mode = CK_NU_HALF;
  else if (nu == 1.5) P->mode = CK_NU_3HALF;
  else if (nu == 2.5) P->mode = CK_NU_5HALF;
  else if (nu == 3.5) P->mode = CK_NU_7HALF;
  else P->mode = CK_NU_GENERIC;
  P->x_cut = 1.0e308;  // generic nu: set after the K_nu constants below (the predicate needs them)
  if (P->mode != CK_NU_GENERIC) {  // smallest double x at which the reference-order underflow predicate fires
    double lo = 600.0, hi = 760.0;  // predicate false at lo, true at hi
    for (int it = 0; it < 200 && nextafter(lo, hi) < hi; ++it) {
      const double mid = lo + 0.5 * (hi - lo);
      if (ck_knu_half_underflows(P->mode, mid)) hi = mid; else lo = mid;
    }
    P->x_cut = hi;
  }
  const int nl = (int)(nu + 0.5);
  const long double mu = (long double)nu - (long double)nl;
  const long double m2 = mu * mu;
  long double even = 0.0L, odd = 0.0L;  // sum c_2k mu^2k, sum c_(2k+1) mu^2k
  for (int k = 26; k >= 0; k -= 2) even = even * m2 + (long double)CK_RGAMMA1P[k];
  odd = 0.0L;
  for (int k = 25; k >= 1; k -= 2) odd = odd * m2 + (long double)CK_RGAMMA1P[k];
  P->nl = nl;
  P->mu = (double)mu;
  P->mu2 = (double)m2;
  P->temme_g1 = (double)(-odd);             // (1/G(1-mu) - 1/G(1+mu)) / (2 mu)
  P->temme_g2 = (double)even;               // (1/G(1-mu) + 1/G(1+mu)) / 2
  P->rgam_plus = (double)(even + mu * odd); // 1/Gamma(1+mu)
  P->rgam_minus = (double)(even - mu * odd); // 1/Gamma(1-mu)
  for (int i = 0; i < CK_KNU_TT; ++i) {
    const long double fi = (long double)i;
    P->t_rq[i] = i ? (double)(1.0L / (fi * fi - m2)) : 0.0;
    P->t_ri[i] = i ? (double)(1.0L / fi) : 0.0;
    P->t_rm[i] = i ? (double)(1.0L / (fi - mu)) : 0.0;
    P->t_rp[i] = i ? (double)(1.0L / (fi + mu)) : 0.0;
  }
  for (int i = 0; i < CK_KNU_CT; ++i) {
    const long double ai = -(0.25L - m2) - (long double)i * (long double)(i - 1);  // a after the step-i decrement
    P->c_ra[i] = i >= 2 ? (double)(1.0L / ai) : 0.0;
    P->c_cc[i] = i >= 2 ? (double)(-ai / (long double)i) : 0.0;
  }
  P->cheb_ok = 0;
  if (P->mode == CK_NU_GENERIC) {
    // the fit depends on mu only and costs ~0.4 ms of host time: keep the last few (the three blocks of a bivariate model
    // are set up again for every assembly, cross-covariance and local-prediction call of one parameter vector)
    struct ChebCache { double mu; int valid; double cheb[2][3][CK_KNU_CN]; };
    static thread_local ChebCache cache[8];
    static thread_local int next_slot = 0;
    const ChebCache* hit = nullptr;
    for (int i = 0; i < 8; ++i)
      if (cache[i].valid && cache[i].mu == (double)mu) hit = &cache[i];
    if (!hit) {
      static const double seg_lo[3] = {0.0, 0.25, 0.5}, seg_hi[3] = {0.25, 0.5, 1.0};
      const long double pi = 3.14159265358979323846264338327950288L;
      static thread_local long double cosjk[CK_KNU_CN][CK_KNU_CN];
      static thread_local int cos_ready = 0;
      if (!cos_ready) {
        for (int j = 0; j < CK_KNU_CN; ++j)
          for (int k = 0; k < CK_KNU_CN; ++k) cosjk[j][k] = cosl(pi * (long double)j * ((long double)k + 0.5L) / CK_KNU_CN);
        cos_ready = 1;
      }
      ChebCache* slot = &cache[next_slot];
      next_slot = (next_slot + 1) % 8;
      slot->valid = 0;
      for (int sg = 0; sg < 3; ++sg) {
        long double y[2][CK_KNU_CN];
        for (int k = 0; k < CK_KNU_CN; ++k) {
          const long double t = 0.5L * (seg_lo[sg] + seg_hi[sg]) + 0.5L * (seg_hi[sg] - seg_lo[sg]) * cosjk[1][k];
          ck_knu_cf2_nodes(mu, 2.0L / t, &y[0][k], &y[1][k]);
        }
        for (int o = 0; o < 2; ++o)
          for (int j = 0; j < CK_KNU_CN; ++j) {
            long double acc = 0.0L;
            for (int k = 0; k < CK_KNU_CN; ++k) acc += y[o][k] * cosjk[j][k];
            slot->cheb[o][sg][j] = (double)((j == 0 ? 1.0L : 2.0L) * acc / CK_KNU_CN);
          }
      }
      slot->mu = (double)mu;
      slot->valid = 1;
      hit = slot;
    }
    memcpy(P->cheb, hit->cheb, sizeof(P->cheb));
    P->cheb_ok = 1;
  }
  const long double pm = 3.14159265358979323846264338327950288L * mu;
  P->mu_pi_ratio = (fabsl(pm) < 1.0e-9L) ? 1.0 : (double)(pm / sinl(pm));
  if (P->mode == CK_NU_GENERIC) {  // smallest double x with K_nu(x) below the AMOS cut-off (K_nu is decreasing in x)
    double lo = 600.0, hi = 800.0;
    for (int it = 0; it < 200 && nextafter(lo, hi) < hi; ++it) {
      const double mid = lo + 0.5 * (hi - lo);
      if (ck_besselk(*P, mid) < CK_KV_UNDERFLOW) hi = mid; else lo = mid;
    }
    P->x_cut = hi;
  }
  return 0;
}

// Piecewise Chebyshev table of g(x) = rho(x) e^x for a generic-nu block (see ck_math.cuh).  Node values: x >= 2 from the
// long-double continued fraction (sqrt(x) e^x K_mu, K_mu+1 without any exponential) and the forward recurrence in long
// double; x < 2 from the double-precision Temme series.  Cached by nu (a bivariate model has three blocks, re-created for
// every call of one parameter vector).  Returns 0, or -1 if P is not a generic-nu block.
static inline int ck_matern_table_setup(const CkMatern& P, CkMaternTable* T) {
  if (P.mode != CK_NU_GENERIC) return -1;
  struct TabCache { double nu; int valid; CkMaternTable tab; };
  static thread_local TabCache cache[6];
  static thread_local int next_slot = 0;
  for (int i = 0; i < 6; ++i)
    if (cache[i].valid && cache[i].nu == P.nu) { memcpy(T, &cache[i].tab, sizeof(*T)); return 0; }
  const long double pi = 3.14159265358979323846264338327950288L;
  const long double nu = (long double)P.nu, mu = (long double)P.mu;
  const long double lc = (1.0L - nu) * logl(2.0L) - lgammal(nu);
  auto g_of = [&](long double x) -> long double {
    if (x >= 2.0L) {
      long double k0, k1;  // sqrt(x) e^x K_mu, K_(mu+1)
      ck_knu_cf2_nodes(mu, x, &k0, &k1);
      for (int i = 1; i <= P.nl; ++i) {
        const long double t = (mu + (long double)i) * (2.0L / x) * k1 + k0;
        k0 = k1;
        k1 = t;
      }
      return expl(lc + (nu - 0.5L) * logl(x)) * k0;  // c x^nu K_nu e^x = c x^(nu - 1/2) [sqrt(x) e^x K_nu]
    }
    const double xd = (double)x;
    return expl(lc + nu * logl(x) + x) * (long double)ck_besselk(P, xd);
  };
  TabCache* slot = &cache[next_slot];
  next_slot = (next_slot + 1) % 6;
  slot->valid = 0;
  long double cosjk[CK_TAB_NC][CK_TAB_NC];
  for (int j = 0; j < CK_TAB_NC; ++j)
    for (int k = 0; k < CK_TAB_NC; ++k) cosjk[j][k] = cosl(pi * (long double)j * ((long double)k + 0.5L) / CK_TAB_NC);
  for (int sg = 0; sg < CK_TAB_NSEG; ++sg) {
    const int e = CK_TAB_EMIN + sg / 4, q = sg % 4;
    const long double a = ldexpl(1.0L + 0.25L * q, e), b = ldexpl(1.0L + 0.25L * (q + 1), e);
    long double y[CK_TAB_NC];
    for (int k = 0; k < CK_TAB_NC; ++k) y[k] = g_of(0.5L * (a + b) + 0.5L * (b - a) * cosjk[1][k]);
    for (int j = 0; j < CK_TAB_NC; ++j) {
      long double acc = 0.0L;
      for (int k = 0; k < CK_TAB_NC; ++k) acc += y[k] * cosjk[j][k];
      slot->tab.c[j][sg] = (double)((j == 0 ? 1.0L : 2.0L) * acc / CK_TAB_NC);
    }
  }
  slot->nu = P.nu;
  slot->valid = 1;
  memcpy(T, &slot->tab, sizeof(*T));
  return 0;
}

// Host instantiation of csrc/ck_math.cuh for NUMERICAL VALIDATION ONLY (tests tier, no GPU).
// Built by __graft_entry__.build() / tests/conftest.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC ck_hostmath.cpp -o _ck_hostmath.so
// It lets the CPU test tier compare the exact scalar code the sm_100a kernels inline against
// scipy/mpmath.  It is NOT part of the product library and is never used as a fallback.
#include "../../sif-xco2-cokriging_b200/csrc/ck_math.cuh"
#include "../../sif-xco2-cokriging_b200/csrc/ck_matern_setup.h"

extern "C" {

int ckh_besselk(double nu, const double* x, long n, double* out) {
  CkMatern P;
  if (ck_matern_setup(&P, 1.0, nu, 1.0, 0.0)) return -1;
  for (long i = 0; i < n; ++i) out[i] = ck_besselk(P, x[i]);
  return 0;
}

int ckh_matern_cov(double scale, double nu, double len_scale, double nugget, const double* h, long n, double* out) {
  CkMatern P;
  if (ck_matern_setup(&P, scale, nu, len_scale, nugget)) return -1;
  for (long i = 0; i < n; ++i) out[i] = ck_matern_cov_dyn(P, h[i]);
  return 0;
}

// generic nu through the piecewise Chebyshev table of the assembly kernel (ck_matern_corr_tab)
int ckh_matern_cov_table(double scale, double nu, double len_scale, double nugget, const double* h, long n, double* out) {
  CkMatern P;
  if (ck_matern_setup(&P, scale, nu, len_scale, nugget)) return -1;
  static CkMaternTable T;
  if (ck_matern_table_setup(P, &T)) return -2;
  for (long i = 0; i < n; ++i) {
    double c = P.scale * ck_matern_corr_tab(P, &T.c[0][0], h[i]);
    if (h[i] == 0.0) c += P.nugget;
    out[i] = c;
  }
  return 0;
}

int ckh_distance(int metric, const double* X1, long n1, const double* X2, long n2, double* out) {
  for (long i = 0; i < n1; ++i) {
    const CkPoint p = ck_prepare_point(metric, X1[2 * i], X1[2 * i + 1]);
    for (long j = 0; j < n2; ++j) {
      const CkPoint q = ck_prepare_point(metric, X2[2 * j], X2[2 * j + 1]);
      out[i * n2 + j] = metric == CK_METRIC_HAVERSINE ? ck_dist_haversine(p, q) : ck_dist_euclid(p, q);
    }
  }
  return 0;
}

// fast assembly-path variants (K1, half-integer nu)
int ckh_distance_fast(int metric, const double* X1, long n1, const double* X2, long n2, double* out) {
  for (long i = 0; i < n1; ++i) {
    const CkPoint p = ck_prepare_point(metric, X1[2 * i], X1[2 * i + 1]);
    for (long j = 0; j < n2; ++j) {
      const CkPoint q = ck_prepare_point(metric, X2[2 * j], X2[2 * j + 1]);
      out[i * n2 + j] = metric == CK_METRIC_HAVERSINE ? ck_dist_haversine_fast(p, q) : ck_dist_euclid_fast(p, q);
    }
  }
  return 0;
}

int ckh_distance_pre(const double* X1, long n1, const double* X2, long n2, double* out) {
  for (long i = 0; i < n1; ++i) {
    const CkPointH p = ck_prepare_point_h(X1[2 * i], X1[2 * i + 1]);
    for (long j = 0; j < n2; ++j) out[i * n2 + j] = ck_dist_haversine_pre(p, ck_prepare_point_h(X2[2 * j], X2[2 * j + 1]));
  }
  return 0;
}

int ckh_matern_cov_fast(double scale, double nu, double len_scale, double nugget, const double* h, long n, double* out) {
  CkMatern P;
  if (ck_matern_setup(&P, scale, nu, len_scale, nugget)) return -1;
  for (long i = 0; i < n; ++i) {
    switch (P.mode) {
      case CK_NU_HALF: out[i] = ck_matern_cov_fast<CK_NU_HALF>(P, h[i]); break;
      case CK_NU_3HALF: out[i] = ck_matern_cov_fast<CK_NU_3HALF>(P, h[i]); break;
      case CK_NU_5HALF: out[i] = ck_matern_cov_fast<CK_NU_5HALF>(P, h[i]); break;
      case CK_NU_7HALF: out[i] = ck_matern_cov_fast<CK_NU_7HALF>(P, h[i]); break;
      default: return -2;
    }
  }
  return 0;
}

int ckh_fast_pieces(const double* x, long n, double* sin_out, double* asin_sqrt_out, double* exp_neg_out) {
  for (long i = 0; i < n; ++i) {
    sin_out[i] = ck_fast_sin(x[i]);
    asin_sqrt_out[i] = ck_fast_asin_sqrt(x[i]);
    exp_neg_out[i] = ck_fast_exp_neg(x[i]);
  }
  return 0;
}
}

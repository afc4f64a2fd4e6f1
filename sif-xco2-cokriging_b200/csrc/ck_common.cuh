// ck_common.cuh -- internal helpers shared by the translation units of libcokrig_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "../../include/cokrig.h"
#include "ck_math.cuh"
#include "ck_matern_setup.h"

#define CK_NB 128  // Cholesky panel width == diagonal block size == GEMM K depth per step

void ck_set_error(const char* fmt, ...);

#define CK_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      ck_set_error(__VA_ARGS__);   \
      return CK_ERR_ARG;           \
    }                              \
  } while (0)

#define CK_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ck_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return CK_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// every kernel launch of the library is counted (bench.py reports it as `gpu_launches`)
void ck_count_launches(int n);
#define CK_LAUNCH_CHECK_N(n)        \
  do {                              \
    ck_count_launches(n);           \
    CK_CUDA(cudaGetLastError());    \
  } while (0)
#define CK_LAUNCH_CHECK() CK_LAUNCH_CHECK_N(1)

// The reference's parameter matrices unpacked from the flat vector (src/model.py:130-152).
struct CkParams {
  int n_procs;
  double sigma[2], nugget[2];
  double nu[2][2], len_scale[2][2];  // symmetric fill
  double rho01;
  double sigma_prod;  // np.nanprod(sigma matrix) -- product of ALL marginal sigmas (src/model.py:203-207)
};

int ck_unpack_params(const double* params, int n_procs, CkParams* out);
// Block (i, j) of the joint model: i == j -> covariance(i, ., use_nugget), else cross_covariance(i, j, .)
int ck_block_matern(const CkParams& p, int i, int j, int use_nugget, CkMatern* out);

// One block of distances (value = 0) or covariances (value = 1) from coordinates (ck_matern.cu).
int ck_block_launch(const double* xy1, ck_i64 n1, const double* xy2, ck_i64 n2, int metric, const CkMatern& P, int value,
                    double* out, ck_i64 ld, double* out_t, ck_i64 ld_t, int symmetric, cudaStream_t st);

// INT8 update kernel (ck_ozaki.cu): SM count of the current device
int ck_oz_num_sms();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: every launcher keeps one flag per device
// (a process may drive several GPUs) and sets the attribute the first time it launches on each of them.
struct CkPerDevice {
  bool done[64] = {};
};
#define CK_SET_SMEM_ONCE(flags, kernel, bytes)                                                                  \
  do {                                                                                                          \
    int dev_ = 0;                                                                                               \
    CK_CUDA(cudaGetDevice(&dev_));                                                                              \
    if (dev_ < 0 || dev_ >= 64 || !(flags).done[dev_]) {                                                        \
      CK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));         \
      if (dev_ >= 0 && dev_ < 64) (flags).done[dev_] = true;                                                    \
    }                                                                                                           \
  } while (0)

static inline cudaStream_t ck_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

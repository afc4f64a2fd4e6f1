/* mg_cokrige.c -- the multi-GPU joint cokriging sweep driven from plain C through include/cokrig.h
 * (ck_mg_create / ck_mg_joint_cov / ck_mg_potrf / ck_mg_potrs_predict), one process per GPU.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/mg_cokrige.c \
 *       -L sif-xco2-cokriging_b200/cokrig_b200 -lcokrig_b200 -L /usr/local/cuda/lib64 -lcudart -lm -o mg_cokrige
 *   # one process per GPU; rank 0 writes the NCCL id to a file the other ranks read:
 *   for r in 0 1; do ./mg_cokrige $r 2 /tmp/ck_id & done; wait        # optional: n0 n1 m after the id path
 *
 * It solves a synthetic bivariate system (the reference's src/joint_prediction.py:50-78 for one large system) and
 * prints the first predictions; every rank prints the same numbers.  tests/test_abi.py compiles this file (syntax and
 * prototypes only) so that the header stays valid C. */
#define _DEFAULT_SOURCE /* usleep */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include "cokrig.h"

#define CHECK(call)                                                         \
  do {                                                                      \
    int rc_ = (call);                                                       \
    if (rc_ != CK_OK) {                                                     \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ck_last_error()); \
      exit(1);                                                              \
    }                                                                       \
  } while (0)

static double* to_device(const double* h, size_t n) {
  double* d = NULL;
  if (cudaMalloc((void**)&d, n * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    fprintf(stderr, "device copy failed\n");
    exit(1);
  }
  return d;
}

int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 0, world = argc > 2 ? atoi(argv[2]) : 1;
  const char* id_path = argc > 3 ? argv[3] : "/tmp/ck_mg_id";
  const ck_i64 n0 = argc > 4 ? atoll(argv[4]) : 6000, n1 = argc > 5 ? atoll(argv[5]) : 5000, m = argc > 6 ? atoll(argv[6]) : 2000;
  /* sigma11, sigma22, nu11, nu12, nu22, len11, len12, len22, nugget11, nugget22, rho12  (src/model.py:130-152) */
  const double params[11] = {1.0, 0.8, 1.5, 1.5, 1.5, 0.1, 0.1, 0.1, 0.02, 0.02, -0.2};
  unsigned char id[128] = {0};
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (ndev < 1) {
    fprintf(stderr, "no CUDA device: there is no CPU fallback\n");
    return 1;
  }
  cudaSetDevice(rank % ndev);

  /* replicated synthetic inputs (same seed on every rank) */
  double* h = (double*)malloc(sizeof(double) * (size_t)(2 * (n0 + n1 + m) + n0 + n1));
  srand(7);
  for (ck_i64 i = 0; i < 2 * (n0 + n1 + m) + n0 + n1; ++i) h[i] = rand() / (double)RAND_MAX;
  double* xy0 = to_device(h, (size_t)(2 * n0));
  double* xy1 = to_device(h + 2 * n0, (size_t)(2 * n1));
  double* xyp = to_device(h + 2 * (n0 + n1), (size_t)(2 * m));
  double* z = to_device(h + 2 * (n0 + n1 + m), (size_t)(n0 + n1));

  /* the 128-byte NCCL id travels through a file here; MPI_Bcast or a socket does as well */
  if (world > 1) {
    if (rank == 0) {
      CHECK(ck_mg_unique_id(id));
      FILE* f = fopen(id_path, "wb");
      fwrite(id, 1, sizeof id, f);
      fclose(f);
    } else {
      FILE* f = NULL;
      while (!(f = fopen(id_path, "rb")) || fread(id, 1, sizeof id, f) != sizeof id) {
        if (f) fclose(f);
        usleep(100000);
      }
      fclose(f);
    }
  }

  ck_mg_ctx* ctx = NULL;
  const int Q = world >= 4 && world % 2 == 0 ? 2 : 1;
  CHECK(ck_mg_create(&ctx, world, rank, world / Q, Q, 1024, world > 1 ? id : NULL));
  const size_t bytes = ck_mg_workspace_bytes(ctx, n0 + n1, m);
  void* ws = NULL;
  double *pred = NULL, *var = NULL, logdet_h = 0.0, times[3];
  int *info = NULL, info_h = 0;
  cudaMalloc(&ws, bytes);
  cudaMalloc((void**)&pred, sizeof(double) * (size_t)(m + 1));
  cudaMalloc((void**)&var, sizeof(double) * (size_t)m);
  cudaMalloc((void**)&info, sizeof(int));

  CHECK(ck_mg_joint_cov(ctx, xy0, n0, xy1, n1, xyp, m, z, params, 2, /*i_pred*/ 0, CK_METRIC_EUCLID, ws, bytes, NULL));
  CHECK(ck_mg_potrf(ctx, NULL));
  CHECK(ck_mg_potrs_predict(ctx, pred, var, info, NULL));
  CHECK(ck_mg_logdet(ctx, pred + m, NULL));
  CHECK(ck_mg_times_ms(ctx, times));

  double out[4] = {0, 0, 0, 0};
  cudaMemcpy(out, pred, sizeof(double) * (size_t)(m < 4 ? m : 4), cudaMemcpyDeviceToHost);
  cudaMemcpy(&logdet_h, pred + m, sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(&info_h, info, sizeof(int), cudaMemcpyDeviceToHost);
  printf("rank %d/%d: info %d, logdet %.9f, pred[0..3] = %.12f %.12f %.12f %.12f; assemble %.2f ms, sweep %.2f ms, predict %.2f ms, "
         "workspace %.2f GB\n", rank, world, info_h, logdet_h, out[0], out[1], out[2], out[3], times[0], times[1], times[2], bytes / 1e9);

  CHECK(ck_mg_destroy(ctx));
  cudaFree(ws); cudaFree(pred); cudaFree(var); cudaFree(info);
  cudaFree(xy0); cudaFree(xy1); cudaFree(xyp); cudaFree(z);
  free(h);
  return info_h != 0;
}

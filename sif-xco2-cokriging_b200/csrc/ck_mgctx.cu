// ck_mgctx.cu -- the multi-GPU handle API (ck_mg_create / ck_mg_joint_cov / ck_mg_potrf /
// ck_mg_potrs_predict / ck_mg_destroy): the whole 2-D block-cyclic sweep of one large joint cokriging
// system (src/joint_prediction.py:50-78) driven from C, one process per GPU, NCCL over NVLink / NVSwitch.
//
// A context owns three NCCL communicators -- world, my process ROW (fixed p: Q ranks) and my process
// COLUMN (fixed q: P ranks) -- and a high-priority panel stream.  The schedule is the one of
// cokrig_b200/parallel.py (which issues the same per-rank C-ABI calls through torch.distributed):
//
//   per tile column k (right-looking, one-panel look-ahead on the panel stream):
//     owners of column k     bring column k up to date with panel k-1, then
//     diagonal owner         potrf(tile k,k)                         -> ncclBroadcast down the process column
//     process column k%Q     rows I>k:  A_Ik <- A_Ik L_kk^-T         -> ncclBroadcast along each process ROW
//                                                                       (a rank receives only tiles I = p mod P)
//     every rank             panel tiles J = q mod Q (B operand)     <- ncclAllGather / ncclBroadcast inside its
//                                                                       process COLUMN
//     every rank (main)      A_IJ -= L_Ik L_Jk^T on its tiles I>k, J>k, J<=I  (INT8 tensor-core kernel; FP64 DMMA
//                                                                       below its break-even), minus column k+1
//   the target rows (C^T and z) ride along as extra row tiles: after the sweep they hold L^-1 c, L^-1 z.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process this is the copy torch has
// already loaded; a plain C program gets the system library), so the library itself has no link-time
// dependency on it and world = 1 contexts never touch it.  No hidden device allocations: every buffer
// is carved from the workspace the caller passes (ck_mg_workspace_bytes).
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include <vector>

#include "ck_common.cuh"

namespace {

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommSplit) CommSplit = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.lib) return CK_OK;
  const char* names[] = {getenv("CK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names)
    if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) {
    ck_set_error("cannot load NCCL (libnccl.so.2; set CK_NCCL_LIB): %s", dlerror());
    return CK_ERR_UNSUPPORTED;
  }
#define CK_SYM(field, name)                                               \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  if (!g_nccl.field) {                                                    \
    ck_set_error("NCCL symbol %s not found", name);                       \
    return CK_ERR_UNSUPPORTED;                                            \
  }
  CK_SYM(GetUniqueId, "ncclGetUniqueId")
  CK_SYM(CommInitRank, "ncclCommInitRank")
  CK_SYM(CommSplit, "ncclCommSplit")
  CK_SYM(CommDestroy, "ncclCommDestroy")
  CK_SYM(Broadcast, "ncclBroadcast")
  CK_SYM(AllGather, "ncclAllGather")
  CK_SYM(AllReduce, "ncclAllReduce")
  CK_SYM(GetErrorString, "ncclGetErrorString")
#undef CK_SYM
  g_nccl.lib = h;
  return CK_OK;
}

#define CK_NCCL(call)                                                                                  \
  do {                                                                                                 \
    ncclResult_t r_ = (call);                                                                          \
    if (r_ != ncclSuccess) {                                                                           \
      ck_set_error("%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));      \
      return CK_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)
#define CK_TRY(call)         \
  do {                       \
    int rc_ = (call);        \
    if (rc_ != CK_OK) return rc_; \
  } while (0)

inline ck_i64 first_local_after(ck_i64 k, ck_i64 nprocs, ck_i64 rank) {  // smallest l with l * nprocs + rank > k
  return k >= rank ? (k - rank) / nprocs + 1 : 0;
}

__global__ void __launch_bounds__(256) ck_mg_eye_kernel(double* __restrict__ x, int n) {
  const int c = blockIdx.x * 256 + threadIdx.x, r = blockIdx.y;
  if (c < n) x[(size_t)r * n + c] = r == c ? 1.0 : 0.0;
}

// out[r][c] = -in[c][r] (n x n, 32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) ck_mg_neg_transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int n) {
  __shared__ double t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) t[i][tx] = in[(size_t)(by + i) * n + bx + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8) out[(size_t)(bx + i) * n + by + tx] = -t[tx][i];
}

// pred[i] = sum_r allp[r][0][i], var[i] = c0 - sum_r allp[r][1][i], ranks in fixed order
__global__ void __launch_bounds__(256) ck_mg_combine_kernel(const double* __restrict__ allp, int world, ck_i64 m, double c0,
                                                            double* __restrict__ pred, double* __restrict__ var) {
  const ck_i64 i = (ck_i64)blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  double a = 0.0, b = 0.0;
  for (int r = 0; r < world; ++r) {
    a += allp[(size_t)r * 2 * m + i];
    b += allp[(size_t)r * 2 * m + m + i];
  }
  pred[i] = a;
  var[i] = c0 - b;
}

// LAPACK-style info of the whole system from the per-tile-column flags: tile column k failing at local order j -> k*tb + j
__global__ void ck_mg_info_kernel(const int* __restrict__ info, ck_i64 TC, ck_i64 tb, ck_i64 N, int* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  ck_i64 bad = 0;
  for (ck_i64 k = 0; k < TC; ++k)
    if (info[k] > 0) {
      bad = k * tb + info[k];
      break;
    }
  *out = bad > N ? 0 : (int)bad;  // the pad is the identity: cannot fail
}

__global__ void ck_mg_sum_kernel(const double* __restrict__ x, ck_i64 n, double* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  double s = 0.0;
  for (ck_i64 i = 0; i < n; ++i) s += x[i];  // fixed order
  *out = s;
}

}  // namespace

struct ck_mg_ctx {
  int world = 1, rank = 0, P = 1, Q = 1, p = 0, q = 0;
  ck_i64 tb = 1024;
  int device = 0, nsm = 148;
  int lookahead = 1, panel_sms_fixed = -1, int8_min_tiles = 600, panel_min = 24;  // floor of the panel stream's SM share: see parallel._panel_share
  ncclComm_t comm_world = nullptr, comm_row = nullptr, comm_col = nullptr;
  cudaStream_t panel = nullptr;
  cudaEvent_t tev[4] = {};
  // layout of the current problem
  ck_i64 N = 0, m = 0, TC = 0, TE = 0, TR = 0, LRt = 0, LCt = 0, LRmax = 0, cap = 0, ld = 0, pack_len = 0;
  double c0 = 0.0;
  int state = 0;  // 0 created, 1 assembled, 2 factored
  // workspace carve-up
  double *local = nullptr, *stage[2] = {}, *bcols[2] = {}, *xsend = nullptr, *xrecv = nullptr, *packs[2] = {};
  double *xt = nullptr, *negx = nullptr, *part = nullptr, *allp = nullptr, *dl = nullptr;
  int* info = nullptr;
  struct Oz {
    void *fa = nullptr, *fb = nullptr;
    double *sa = nullptr, *sb = nullptr;
  } oz[2];  // [0] main stream, [1] panel stream
  int panel_sms = 0;
};

namespace {

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* ptr = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return ptr;
  }
};

void set_layout(ck_mg_ctx* h, ck_i64 N, ck_i64 m) {
  const ck_i64 tb = h->tb;
  h->N = N;
  h->m = m;
  h->TC = (N + tb - 1) / tb;
  h->cap = tb - 1;
  h->TE = (m + h->cap - 1) / h->cap;
  h->TR = h->TC + h->TE;
  h->LRt = ck_mg_local_tiles(h->TR, h->P, h->p);
  h->LCt = ck_mg_local_tiles(h->TC, h->Q, h->q);
  h->LRmax = ck_mg_local_tiles(h->TR, h->P, 0);
  h->ld = (h->LCt > 0 ? h->LCt : 1) * tb;
  h->pack_len = tb * tb + (ck_i64)(ck_potrf_workspace_bytes(tb) / 8);
}

// carve (ws != NULL) or just measure (ws == NULL) the workspace of the current layout
size_t carve(ck_mg_ctx* h, void* ws) {
  Carver c(ws);
  const size_t tb = (size_t)h->tb, t2 = tb * tb;
  const size_t lrt = (size_t)(h->LRt > 0 ? h->LRt : 1), lct = (size_t)(h->LCt > 0 ? h->LCt : 1);
  const size_t lrmax = (size_t)(h->LRmax > 0 ? h->LRmax : 1);
  h->local = c.take<double>(lrt * tb * lct * tb);
  for (int b = 0; b < 2; ++b) h->stage[b] = c.take<double>(lrmax * t2);
  for (int b = 0; b < 2; ++b) h->bcols[b] = c.take<double>(lct * t2);
  h->xsend = h->P > 1 ? c.take<double>(lct * t2) : nullptr;
  h->xrecv = h->P > 1 ? c.take<double>((size_t)h->P * lct * t2) : nullptr;
  for (int b = 0; b < 2; ++b) h->packs[b] = c.take<double>((size_t)h->pack_len);
  h->xt = c.take<double>(t2);
  h->negx = c.take<double>(t2);
  const size_t mm = (size_t)(h->m > 0 ? h->m : 1);
  h->part = c.take<double>(2 * mm);
  h->allp = c.take<double>((size_t)h->world * 2 * mm);
  h->dl = c.take<double>((size_t)(h->TC > 0 ? h->TC : 1) + 1);
  h->info = c.take<int>((size_t)(h->TC > 0 ? h->TC : 1) + 1);
  for (int s = 0; s < 2; ++s) {
    const ck_i64 rows_a = (ck_i64)(lrt * tb), rows_b = (ck_i64)((s == 0 ? lct : 1) * tb);
    h->oz[s].fa = c.take<char>(ck_oz_slices_bytes(rows_a, h->tb, 0));
    h->oz[s].fb = c.take<char>(ck_oz_slices_bytes(rows_b, h->tb, 1));
    h->oz[s].sa = c.take<double>((size_t)ck_oz_scales_len(rows_a));
    h->oz[s].sb = c.take<double>((size_t)ck_oz_scales_len(rows_b));
  }
  return (c.off + 255) & ~size_t(255);
}

bool int8_ok(const ck_mg_ctx* h, ck_i64 m, ck_i64 n, ck_i64 k) {
  return ck_oz_active((ck_i64)1 << 20) && k % 32 == 0 && k >= 256 && k <= 1024 && (m / 128) * (n / 64) >= h->int8_min_tiles;
}

// SMs this launch may use: the persistent INT8 kernel of the main stream leaves panel_sms SMs to the panel stream
int cta_cap(const ck_mg_ctx* h, int on_panel) {
  const int r = h->panel_sms;
  if (r <= 0) return 0;
  return on_panel ? r : (h->nsm - r > 1 ? h->nsm - r : 1);
}

// C -= A B^T through digit slices (tb > 0: block-cyclic mask of ck_mg_update)
int oz_product(ck_mg_ctx* h, int on_panel, const double* A, ck_i64 lda, ck_i64 m, const double* B, ck_i64 ldb, ck_i64 n, ck_i64 k,
               double* C, ck_i64 ldc, ck_i64 tbm, ck_i64 gi0, ck_i64 gis, ck_i64 gj0, ck_i64 gjs, cudaStream_t st) {
  ck_mg_ctx::Oz& z = h->oz[on_panel];
  CK_TRY(ck_oz_split(A, lda, m, k, z.fa, nullptr, z.sa, st));
  CK_TRY(ck_oz_split(B, ldb, n, k, nullptr, z.fb, z.sb, st));
  const int cap = cta_cap(h, on_panel);
  if (tbm) return ck_oz_mg_update(z.fa, z.sa, m, z.fb, z.sb, n, k, C, ldc, tbm, gi0, gis, gj0, gjs, cap, st);
  return ck_oz_gemm(z.fa, z.sa, m, z.fb, z.sb, n, k, C, ldc, 0, cap, st);
}

int mg_update(ck_mg_ctx* h, int on_panel, const double* A, ck_i64 m, const double* B, ck_i64 n, double* C, ck_i64 ldc, ck_i64 gi0,
              ck_i64 gis, ck_i64 gj0, ck_i64 gjs, cudaStream_t st) {
  const ck_i64 tb = h->tb;
  if (int8_ok(h, m, n, tb)) return oz_product(h, on_panel, A, tb, m, B, tb, n, tb, C, ldc, tb, gi0, gis, gj0, gjs, st);
  return ck_mg_update(A, tb, B, tb, C, ldc, m, n, tb, tb, gi0, gis, gj0, gjs, st);
}

// rows <- rows L^-T for the (nrows x tb) block `rows` (leading dimension ld); out (nrows x tb, dense) gets a copy
int panel_trsm(ck_mg_ctx* h, const double* pack, double* rows, ck_i64 nrows, ck_i64 ld, double* out, cudaStream_t st) {
  const ck_i64 tb = h->tb;
  const size_t row_bytes = (size_t)tb * sizeof(double);
  if (int8_ok(h, nrows, tb, tb)) {
    // one INT8 product with the explicitly inverted tile: xt = (L^-1)^T (target-major solve of I), negx = -L^-1
    ck_mg_eye_kernel<<<dim3((unsigned)((tb + 255) / 256), (unsigned)tb), 256, 0, st>>>(h->xt, (int)tb);
    CK_LAUNCH_CHECK();
    CK_TRY(ck_trsm_lower(pack, tb, tb, pack + tb * tb, h->xt, tb, tb, st));
    ck_mg_neg_transpose_kernel<<<dim3((unsigned)(tb / 32), (unsigned)(tb / 32)), 256, 0, st>>>(h->xt, h->negx, (int)tb);
    CK_LAUNCH_CHECK();
    CK_CUDA(cudaMemsetAsync(out, 0, (size_t)nrows * row_bytes, st));
    CK_TRY(oz_product(h, 1, rows, ld, nrows, h->negx, tb, tb, tb, out, tb, 0, 0, 1, 0, 1, st));  // out = 0 - rows (-L^-1)^T
    CK_CUDA(cudaMemcpy2DAsync(rows, (size_t)ld * sizeof(double), out, row_bytes, row_bytes, (size_t)nrows,
                              cudaMemcpyDeviceToDevice, st));
    return CK_OK;
  }
  CK_TRY(ck_trsm_lower(pack, tb, tb, pack + tb * tb, rows, nrows, ld, st));
  CK_CUDA(cudaMemcpy2DAsync(out, row_bytes, rows, (size_t)ld * sizeof(double), row_bytes, (size_t)nrows, cudaMemcpyDeviceToDevice,
                            st));
  return CK_OK;
}

// SM share of the panel stream while trailing update k and panel chain k+1 run side by side (the two-term cost model of
// parallel.BlockCyclicCokriging._panel_share, fitted to the C3 traces: profiles/r02_mg_trace_*)
int panel_share(const ck_mg_ctx* h, ck_i64 k) {
  const double rows = (double)(h->TR - k - 1 > 0 ? h->TR - k - 1 : 0), cols = (double)(h->TC - k - 1 > 0 ? h->TC - k - 1 : 0);
  const double w_main = rows * cols / (2.0 * h->P * h->Q);
  const double w_panel = 2.0 * (double)(h->TR - k - 2 > 0 ? h->TR - k - 2 : 0) / h->P;
  const double s = (double)h->tb / 1024.0, tau = 3.65 * s * s * s, fixed = 1.7;
  int best = h->panel_min;
  double best_t = 1e300;
  for (int r = h->panel_min; r < h->nsm - 23; r += 4) {
    const double a = w_main * tau / (h->nsm - r), b = w_panel * tau / r + fixed;
    const double t = a > b ? a : b;
    if (t < best_t) best_t = t, best = r;
  }
  return best;
}

struct Pending {
  ck_i64 k;
  const double *stage, *bcols;
  cudaEvent_t ready;
};

// Panel stream: [apply `pending` to column k on its owners] -> potrf(k,k) -> broadcast down the process column -> TRSM of
// the rows below -> row broadcast -> column exchange.  *ready is recorded when stage_b / bcols_b hold my share of panel k.
int factor_panel(ck_mg_ctx* h, ck_i64 k, double* stage_b, double* bcols_b, double* pack, const Pending* pending,
                 cudaEvent_t stage_free, cudaEvent_t ready) {
  const ck_i64 tb = h->tb, t2 = tb * tb, P = h->P, Q = h->Q, p = h->p, q = h->q, ld = h->ld;
  const ck_i64 qk = k % Q, pk = k % P, ljk = k / Q;
  cudaStream_t st = h->panel;
  if (stage_free) CK_CUDA(cudaStreamWaitEvent(st, stage_free, 0));
  if (pending) {
    CK_CUDA(cudaStreamWaitEvent(st, pending->ready, 0));
    if (q == qk) {  // bring column k up to date with panel k-1 (rows I >= k)
      const ck_i64 li0 = first_local_after(k - 1, P, p);
      if (li0 < h->LRt)
        CK_TRY(mg_update(h, 1, pending->stage + li0 * t2, (h->LRt - li0) * tb, pending->bcols + ljk * t2, tb,
                         h->local + li0 * tb * ld + ljk * tb, ld, li0 * P + p, P, k, Q, st));
    }
  }
  const ck_i64 l0 = first_local_after(k, P, p);
  if (q == qk) {
    if (p == pk) {
      double* tile = h->local + (k / P) * tb * ld + ljk * tb;
      CK_TRY(ck_potrf(tile, tb, ld, pack + t2, h->info + k, st));
      CK_CUDA(cudaMemcpy2DAsync(pack, (size_t)tb * 8, tile, (size_t)ld * 8, (size_t)tb * 8, (size_t)tb, cudaMemcpyDeviceToDevice, st));
    }
    if (P > 1) CK_NCCL(g_nccl.Broadcast(pack, pack, (size_t)h->pack_len, ncclDouble, (int)pk, h->comm_col, st));
    if (l0 < h->LRt) CK_TRY(panel_trsm(h, pack, h->local + l0 * tb * ld + ljk * tb, (h->LRt - l0) * tb, ld, stage_b + l0 * t2, st));
  }
  // A operand: the panel tiles of my process row, from the member of my row that owns tile column k
  if (Q > 1 && l0 < h->LRt)
    CK_NCCL(g_nccl.Broadcast(stage_b + l0 * t2, stage_b + l0 * t2, (size_t)((h->LRt - l0) * t2), ncclDouble, (int)qk, h->comm_row, st));
  // B operand: the panel tiles J of my tile columns (J > k); tile J sits with process row J mod P after the row broadcast
  const ck_i64 lj0 = first_local_after(k, Q, q);
  if (lj0 < h->LCt) {
    const size_t tile_bytes = (size_t)t2 * sizeof(double);
    if (P == 1) {
      for (ck_i64 lj = lj0; lj < h->LCt; ++lj)
        CK_CUDA(cudaMemcpyAsync(bcols_b + lj * t2, stage_b + (lj * Q + q) * t2, tile_bytes, cudaMemcpyDeviceToDevice, st));
    } else {
      // share[pp] = my column tiles held by process row pp, in increasing J; cnt = the largest share
      std::vector<ck_i64> count(P, 0), slot((size_t)(h->LCt - lj0));
      for (ck_i64 lj = lj0; lj < h->LCt; ++lj) {
        const ck_i64 J = lj * Q + q;
        slot[(size_t)(lj - lj0)] = count[(size_t)(J % P)]++;
      }
      ck_i64 cnt = 0, holders = 0, holder = -1;
      for (ck_i64 pp = 0; pp < P; ++pp) {
        if (count[(size_t)pp] > cnt) cnt = count[(size_t)pp];
        if (count[(size_t)pp]) ++holders, holder = pp;
      }
      for (ck_i64 lj = lj0; lj < h->LCt; ++lj) {  // my share, packed
        const ck_i64 J = lj * Q + q;
        if (J % P == p)
          CK_CUDA(cudaMemcpyAsync(h->xsend + slot[(size_t)(lj - lj0)] * t2, stage_b + (J / P) * t2, tile_bytes, cudaMemcpyDeviceToDevice, st));
      }
      if (holders == 1) {
        // P divides Q (e.g. 2 x 4): every tile of my columns sits with ONE process row -> a broadcast from it
        if (p == holder)
          CK_CUDA(cudaMemcpyAsync(bcols_b + lj0 * t2, h->xsend, (size_t)(h->LCt - lj0) * tile_bytes, cudaMemcpyDeviceToDevice, st));
        CK_NCCL(g_nccl.Broadcast(bcols_b + lj0 * t2, bcols_b + lj0 * t2, (size_t)((h->LCt - lj0) * t2), ncclDouble, (int)holder,
                                 h->comm_col, st));
      } else {
        CK_NCCL(g_nccl.AllGather(h->xsend, h->xrecv, (size_t)(cnt * t2), ncclDouble, h->comm_col, st));
        for (ck_i64 lj = lj0; lj < h->LCt; ++lj) {
          const ck_i64 J = lj * Q + q;
          CK_CUDA(cudaMemcpyAsync(bcols_b + lj * t2, h->xrecv + ((J % P) * cnt + slot[(size_t)(lj - lj0)]) * t2, tile_bytes,
                                  cudaMemcpyDeviceToDevice, st));
        }
      }
    }
  }
  CK_CUDA(cudaEventRecord(ready, st));
  return CK_OK;
}

// A_IJ -= L_Ik L_Jk^T on my tiles with I > k, J > k (J <= I), minus column `skip_col` (done by the look-ahead)
int trailing_update(ck_mg_ctx* h, ck_i64 k, const double* stage_b, const double* bcols_b, ck_i64 skip_col, cudaStream_t st) {
  const ck_i64 tb = h->tb, t2 = tb * tb, P = h->P, Q = h->Q;
  const ck_i64 li0 = first_local_after(k, P, h->p);
  ck_i64 lj0 = first_local_after(k, Q, h->q);
  if (skip_col >= 0 && lj0 < h->LCt && lj0 * Q + h->q == skip_col) ++lj0;
  if (li0 >= h->LRt || lj0 >= h->LCt) return CK_OK;
  return mg_update(h, 0, stage_b + li0 * t2, (h->LRt - li0) * tb, bcols_b + lj0 * t2, (h->LCt - lj0) * tb,
                   h->local + li0 * tb * h->ld + lj0 * tb, h->ld, li0 * P + h->p, P, lj0 * Q + h->q, Q, st);
}

}  // namespace

extern "C" int ck_mg_unique_id(void* id128) {
  CK_REQUIRE(id128, "null pointer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  CK_TRY(nccl_load());
  ncclUniqueId id;
  CK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return CK_OK;
}

extern "C" int ck_mg_create(ck_mg_ctx** out, int world, int rank, int P, int Q, ck_i64 tile, const void* nccl_unique_id128) {
  CK_REQUIRE(out, "null pointer");
  *out = nullptr;
  CK_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
  CK_REQUIRE(P >= 1 && Q >= 1 && (ck_i64)P * Q == world, "process grid %dx%d does not match world size %d", P, Q, world);
  CK_REQUIRE(tile >= 128 && tile % 128 == 0 && tile <= 1024, "tile must be a multiple of 128, at most 1024 (got %lld)", (long long)tile);
  CK_REQUIRE(world == 1 || nccl_unique_id128, "a NCCL unique id (ck_mg_unique_id on rank 0, distributed by the host) is required");
  ck_mg_ctx* h = new ck_mg_ctx();
  h->world = world, h->rank = rank, h->P = P, h->Q = Q, h->p = rank / Q, h->q = rank % Q, h->tb = tile;
  auto fail = [&](int rc) {
    delete h;
    return rc;
  };
  if (cudaGetDevice(&h->device) != cudaSuccess) {
    ck_set_error("no CUDA device");
    return fail(CK_ERR_CUDA);
  }
  h->nsm = ck_oz_num_sms();
  if (const char* e = getenv("CK_MG_LOOKAHEAD")) h->lookahead = atoi(e) != 0;
  if (const char* e = getenv("CK_MG_PANEL_SMS")) h->panel_sms_fixed = atoi(e);
  if (const char* e = getenv("CK_MG_INT8_MIN_TILES")) h->int8_min_tiles = atoi(e);
  if (const char* e = getenv("CK_MG_PANEL_MIN")) h->panel_min = atoi(e) > 4 ? atoi(e) : 4;
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->panel, cudaStreamNonBlocking, hi) != cudaSuccess) {
    ck_set_error("cannot create the panel stream: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(CK_ERR_CUDA);
  }
  for (auto& e : h->tev)
    if (cudaEventCreate(&e) != cudaSuccess) return fail(CK_ERR_CUDA);
  if (world > 1) {
    int rc = nccl_load();
    if (rc) return fail(rc);
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&h->comm_world, world, id, rank);
    // row communicator: the Q ranks of my process row, rank inside = q; column communicator: the P ranks of my process
    // column, rank inside = p (every rank takes part in both splits)
    if (r == ncclSuccess) r = g_nccl.CommSplit(h->comm_world, h->p, h->q, &h->comm_row, nullptr);
    if (r == ncclSuccess) r = g_nccl.CommSplit(h->comm_world, h->q, h->p, &h->comm_col, nullptr);
    if (r != ncclSuccess) {
      ck_set_error("NCCL communicator setup failed: %s", g_nccl.GetErrorString(r));
      return fail(CK_ERR_CUDA);
    }
  }
  *out = h;
  return CK_OK;
}

extern "C" int ck_mg_destroy(ck_mg_ctx* h) {
  if (!h) return CK_OK;
  if (h->comm_row) g_nccl.CommDestroy(h->comm_row);
  if (h->comm_col) g_nccl.CommDestroy(h->comm_col);
  if (h->comm_world) g_nccl.CommDestroy(h->comm_world);
  for (auto& e : h->tev)
    if (e) cudaEventDestroy(e);
  if (h->panel) cudaStreamDestroy(h->panel);
  delete h;
  return CK_OK;
}

extern "C" size_t ck_mg_workspace_bytes(ck_mg_ctx* h, ck_i64 n_data, ck_i64 m) {
  if (!h || n_data < 0 || m < 0) return 0;
  ck_mg_ctx tmp = *h;
  set_layout(&tmp, n_data, m);
  return carve(&tmp, nullptr);
}

extern "C" int ck_mg_grid(const ck_mg_ctx* h, int* pq4 /*HOST: P, Q, p, q*/) {
  CK_REQUIRE(h && pq4, "null pointer");
  pq4[0] = h->P, pq4[1] = h->Q, pq4[2] = h->p, pq4[3] = h->q;
  return CK_OK;
}

extern "C" int ck_mg_joint_cov(ck_mg_ctx* h, const double* xy0, ck_i64 n0, const double* xy1, ck_i64 n1, const double* xyp, ck_i64 m,
                               const double* z, const double* params, int n_procs, int i_pred, int metric, void* ws, size_t ws_bytes,
                               void* stream) {
  CK_REQUIRE(h, "null context");
  CK_REQUIRE(n_procs == 1 || n_procs == 2, "n_procs=%d unsupported (1 or 2)", n_procs);
  CK_REQUIRE(params, "params is NULL");
  if (n_procs == 1) n1 = 0;
  CK_REQUIRE(n0 >= 0 && n1 >= 0 && m >= 0 && n0 + n1 > 0, "bad sizes");
  CK_REQUIRE(i_pred >= 0 && i_pred < n_procs, "i_pred out of range");
  set_layout(h, n0 + n1, m);
  const size_t need = carve(h, nullptr);
  CK_REQUIRE(ws && ws_bytes >= need, "workspace too small: %zu bytes given, %zu needed (ck_mg_workspace_bytes)", ws_bytes, need);
  carve(h, ws);
  const double sig = n_procs == 2 ? params[i_pred] : params[0], nug = n_procs == 2 ? params[8 + i_pred] : params[3];
  h->c0 = sig * sig + nug;
  cudaStream_t st = ck_stream(stream);
  CK_CUDA(cudaEventRecord(h->tev[0], st));
  CK_CUDA(cudaMemsetAsync(h->info, 0, (size_t)(h->TC + 1) * sizeof(int), st));
  CK_TRY(ck_mg_assemble(xy0, n0, xy1, n1, xyp, m, z, params, n_procs, i_pred, metric, h->tb, h->P, h->p, h->Q, h->q, h->local, h->ld, st));
  CK_CUDA(cudaEventRecord(h->tev[1], st));
  h->state = 1;
  return CK_OK;
}

extern "C" int ck_mg_potrf(ck_mg_ctx* h, void* stream) {
  CK_REQUIRE(h && h->state >= 1, "ck_mg_joint_cov must come first");
  cudaStream_t mainst = ck_stream(stream);
  const ck_i64 TC = h->TC;
  // events: [0] assembled, then per tile column its panel-ready and main-done events
  std::vector<cudaEvent_t> ev((size_t)(2 * TC + 1));
  for (auto& e : ev) CK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  int rc = CK_OK;
  auto run = [&]() -> int {
    cudaEvent_t assembled = ev[0];
    CK_CUDA(cudaEventRecord(assembled, mainst));  // the panel stream must not touch `local` before the assembly is done
    cudaEvent_t main_done[2] = {nullptr, nullptr};
    h->panel_sms = h->lookahead ? (h->panel_sms_fixed >= 0 ? h->panel_sms_fixed : 40) : 0;
    cudaEvent_t panel_ready = nullptr;
    if (TC) {
      panel_ready = ev[1];
      CK_TRY(factor_panel(h, 0, h->stage[0], h->bcols[0], h->packs[0], nullptr, assembled, panel_ready));
    }
    for (ck_i64 k = 0; k < TC; ++k) {
      const int buf = (int)(k & 1);
      cudaEvent_t nxt = nullptr;
      if (h->lookahead && h->panel_sms_fixed < 0) h->panel_sms = panel_share(h, k);
      if (h->lookahead && k + 1 < TC) {
        // column k+1 first (its owners), then its panel, all on the panel stream
        Pending pend{k, h->stage[buf], h->bcols[buf], panel_ready};
        nxt = ev[(size_t)(1 + 2 * (k + 1))];
        CK_TRY(factor_panel(h, k + 1, h->stage[1 - buf], h->bcols[1 - buf], h->packs[1 - buf], &pend, main_done[1 - buf], nxt));
      }
      CK_CUDA(cudaStreamWaitEvent(mainst, panel_ready, 0));
      CK_TRY(trailing_update(h, k, h->stage[buf], h->bcols[buf], (h->lookahead && k + 1 < TC) ? k + 1 : -1, mainst));
      main_done[buf] = ev[(size_t)(2 + 2 * k)];
      CK_CUDA(cudaEventRecord(main_done[buf], mainst));
      if (!h->lookahead && k + 1 < TC) {
        nxt = ev[(size_t)(1 + 2 * (k + 1))];
        CK_TRY(factor_panel(h, k + 1, h->stage[1 - buf], h->bcols[1 - buf], h->packs[1 - buf], nullptr, main_done[buf], nxt));
      }
      panel_ready = nxt;
    }
    CK_CUDA(cudaEventRecord(h->tev[2], mainst));
    return CK_OK;
  };
  rc = run();
  for (auto& e : ev) cudaEventDestroy(e);  // released by the runtime once the recorded work has completed
  if (rc == CK_OK) h->state = 2;
  return rc;
}

extern "C" int ck_mg_potrs_predict(ck_mg_ctx* h, double* pred_dev, double* var_dev, int* info_dev, void* stream) {
  CK_REQUIRE(h && h->state >= 2, "ck_mg_potrf must come first");
  CK_REQUIRE((h->m == 0 || (pred_dev && var_dev)) && info_dev, "null pointer");
  cudaStream_t st = ck_stream(stream);
  const ck_i64 tb = h->tb, m = h->m, mm = m > 0 ? m : 1;
  CK_CUDA(cudaMemsetAsync(h->part, 0, (size_t)(2 * mm) * sizeof(double), st));
  for (ck_i64 li = 0; li < h->LRt; ++li) {
    const ck_i64 I = li * h->P + h->p;
    if (I < h->TC || h->LCt == 0) continue;
    const ck_i64 t_lo = (I - h->TC) * h->cap, nt = (h->cap < m - t_lo) ? h->cap : m - t_lo;
    const double* rows = h->local + li * tb * h->ld;
    CK_TRY(ck_row_dots(rows, h->ld, nt, h->LCt * tb, rows + (tb - 1) * h->ld, 0, h->part + t_lo, h->part + mm + t_lo, st));
  }
  const double* total = h->part;
  if (h->world > 1) {
    CK_NCCL(g_nccl.AllGather(h->part, h->allp, (size_t)(2 * mm), ncclDouble, h->comm_world, st));
    CK_NCCL(g_nccl.AllReduce(h->info, h->info, (size_t)(h->TC > 0 ? h->TC : 1), ncclInt32, ncclMax, h->comm_world, st));
    total = h->allp;
  }
  if (m > 0) {
    ck_mg_combine_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(total, h->world > 1 ? h->world : 1, mm, h->c0, pred_dev, var_dev);
    CK_LAUNCH_CHECK();
  }
  ck_mg_info_kernel<<<1, 32, 0, st>>>(h->info, h->TC, tb, h->N, info_dev);
  CK_LAUNCH_CHECK();
  CK_CUDA(cudaEventRecord(h->tev[3], st));
  return CK_OK;
}

extern "C" int ck_mg_logdet(ck_mg_ctx* h, double* out_dev, void* stream) {
  CK_REQUIRE(h && h->state >= 2 && out_dev, "ck_mg_potrf must come first");
  cudaStream_t st = ck_stream(stream);
  const ck_i64 tb = h->tb, TC = h->TC;
  CK_CUDA(cudaMemsetAsync(h->dl, 0, (size_t)(TC + 1) * sizeof(double), st));
  for (ck_i64 k = 0; k < TC; ++k)
    if (k % h->P == h->p && k % h->Q == h->q)
      CK_TRY(ck_logdet(h->local + (k / h->P) * tb * h->ld + (k / h->Q) * tb, tb, h->ld, h->dl + k, st));
  if (h->world > 1)  // every entry has exactly one non-zero contributor: exact
    CK_NCCL(g_nccl.AllReduce(h->dl, h->dl, (size_t)TC, ncclDouble, ncclSum, h->comm_world, st));
  ck_mg_sum_kernel<<<1, 32, 0, st>>>(h->dl, TC, out_dev);
  CK_LAUNCH_CHECK();
  return CK_OK;
}

extern "C" int ck_mg_times_ms(ck_mg_ctx* h, double* out3 /*HOST*/) {
  CK_REQUIRE(h && out3 && h->state >= 2, "nothing to report");
  CK_CUDA(cudaEventSynchronize(h->tev[3]));
  for (int i = 0; i < 3; ++i) {
    float ms = 0.f;
    CK_CUDA(cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]));
    out3[i] = ms;
  }
  return CK_OK;
}

extern "C" int ck_mg_local_factor(ck_mg_ctx* h, double** local_dev /*HOST*/, ck_i64* ld /*HOST*/, ck_i64* local_row_tiles /*HOST*/,
                                  ck_i64* local_col_tiles /*HOST*/) {
  CK_REQUIRE(h && h->state >= 1 && local_dev && ld && local_row_tiles && local_col_tiles, "no assembled system");
  *local_dev = h->local, *ld = h->ld, *local_row_tiles = h->LRt, *local_col_tiles = h->LCt;
  return CK_OK;
}

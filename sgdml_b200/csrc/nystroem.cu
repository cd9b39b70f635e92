// Building blocks of the Nystroem-preconditioned CG solver (SURVEY.md section 8 row a-S2) --
// reference sgdml/solvers/iterative.py:208-351 (_nystroem_cholesky_factor), 83-142
// (_init_precon_operator), 414-471 (_cho_factor_stable).
//
// Everything operates on X = K_nm, the (n x m) block of kernel columns at the inducing columns,
// row-major and resident in HBM from assembly to the end of the solve (the reference keeps it in
// host memory and calls SciPy):
//   K_mm = -X[cols, :]                               gather_rows_neg     (iterative.py:253)
//   L_mm = chol(K_mm + eps I)                        sgdml_b200_potrf    (iterative.py:267)
//   X   <- X L_mm^-T                                 trsm_right_lt       (iterative.py:278-287)
//   inner = X^T X + lam I                            gram_tn             (iterative.py:293-295)
//   L    = chol(inner)                               sgdml_b200_potrf    (iterative.py:305-311)
//   X   <- X L^-T          (= B^T, B = L_inv_K_mn)   trsm_right_lt       (iterative.py:337-347)
//   leverage scores = row norms^2 of X               row_sqnorms         (iterative.py:107-109)
//   P v = (X (X^T v) - v)/lam                        nystroem_apply      (iterative.py:136-138)
#include <algorithm>

#include "common.cuh"
#include "solve.cuh"

namespace sgdml {

__global__ void k_gather_rows_neg(const double* __restrict__ X, int64_t ldx, int64_t m, const int64_t* __restrict__ idx,
                                  double* __restrict__ out, int64_t ldo) {
  const int64_t r = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m || c >= m) return;
  out[r * ldo + c] = -X[idx[r] * ldx + c];
}

// tiled transpose: src (rows x cols, lds) -> dst (cols x rows_pad, ldd), padding columns zeroed by the caller
__global__ void k_transpose(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t lds,
                            double* __restrict__ dst, int64_t ldd) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[c * ldd + r] = tile[threadIdx.x][i];
  }
}

__global__ void k_add_diag2(double* __restrict__ A, int64_t n, int64_t lda, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i * lda + i] += v;
}

// out[r] = |X[r, :]|^2 ; one warp per row
__global__ void k_row_sqnorms(const double* __restrict__ X, int64_t n_rows, int64_t m, int64_t ldx,
                              double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  double s = 0.0;
  for (int64_t j = lane; j < m; j += 32) {
    const double v = X[r * ldx + j];
    s = fma(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = s;
}

// part[chunk][j] = sum_{r in chunk} X[r][j] v[r] ; grid.y = row chunks, threads over columns.
// Deterministic two-pass reduction (no atomics): with several ranks every rank must compute
// bit-identical CG scalars, otherwise their iteration counts (and collectives) diverge.
__global__ void __launch_bounds__(256) k_xt_v(const double* __restrict__ X, int64_t n_rows, int64_t m, int64_t ldx,
                                             const double* __restrict__ v, double* __restrict__ part,
                                             int rows_per_cta) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, n_rows);
  if (j >= m) return;
  // four independent accumulators: four loads in flight per thread (fixed order -> deterministic)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 = fma(X[r * ldx + j], v[r], s0);
    s1 = fma(X[(r + 1) * ldx + j], v[r + 1], s1);
    s2 = fma(X[(r + 2) * ldx + j], v[r + 2], s2);
    s3 = fma(X[(r + 3) * ldx + j], v[r + 3], s3);
  }
  for (; r < r1; ++r) s0 = fma(X[r * ldx + j], v[r], s0);
  part[(int64_t)blockIdx.y * m + j] = (s0 + s1) + (s2 + s3);
}

__global__ void k_reduce_chunks(const double* __restrict__ part, int64_t n_chunks, int64_t m, double* __restrict__ t) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double s = 0.0;
  for (int64_t c = 0; c < n_chunks; ++c) s += part[c * m + j];
  t[j] = s;
}

// out[r] = (X[r, :] . t - v[r]) / lam ; one warp per row
__global__ void k_x_t_minus_v(const double* __restrict__ X, int64_t n_rows, int64_t m, int64_t ldx,
                              const double* __restrict__ t, const double* __restrict__ v, double lam_inv,
                              double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const double* __restrict__ x = X + r * ldx;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t j = lane;
  for (; j + 96 < m; j += 128) {
    s0 = fma(x[j], t[j], s0);
    s1 = fma(x[j + 32], t[j + 32], s1);
    s2 = fma(x[j + 64], t[j + 64], s2);
    s3 = fma(x[j + 96], t[j + 96], s3);
  }
  for (; j < m; j += 32) s0 = fma(x[j], t[j], s0);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = (s - v[r]) * lam_inv;
}

}  // namespace sgdml

static const int XTV_ROWS_PER_CTA = 256;
namespace sgdml {
int64_t xtv_chunks(int64_t n_rows) { return (n_rows + XTV_ROWS_PER_CTA - 1) / XTV_ROWS_PER_CTA; }

// out = (X t - v)/lam on device vectors, stream-ordered, no synchronisation (csrc/pcg.cu)
int x_t_minus_v_device(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double lam, const double* t_dev,
                       const double* v_dev, double* out_dev, cudaStream_t s) {
  k_x_t_minus_v<<<ceil_div(n_rows, 8), 256, 0, s>>>(X, n_rows, m, ldx, t_dev, v_dev, 1.0 / lam, out_dev);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  return 0;
}

// t = X^T v (m doubles, device), deterministic two-pass reduction; part: m * xtv_chunks(n_rows) doubles
int xt_v_device(const double* X, int64_t n_rows, int64_t m, int64_t ldx, const double* v_dev, double* t_dev,
                       double* part, cudaStream_t s) {
  const int64_t n_chunks = xtv_chunks(n_rows);
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)n_chunks);
  SG_ARG(grid.y <= 65535);
  k_xt_v<<<grid, 256, 0, s>>>(X, n_rows, m, ldx, v_dev, part, XTV_ROWS_PER_CTA);
  SG_CUDA(cudaGetLastError());
  k_reduce_chunks<<<ceil_div(m, 256), 256, 0, s>>>(part, n_chunks, m, t_dev);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC, 2);
  return 0;
}
}  // namespace sgdml

using namespace sgdml;

extern "C" {

int sgdml_b200_gather_rows_neg(const double* X, int64_t ldx, int64_t m, const int64_t* row_idxs, double* out,
                               int64_t ldo, void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && row_idxs != nullptr && out != nullptr && m >= 1 && ldx >= m && ldo >= m);
  SG_ARG(is_device_ptr(X) && is_device_ptr(out));
  cudaStream_t s = (cudaStream_t)stream;
  Staged sI;
  SG_TRY(sI.init(row_idxs, sizeof(int64_t) * (size_t)m, true, s));
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)std::min<int64_t>(m, 65535));
  SG_ARG(m <= 65535);
  k_gather_rows_neg<<<grid, 256, 0, s>>>(X, ldx, m, (const int64_t*)sI.dev(), out, ldo);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  if (sI.staged()) SG_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int sgdml_b200_add_diag(double* A, int64_t n, int64_t lda, double value, void* stream) {
  SG_TRY(require_device());
  SG_ARG(A != nullptr && n >= 1 && lda >= n && is_device_ptr(A));
  cudaStream_t s = (cudaStream_t)stream;
  k_add_diag2<<<ceil_div(n, 256), 256, 0, s>>>(A, n, lda, value);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  return 0;
}

int sgdml_b200_trsm_right_lt(const double* L, int64_t m, int64_t ldl, double* X, int64_t n_rows, int64_t ldx,
                             void* stream) {
  SG_TRY(require_device());
  SG_ARG(L != nullptr && X != nullptr && m >= 1 && n_rows >= 1 && ldl >= m && ldx >= m);
  SG_ARG(is_device_ptr(L) && is_device_ptr(X));
  return trsm_right_lt_device(L, m, ldl, X, n_rows, ldx, (cudaStream_t)stream);
}

int sgdml_b200_gram_tn(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double lam, double* C, int64_t ldc,
                       void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && C != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && ldc >= m);
  SG_ARG(is_device_ptr(X) && is_device_ptr(C));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t ldt = (n_rows + 1) / 2 * 2;  // even row stride for the aligned DMMA path
  double* Xt = nullptr;
  SG_CUDA(cudaMalloc(&Xt, sizeof(double) * (size_t)m * ldt));
  auto body = [&]() -> int {
    if (ldt != n_rows) SG_CUDA(cudaMemsetAsync(Xt, 0, sizeof(double) * (size_t)m * ldt, s));
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n_rows + 31) / 32));
    SG_ARG((n_rows + 31) / 32 <= 65535);
    k_transpose<<<grid, dim3(32, 8), 0, s>>>(X, n_rows, m, ldx, Xt, ldt);
    SG_CUDA(cudaGetLastError());
    count_launch(KID_MISC);
    GemmArgs g;
    g.m = m;
    g.n = m;
    g.k = ldt;
    g.A = Xt;
    g.lda = ldt;
    g.B = Xt;
    g.ldb = ldt;
    g.C = C;
    g.ldc = ldc;
    g.alpha = 1.0;
    g.beta = 0.0;
    g.mode = 0;
    g.tri = 1;  // lower triangle (all that potrf reads)
    g.abort_flag = nullptr;
    SG_TRY(launch_gemm(g, s));
    k_add_diag2<<<ceil_div(m, 256), 256, 0, s>>>(C, m, ldc, lam);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(s));
    return 0;
  };
  int rc = body();
  cudaFree(Xt);
  return rc;
}

int sgdml_b200_row_sqnorms(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double* out, void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && out != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  Staged sO;
  SG_TRY(sO.init(out, sizeof(double) * (size_t)n_rows, false, s));
  k_row_sqnorms<<<ceil_div(n_rows, 8), 256, 0, s>>>(X, n_rows, m, ldx, (double*)sO.dev());
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  SG_TRY(sO.finish(s));
  if (sO.staged()) SG_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// Scratch buffer of the preconditioner application: P v runs once per CG iteration, so its
// temporaries (staged vectors, per-chunk partial sums) are kept between calls instead of being
// cudaMalloc'ed and cudaFree'd (which synchronises the device) every time.  One per host thread
// and device; grown on demand.
struct Scratch {
  double* p = nullptr;
  size_t cap = 0;  // doubles
  int dev = -1;
  ~Scratch() {}  // freed with the context at process exit
};
static thread_local Scratch g_scratch;

static int scratch_get(size_t n_doubles, double** out) {
  int dev = 0;
  SG_CUDA(cudaGetDevice(&dev));
  if (g_scratch.dev != dev || g_scratch.cap < n_doubles) {
    if (g_scratch.p != nullptr && g_scratch.dev == dev) cudaFree(g_scratch.p);
    g_scratch.p = nullptr;
    g_scratch.cap = 0;
    SG_CUDA(cudaMalloc(&g_scratch.p, sizeof(double) * n_doubles));
    g_scratch.cap = n_doubles;
    g_scratch.dev = dev;
  }
  *out = g_scratch.p;
  return 0;
}

// device view of a possibly-host input vector: host data is copied into `slot` (scratch memory)
static int stage_in(const double* user, size_t n, double* slot, cudaStream_t s, const double** dev) {
  if (is_device_ptr(user)) {
    *dev = user;
    return 0;
  }
  SG_CUDA(cudaMemcpyAsync(slot, user, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  *dev = slot;
  return 0;
}

int sgdml_b200_nystroem_apply(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double lam, const double* v,
                              double* out, void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && v != nullptr && out != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && lam > 0.0);
  SG_ARG(is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)n_rows, mm = (size_t)m;
  double* sc = nullptr;
  SG_TRY(scratch_get(mm + mm * (size_t)xtv_chunks(n_rows) + 2 * n, &sc));
  double *t = sc, *part = sc + mm, *v_slot = part + mm * (size_t)xtv_chunks(n_rows), *out_slot = v_slot + n;
  const double* v_dev = nullptr;
  SG_TRY(stage_in(v, n, v_slot, s, &v_dev));
  const bool out_host = !is_device_ptr(out);
  double* out_dev = out_host ? out_slot : out;
  SG_TRY(xt_v_device(X, n_rows, m, ldx, v_dev, t, part, s));
  k_x_t_minus_v<<<ceil_div(n_rows, 8), 256, 0, s>>>(X, n_rows, m, ldx, t, v_dev, 1.0 / lam, out_dev);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  if (out_host) SG_CUDA(cudaMemcpyAsync(out, out_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  SG_CUDA(cudaStreamSynchronize(s));  // the scratch buffer is reused by the next call
  return 0;
}

int sgdml_b200_nystroem_project(const double* X, int64_t n_rows, int64_t m, int64_t ldx, const double* v, double* t,
                                void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && v != nullptr && t != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)n_rows, mm = (size_t)m;
  double* sc = nullptr;
  SG_TRY(scratch_get(mm + mm * (size_t)xtv_chunks(n_rows) + n, &sc));
  double *t_slot = sc, *part = sc + mm, *v_slot = part + mm * (size_t)xtv_chunks(n_rows);
  const double* v_dev = nullptr;
  SG_TRY(stage_in(v, n, v_slot, s, &v_dev));
  const bool t_host = !is_device_ptr(t);
  double* t_dev = t_host ? t_slot : t;
  SG_TRY(xt_v_device(X, n_rows, m, ldx, v_dev, t_dev, part, s));
  if (t_host) SG_CUDA(cudaMemcpyAsync(t, t_dev, sizeof(double) * mm, cudaMemcpyDeviceToHost, s));
  SG_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int sgdml_b200_nystroem_expand(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double lam, const double* t,
                               const double* v, double* out, void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && t != nullptr && v != nullptr && out != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && lam > 0.0);
  SG_ARG(is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)n_rows, mm = (size_t)m;
  double* sc = nullptr;
  SG_TRY(scratch_get(mm + 2 * n, &sc));
  double *t_slot = sc, *v_slot = sc + mm, *out_slot = v_slot + n;
  const double *t_dev = nullptr, *v_dev = nullptr;
  SG_TRY(stage_in(t, mm, t_slot, s, &t_dev));
  SG_TRY(stage_in(v, n, v_slot, s, &v_dev));
  const bool out_host = !is_device_ptr(out);
  double* out_dev = out_host ? out_slot : out;
  k_x_t_minus_v<<<ceil_div(n_rows, 8), 256, 0, s>>>(X, n_rows, m, ldx, t_dev, v_dev, 1.0 / lam, out_dev);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  if (out_host) SG_CUDA(cudaMemcpyAsync(out, out_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  SG_CUDA(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"

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
  double s = 0.0;
  for (int64_t r = r0; r < r1; ++r) s = fma(X[r * ldx + j], v[r], s);
  part[(int64_t)blockIdx.y * m + j] = s;
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
  double s = 0.0;
  for (int64_t j = lane; j < m; j += 32) s = fma(X[r * ldx + j], t[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[r] = (s - v[r]) * lam_inv;
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

// t = X^T v (m doubles, device), deterministic two-pass reduction
static int xt_v_device(const double* X, int64_t n_rows, int64_t m, int64_t ldx, const double* v_dev, double* t_dev,
                       cudaStream_t s) {
  const int rows_per_cta = 256;
  const int64_t n_chunks = (n_rows + rows_per_cta - 1) / rows_per_cta;
  double* part = nullptr;
  SG_CUDA(cudaMalloc(&part, sizeof(double) * (size_t)m * n_chunks));
  auto body = [&]() -> int {
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)n_chunks);
    SG_ARG(grid.y <= 65535);
    k_xt_v<<<grid, 256, 0, s>>>(X, n_rows, m, ldx, v_dev, part, rows_per_cta);
    SG_CUDA(cudaGetLastError());
    k_reduce_chunks<<<ceil_div(m, 256), 256, 0, s>>>(part, n_chunks, m, t_dev);
    SG_CUDA(cudaGetLastError());
    count_launch(KID_MISC, 2);
    SG_CUDA(cudaStreamSynchronize(s));  // `part` is freed below
    return 0;
  };
  int rc = body();
  cudaFree(part);
  return rc;
}

int sgdml_b200_nystroem_apply(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double lam, const double* v,
                              double* out, void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && v != nullptr && out != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && lam > 0.0);
  SG_ARG(is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  Staged sV, sO;
  SG_TRY(sV.init(v, sizeof(double) * (size_t)n_rows, true, s));
  SG_TRY(sO.init(out, sizeof(double) * (size_t)n_rows, false, s));
  double* t = nullptr;
  SG_CUDA(cudaMalloc(&t, sizeof(double) * (size_t)m));
  auto body = [&]() -> int {
    SG_TRY(xt_v_device(X, n_rows, m, ldx, (const double*)sV.dev(), t, s));
    k_x_t_minus_v<<<ceil_div(n_rows, 8), 256, 0, s>>>(X, n_rows, m, ldx, t, (const double*)sV.dev(), 1.0 / lam,
                                                      (double*)sO.dev());
    SG_CUDA(cudaGetLastError());
    count_launch(KID_MISC);
    SG_TRY(sO.finish(s));
    SG_CUDA(cudaStreamSynchronize(s));
    return 0;
  };
  int rc = body();
  cudaFree(t);
  return rc;
}

int sgdml_b200_nystroem_project(const double* X, int64_t n_rows, int64_t m, int64_t ldx, const double* v, double* t,
                                void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && v != nullptr && t != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  Staged sV, sT;
  SG_TRY(sV.init(v, sizeof(double) * (size_t)n_rows, true, s));
  SG_TRY(sT.init(t, sizeof(double) * (size_t)m, false, s));
  SG_TRY(xt_v_device(X, n_rows, m, ldx, (const double*)sV.dev(), (double*)sT.dev(), s));
  SG_TRY(sT.finish(s));
  SG_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int sgdml_b200_nystroem_expand(const double* X, int64_t n_rows, int64_t m, int64_t ldx, double lam, const double* t,
                               const double* v, double* out, void* stream) {
  SG_TRY(require_device());
  SG_ARG(X != nullptr && t != nullptr && v != nullptr && out != nullptr && n_rows >= 1 && m >= 1 && ldx >= m && lam > 0.0);
  SG_ARG(is_device_ptr(X));
  cudaStream_t s = (cudaStream_t)stream;
  Staged sT, sV, sO;
  SG_TRY(sT.init(t, sizeof(double) * (size_t)m, true, s));
  SG_TRY(sV.init(v, sizeof(double) * (size_t)n_rows, true, s));
  SG_TRY(sO.init(out, sizeof(double) * (size_t)n_rows, false, s));
  k_x_t_minus_v<<<ceil_div(n_rows, 8), 256, 0, s>>>(X, n_rows, m, ldx, (const double*)sT.dev(), (const double*)sV.dev(),
                                                    1.0 / lam, (double*)sO.dev());
  SG_CUDA(cudaGetLastError());
  count_launch(KID_MISC);
  SG_TRY(sO.finish(s));
  SG_CUDA(cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"

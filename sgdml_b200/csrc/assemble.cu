// Path (a), assembly: the symmetric Matern-5/2 Hessian-kernel matrix K (SURVEY.md section 8
// rows a-K / a-KT) -- reference sgdml/train.py:97-232 (_assemble_kernel_mat_wkr),
// train.py:1260-1535 (_assemble_kernel_mat), torchtools.py:110-392 (GDMLTorchAssemble).
//
// B200 design.  The reference materialises the dense Jacobians (D x 3N, six non-zeros per
// row) and runs three S-fold einsums plus a 3N x D x 3N product per block
// (train.py:209-226).  Here the Jacobian never exists.  With the antisymmetric pair vectors
//   G_m[a][g] = (r_a - r_g)/|r_a - r_g|^3      (J_m[d(a,g), atom a] = -G_m[a][g])
// and P the atom permutation that induces the descriptor permutation perm_p, block (i,j) is
//   K_ij[a][b] = sum_p  c1_p u_p[a] (x) v_p[b]  -  c2_p T_p[a][b]          (3x3 per atom pair)
//   delta_p[d] = x_i[d] - x_j[perm_p[d]],  n_p = sqrt5 |delta_p|,  e_p = exp(-n_p/sig)
//   c1_p = 25 e_p/(3 sig^4),   c2_p = 5 (sig^2 + sig n_p) e_p/(3 sig^4)       (train.py:179-220)
//   u_p[a] = -sum_g G_i[a][g] delta_p[d(a,g)]                                  (= J_i^T delta_p)
//   v_p[b] = -sum_g G_j[b][g] delta_p[d(P^-1 b, P^-1 g)]                        (= J_j^(p)T delta_p)
//   T_p[a][b] = sum_g G_i[a][g] (x) G_j[Pa][Pg]   if b == P a,
//             = -G_i[a][P^-1 b] (x) G_j[Pa][b]    otherwise                     (= J_i^T J_j^(p))
// i.e. ~33 N^2 FMAs per permutation instead of the reference's ~2 D 3N (3S + 3N) per block.
// One CTA owns row point i and a tile of TJ column points; the per-(j,p) vectors are staged in
// shared memory; each thread then owns 3x3 atom-pair sub-blocks and writes them once.
#include <algorithm>

#include "common.cuh"
#include "desc.cuh"

namespace sgdml {

struct AsmArgs {
  const double* R_desc;    // (M, D)
  const double* R_d_desc;  // (M, D, 3)
  const int* dperm;        // (S, D) descriptor perms
  const int* aperm;        // (S, N) atom perms P
  const int* apinv;        // (S, N) inverse atom perms
  const int* jpts;         // (nJ) training point of each block-column
  const int64_t* dest;     // (nJ, 3N) destination column in K or -1
  int N, D, M, S, nJ, TJ;
  double sig, scale;
  double* K;
  int64_t ldk;
};

__device__ __forceinline__ int pair_idx_any(int a, int b) { return a > b ? pair_index(a, b) : pair_index(b, a); }

// expands compressed g (D,3) into the antisymmetric table G (N,N,3), zero diagonal
__device__ void load_pair_table(const double* __restrict__ g, int N, double* __restrict__ G, int tid, int nt) {
  for (int idx = tid; idx < N * N; idx += nt) {
    const int a = idx / N, b = idx - a * N;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    if (a > b) {
      const int d = pair_index(a, b);
      v0 = g[d * 3 + 0];
      v1 = g[d * 3 + 1];
      v2 = g[d * 3 + 2];
    } else if (b > a) {
      const int d = pair_index(b, a);
      v0 = -g[d * 3 + 0];
      v1 = -g[d * 3 + 1];
      v2 = -g[d * 3 + 2];
    }
    G[idx * 3 + 0] = v0;
    G[idx * 3 + 1] = v1;
    G[idx * 3 + 2] = v2;
  }
}

__global__ void __launch_bounds__(256) k_assemble(const AsmArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int N = p.N, D = p.D, S = p.S, TJ = p.TJ;
  const int N3 = 3 * N, NN3 = N * N * 3;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;

  const int i = blockIdx.y;
  const int jt0 = blockIdx.x * TJ;
  const int tj = min(TJ, p.nJ - jt0);

  // shared layout
  double* Gi = sm;                       // NN3
  double* xi = Gi + NN3;                 // D
  double* Gj = xi + D;                   // TJ*NN3
  double* xj = Gj + TJ * NN3;            // TJ*D
  double* del = xj + TJ * D;             // TJ*S*D
  double* cc = del + TJ * S * D;         // TJ*S*2
  double* u = cc + TJ * S * 2;           // TJ*S*N3
  double* v = u + TJ * S * N3;           // TJ*S*N3
  double* Dg = v + TJ * S * N3;          // TJ*S*3*N3
  int* sP = reinterpret_cast<int*>(Dg + TJ * S * 3 * N3);  // S*N
  int* sPi = sP + S * N;                                     // S*N

  // ---- stage 0: pair tables and descriptors of i and of the tile's points
  load_pair_table(p.R_d_desc + (int64_t)i * D * 3, N, Gi, tid, nt);
  for (int d = tid; d < D; d += nt) xi[d] = p.R_desc[(int64_t)i * D + d];
  for (int t = 0; t < tj; ++t) {
    const int j = p.jpts[jt0 + t];
    load_pair_table(p.R_d_desc + (int64_t)j * D * 3, N, Gj + t * NN3, tid, nt);
    for (int d = tid; d < D; d += nt) xj[t * D + d] = p.R_desc[(int64_t)j * D + d];
  }
  for (int idx = tid; idx < S * N; idx += nt) {
    sP[idx] = p.aperm[idx];
    sPi[idx] = p.apinv[idx];
  }
  __syncthreads();

  // ---- stage A: delta_p, n_p -> c1, c2 ; one warp per (j, p)
  const double sig = p.sig;
  const double sig2 = sig * sig;
  const double inv_div = 1.0 / (3.0 * sig2 * sig2);  // 1/mat52_base_div (train.py:179)
  for (int jp = warp; jp < tj * S; jp += nw) {
    const int t = jp / S, pp = jp - t * S;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double dl = xi[d] - xj[t * D + p.dperm[pp * D + d]];  // train.py:199
      del[jp * D + d] = dl;
      s = fma(dl, dl, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const double nrm = sqrt(5.0) * sqrt(s);            // train.py:201
      const double base = exp(-nrm / sig) * inv_div * 5.0;  // train.py:202
      cc[jp * 2 + 0] = base * 5.0;                       // c1 (train.py:211)
      cc[jp * 2 + 1] = (sig2 + sig * nrm) * base;        // c2 (train.py:219)
    }
  }
  __syncthreads();

  // ---- stage B: u_p, v_p (3N each) and the diagonal sums Dg_p (9N) per (j, p)
  {
    const int per = N3 + N3 + 3 * N3;
    for (int idx = tid; idx < tj * S * per; idx += nt) {
      const int jp = idx / per;
      int r = idx - jp * per;
      const int t = jp / S, pp = jp - t * S;
      const double* dl = del + jp * D;
      const double* Gjt = Gj + t * NN3;
      if (r < N3) {  // u_p[a][c]
        const int a = r / 3, c = r - 3 * a;
        double s = 0.0;
        for (int g = 0; g < N; ++g)
          if (g != a) s = fma(Gi[(a * N + g) * 3 + c], dl[pair_idx_any(a, g)], s);
        u[jp * N3 + r] = -s;
      } else if (r < 2 * N3) {  // v_p[b][c]
        r -= N3;
        const int b = r / 3, c = r - 3 * b;
        const int pib = sPi[pp * N + b];
        double s = 0.0;
        for (int g = 0; g < N; ++g)
          if (g != b) s = fma(Gjt[(b * N + g) * 3 + c], dl[pair_idx_any(pib, sPi[pp * N + g])], s);
        v[jp * N3 + r] = -s;
      } else {  // Dg_p[a][c][c']
        r -= 2 * N3;
        const int a = r / 9, c = (r - 9 * a) / 3, c2 = r - 9 * a - 3 * c;
        const int pa = sP[pp * N + a];
        double s = 0.0;
        for (int g = 0; g < N; ++g)
          if (g != a) s = fma(Gi[(a * N + g) * 3 + c], Gjt[(pa * N + sP[pp * N + g]) * 3 + c2], s);
        Dg[jp * 3 * N3 + r] = s;
      }
    }
  }
  __syncthreads();

  // ---- stage C: 3x3 atom-pair sub-blocks
  for (int it = tid; it < tj * N * N; it += nt) {
    const int t = it / (N * N);
    const int ab = it - t * N * N;
    const int a = ab / N, b = ab - a * N;
    const double* Gjt = Gj + t * NN3;
    double acc[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2) acc[c][c2] = 0.0;
    for (int pp = 0; pp < S; ++pp) {
      const int jp = t * S + pp;
      const double c1 = cc[jp * 2 + 0], c2v = cc[jp * 2 + 1];
      const double* ua = u + jp * N3 + 3 * a;
      const double* vb = v + jp * N3 + 3 * b;
      const int pa = sP[pp * N + a];
      double tt[3][3];
      if (b == pa) {
        const double* dg = Dg + jp * 3 * N3 + 9 * a;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int c2 = 0; c2 < 3; ++c2) tt[c][c2] = dg[c * 3 + c2];
      } else {
        const int g = sPi[pp * N + b];
        const double* gi = Gi + (a * N + g) * 3;
        const double* gj = Gjt + (pa * N + b) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int c2 = 0; c2 < 3; ++c2) tt[c][c2] = -gi[c] * gj[c2];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double cu = c1 * ua[c];
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) acc[c][c2] += cu * vb[c2] - c2v * tt[c][c2];
      }
    }
    const int64_t* dst = p.dest + (int64_t)(jt0 + t) * N3 + 3 * b;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double* Krow = p.K + ((int64_t)i * N3 + 3 * a + c) * p.ldk;
#pragma unroll
      for (int c2 = 0; c2 < 3; ++c2) {
        const int64_t col = dst[c2];
        if (col >= 0) Krow[col] = p.scale * acc[c][c2];
      }
    }
  }
}

static size_t asm_smem_bytes(int N, int D, int S, int TJ) {
  const size_t N3 = 3 * (size_t)N, NN3 = (size_t)N * N * 3;
  size_t dbl = NN3 + D + (size_t)TJ * NN3 + (size_t)TJ * D + (size_t)TJ * S * D + (size_t)TJ * S * 2 +
               2 * (size_t)TJ * S * N3 + (size_t)TJ * S * 3 * N3;
  return dbl * 8 + 2 * (size_t)S * N * 4;
}

}  // namespace sgdml

using namespace sgdml;

// Recovers the atom permutation P that induces a descriptor permutation (dperm[d(a,b)] =
// d(P a, P b)); returns false if dperm is not induced by any atom permutation.
static bool atom_perm_from_desc_perm(const int* dperm, int N, int* P) {
  const int D = N * (N - 1) / 2;
  if (N == 2) {
    P[0] = 0;
    P[1] = 1;
    return dperm[0] == 0;
  }
  for (int a = 0; a < N; ++a) {
    // two pairs containing a
    int o1 = (a == 0) ? 1 : 0, o2 = -1;
    for (int o = 0; o < N; ++o)
      if (o != a && o != o1) {
        o2 = o;
        break;
      }
    auto d_of = [](int x, int y) { return x > y ? x * (x - 1) / 2 + y : y * (y - 1) / 2 + x; };
    int e1 = dperm[d_of(a, o1)], e2 = dperm[d_of(a, o2)];
    int a1, b1, a2, b2;
    pair_from_d(e1, a1, b1);
    pair_from_d(e2, a2, b2);
    int common = -1;
    if (a1 == a2 || a1 == b2) common = a1;
    if (b1 == a2 || b1 == b2) common = (common == -1) ? b1 : -2;
    if (common < 0) return false;
    P[a] = common;
  }
  // verify
  std::vector<char> seen((size_t)N, 0);
  for (int a = 0; a < N; ++a) {
    if (P[a] < 0 || P[a] >= N || seen[(size_t)P[a]]) return false;
    seen[(size_t)P[a]] = 1;
  }
  for (int a = 1; a < N; ++a)
    for (int b = 0; b < a; ++b) {
      int pa = P[a], pb = P[b];
      int e = pa > pb ? pa * (pa - 1) / 2 + pb : pb * (pb - 1) / 2 + pa;
      if (dperm[a * (a - 1) / 2 + b] != e) return false;
    }
  (void)D;
  return true;
}

extern "C" int sgdml_b200_assemble(const double* R_desc, const double* R_d_desc, const int64_t* tril_perms_lin,
                                   int64_t n_atoms, int64_t n_train, int64_t n_perms, double sig,
                                   const int64_t* col_idxs, int64_t n_cols, double scale, double* K, int64_t ldk,
                                   void* stream) {
  SG_TRY(require_device());
  SG_ARG(R_desc != nullptr && R_d_desc != nullptr && tril_perms_lin != nullptr && K != nullptr);
  SG_ARG(n_atoms >= 2 && n_train >= 1 && n_perms >= 1 && sig > 0);
  const int N = (int)n_atoms, M = (int)n_train, S = (int)n_perms;
  const int D = N * (N - 1) / 2, N3 = 3 * N;
  const int64_t n = (int64_t)M * N3;
  if (col_idxs == nullptr) SG_ARG(n_cols == n);
  SG_ARG(n_cols >= 1 && n_cols <= n && ldk >= n_cols);
  cudaStream_t s = (cudaStream_t)stream;

  // ---- integer tables (host)
  std::vector<int64_t> lin((size_t)S * D);
  if (is_device_ptr(tril_perms_lin))
    SG_CUDA(cudaMemcpy(lin.data(), tril_perms_lin, sizeof(int64_t) * lin.size(), cudaMemcpyDeviceToHost));
  else
    std::copy(tril_perms_lin, tril_perms_lin + lin.size(), lin.begin());
  std::vector<int> dperm((size_t)S * D), aperm((size_t)S * N), apinv((size_t)S * N);
  for (int pp = 0; pp < S; ++pp) {
    for (int d = 0; d < D; ++d) {
      const int64_t e = lin[(size_t)d * S + pp] - (int64_t)pp * D;
      SG_ARG(e >= 0 && e < D);
      dperm[(size_t)pp * D + d] = (int)e;
    }
    if (!atom_perm_from_desc_perm(&dperm[(size_t)pp * D], N, &aperm[(size_t)pp * N]))
      return fail_arg("tril_perms_lin is not induced by atom permutations (utils/desc.py:509-539)");
    for (int a = 0; a < N; ++a) apinv[(size_t)pp * N + aperm[(size_t)pp * N + a]] = a;
  }
  // block-columns and destination map (train.py:1357-1407)
  std::vector<int> jpts;
  std::vector<int64_t> dest;
  if (col_idxs == nullptr) {
    jpts.resize((size_t)M);
    dest.resize((size_t)M * N3);
    for (int j = 0; j < M; ++j) {
      jpts[(size_t)j] = j;
      for (int k = 0; k < N3; ++k) dest[(size_t)j * N3 + k] = (int64_t)j * N3 + k;
    }
  } else {
    std::vector<int64_t> cols((size_t)n_cols);
    if (is_device_ptr(col_idxs))
      SG_CUDA(cudaMemcpy(cols.data(), col_idxs, sizeof(int64_t) * cols.size(), cudaMemcpyDeviceToHost));
    else
      std::copy(col_idxs, col_idxs + n_cols, cols.begin());
    for (int64_t c = 0; c < n_cols; ++c) {
      SG_ARG(cols[(size_t)c] >= 0 && cols[(size_t)c] < n);
      if (c > 0 && cols[(size_t)c] <= cols[(size_t)c - 1])
        return fail_arg("col_idxs must be sorted ascending without duplicates (train.py:1341-1345)");
      const int j = (int)(cols[(size_t)c] / N3), k = (int)(cols[(size_t)c] % N3);
      if (jpts.empty() || jpts.back() != j) {
        jpts.push_back(j);
        dest.insert(dest.end(), (size_t)N3, (int64_t)-1);
      }
      dest[(jpts.size() - 1) * N3 + k] = c;
    }
  }
  const int nJ = (int)jpts.size();

  // ---- tile size from the shared-memory budget
  int TJ = 0;
  for (int t = 8; t >= 1; --t)
    if (asm_smem_bytes(N, D, S, t) <= 200 * 1024) {
      TJ = t;
      break;
    }
  if (TJ == 0) {
    set_last_error("sgdml_b200_assemble: (N, S) too large for the single-pass assembly kernel");
    return SGDML_B200_ERR_UNSUPPORTED;
  }
  TJ = std::min(TJ, nJ);
  const size_t smem = asm_smem_bytes(N, D, S, TJ);

  Staged sX, sG, sK;
  SG_TRY(sX.init(R_desc, sizeof(double) * (size_t)M * D, true, s));
  SG_TRY(sG.init(R_d_desc, sizeof(double) * (size_t)M * D * 3, true, s));
  const bool K_host = !is_device_ptr(K);
  SG_TRY(sK.init(K, sizeof(double) * (size_t)n * ldk, false, s));
  if (K_host && ldk != n_cols) SG_CUDA(cudaMemsetAsync(sK.dev(), 0, sizeof(double) * (size_t)n * ldk, s));

  int *d_dperm = nullptr, *d_aperm = nullptr, *d_apinv = nullptr, *d_jpts = nullptr;
  int64_t* d_dest = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_dperm);
    cudaFree(d_aperm);
    cudaFree(d_apinv);
    cudaFree(d_jpts);
    cudaFree(d_dest);
  };
  auto body = [&]() -> int {
    SG_CUDA(cudaMalloc(&d_dperm, sizeof(int) * dperm.size()));
    SG_CUDA(cudaMalloc(&d_aperm, sizeof(int) * aperm.size()));
    SG_CUDA(cudaMalloc(&d_apinv, sizeof(int) * apinv.size()));
    SG_CUDA(cudaMalloc(&d_jpts, sizeof(int) * jpts.size()));
    SG_CUDA(cudaMalloc(&d_dest, sizeof(int64_t) * dest.size()));
    SG_CUDA(cudaMemcpyAsync(d_dperm, dperm.data(), sizeof(int) * dperm.size(), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaMemcpyAsync(d_aperm, aperm.data(), sizeof(int) * aperm.size(), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaMemcpyAsync(d_apinv, apinv.data(), sizeof(int) * apinv.size(), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaMemcpyAsync(d_jpts, jpts.data(), sizeof(int) * jpts.size(), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaMemcpyAsync(d_dest, dest.data(), sizeof(int64_t) * dest.size(), cudaMemcpyHostToDevice, s));
    SG_CUDA(cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AsmArgs a;
    a.R_desc = (const double*)sX.dev();
    a.R_d_desc = (const double*)sG.dev();
    a.dperm = d_dperm;
    a.aperm = d_aperm;
    a.apinv = d_apinv;
    a.jpts = d_jpts;
    a.dest = d_dest;
    a.N = N;
    a.D = D;
    a.M = M;
    a.S = S;
    a.nJ = nJ;
    a.TJ = TJ;
    a.sig = sig;
    a.scale = scale;
    a.K = (double*)sK.dev();
    a.ldk = ldk;
    // rows on grid.y (<= 65535), column tiles on grid.x
    SG_ARG(M <= 65535);
    dim3 grid((unsigned)ceil_div(nJ, TJ), (unsigned)M);
    {
      ProfScope ps(KID_ASSEMBLE, s);
      k_assemble<<<grid, 256, smem, s>>>(a);
      SG_CUDA(cudaGetLastError());
      count_launch(KID_ASSEMBLE);
    }
    SG_TRY(sK.finish(s));
    // the integer tables are read by the kernel: wait before freeing them
    SG_CUDA(cudaStreamSynchronize(s));
    return 0;
  };
  int rc = body();
  cleanup();
  return rc;
}

// FP64 GEMM on the 5th-generation tensor cores: C += alpha * A * B^T with A and B cut into signed
// 7-bit slices (int8) and every slice-pair product computed EXACTLY by tcgen05.mma kind::i8 with
// int32 accumulators in tensor memory (Ozaki-style error-free splitting).  This is the tcgen05 form
// of the Cholesky trailing update of path (a) (reference: scipy cho_factor = LAPACK dpotrf,
// sgdml/solvers/analytic.py:94-96): B200 has no f64 tensor-core kind, and the FP64 DMMA pipe tops
// out at 37 TFLOP/s, while 28 exact int8 GEMMs (7 slices) at 4.5 POP/s are worth ~160 TFLOP/s.
//
//   x_ij = 2^(e_i) * sum_{p=1..S} q_ij^(p) 2^(-7p),   |q^(p)| <= 64,   e_i per ROW
//   (A B^T)_ij = 2^(ea_i + eb_j) * sum_{L=2..S+1} 2^(-7L) * sum_{p+q=L} (A^(p) B^(q)T)_ij
// The inner sums are integers below 2^31 (64^2 * k * S with k <= 2^14), so the int32 tensor-core
// accumulation is exact; pairs with p + q > S + 1 are dropped (below the last kept bit).
// tools/ozaki_study.py (CPU, exact) shows what that buys on the sGDML system: with S = 7 the trained
// forces agree with the FP64 factorisation to 2e-11 at cond(K) = 2e10; S = 8 is FP64-equivalent.
//
// Kernel structure (one CTA per 128 x 64 tile of C, 192 threads, warp-specialised):
//   warp 0   TMA producer: 3-D tensor maps over the slice planes [S][rows][k] (int8, K-major, swizzled);
//            a pipeline UNIT is one slice p of one BK-wide k-block (BK = 64: A^(p) 128x64 B = 8 KB +
//            B^(p) 64x64 B = 4 KB); units travel through a ring of up to 18 slots (216 KB, more than two
//            k-blocks of all 7 slices) with full/empty mbarriers
//   warp 1   MMA issuer (one elected lane): per k-block the pairs are issued in groups r = 1, 2, ...
//            (all pairs with min(p, q) = r), after which slices r and S + 1 - r are dead and their
//            slots are handed back with tcgen05.commit -- the same order the producer refills them in
//   warps 2-5 epilogue: the S level accumulators (S x 64 TMEM columns) are read with tcgen05.ld, summed
//            smallest level first in FP64 registers, scaled by 2^(ea_i + eb_j) and added to C
//
// STATUS: written against the PTX ISA / CuTe descriptor definitions without access to a GPU (the
// round-1 GPU budget was spent); it compiles for sm_100a (SASS: UTCIMMA, UTMALDG.3D, LDTM) but has
// NOT been run.  It is therefore not wired into potrf; tests/test_ozaki.py is the bring-up harness.
#include <cuda.h>

#include "common.cuh"
#include "solve.cuh"

namespace sgdml {

constexpr int OZ_BITS = 7;
constexpr int OZ_MAX_S = 7;
constexpr int OZ_BM = 128, OZ_BN = 64;
constexpr int OZ_KPAD = 128;                               // the contraction length is padded to a multiple of this
// BK = bytes (= int8 elements) of k per pipeline unit = width of one swizzle row.  64 (64-byte swizzle) is the
// default: a unit is 12 KB, the ring holds 18 of them = two and a half k-blocks of all 7 slices, so the
// loads of the next k-block never wait for the current one to retire.  128 (128-byte swizzle, the layout every
// library GEMM uses) halves the ring depth to 9 units and is kept selectable for bring-up.
constexpr int OZ_RING_BYTES = 216 * 1024;
constexpr int OZ_MAX_RING = 18;
constexpr int OZ_UMMA_K = 32;                              // k per tcgen05.mma for 8-bit operands
constexpr int OZ_TMEM_COLS = 512;                          // S * 64 <= 448, allocation must be a power of two
template <int BK>
struct OzCfg {
  static_assert(BK == 64 || BK == 128, "unit width = swizzle span");
  static constexpr int A_BYTES = OZ_BM * BK;
  static constexpr int B_BYTES = OZ_BN * BK;
  static constexpr int UNIT_BYTES = A_BYTES + B_BYTES;      // 12 KB / 24 KB; both parts 1024-byte aligned
  static constexpr int MAX_SLOTS = OZ_RING_BYTES / UNIT_BYTES;  // 18 / 9
};
__host__ __device__ inline int oz_ring_slots(int S, int max_slots) { return (2 * S + 4 < max_slots) ? 2 * S + 4 : max_slots; }

// ---------------------------------------------------------------- splitting kernel
// One warp per row: exponent from the row maximum, then S rounds of (scale by 2^7, round to nearest,
// subtract).  planes: [S][rows_pad][kp] int8, zero padded; exps: [rows_pad].
__global__ void __launch_bounds__(256) k_ozaki_split(const double* __restrict__ X, int64_t rows, int64_t k, int64_t ldx,
                                                    int S, int64_t rows_pad, int64_t kp, int8_t* __restrict__ planes,
                                                    int* __restrict__ exps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows_pad) return;
  if (r >= rows) {  // padding rows: zeros
    for (int p = 0; p < S; ++p)
      for (int64_t j = lane; j < kp; j += 32) planes[((int64_t)p * rows_pad + r) * kp + j] = 0;
    if (lane == 0) exps[r] = 0;
    return;
  }
  const double* x = X + r * ldx;
  double amax = 0.0;
  for (int64_t j = lane; j < k; j += 32) amax = fmax(amax, fabs(x[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  int e = 0;
  if (amax > 0.0) {
    frexp(amax, &e);  // amax = f 2^e, f in [0.5, 1)
    e += 1;           // |x| 2^-e < 1/2
  }
  if (lane == 0) exps[r] = e;
  for (int64_t j = lane; j < kp; j += 32) {
    double v = (j < k) ? ldexp(x[j], -e) : 0.0;
    for (int p = 0; p < S; ++p) {
      v *= (double)(1 << OZ_BITS);
      const double q = rint(v);  // |q| <= 64, remainder in [-1/2, 1/2]
      planes[((int64_t)p * rows_pad + r) * kp + j] = (int8_t)(int)q;
      v -= q;
    }
  }
}

// ---------------------------------------------------------------- tcgen05 / TMEM helpers
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives (count 1) on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, single CTA
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Shared-memory matrix descriptor, K-major operand in the canonical 128-byte-swizzle layout TMA writes
// (cute::UMMA::SmemDescriptor, cute/arch/mma_sm100_desc.hpp): rows are 128 bytes apart, an 8-row swizzle
// atom is 1024 bytes.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (1 for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 (Blackwell) |
//   [49,52) base offset = 0 (tiles are 1024-byte aligned) | [61,64) layout type = 2 (SWIZZLE_128B)
// (64-byte swizzle: rows 64 bytes apart, atom 512 bytes, layout type 4)
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * BK) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(BK == 128 ? 2 : 4) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): [4,6) D format = 2 (S32) | [7,10) A format = 1
// (signed 8-bit) | [10,13) B format = 1 | [15] A major = 0 (K) | [16] B major = 0 (K) | [17,23) N >> 3 |
// [24,29) M >> 4; dense, no saturation, no negation.
__host__ __device__ constexpr uint32_t umma_idesc_s8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives row (lane quadrant base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct OzArgs {
  int64_t m, n;      // C is m x n (rows of A, rows of B)
  int64_t kp;        // padded contraction length (multiple of OZ_KPAD)
  int S;             // slices per operand
  int tri;           // 1: only tiles that touch the lower triangle (m == n, A and B the same row set)
  double alpha;      // +1 or -1 (any finite value works)
  const int* ea;     // row exponents of A (m_pad)
  const int* eb;     // row exponents of B (n_pad)
  double* C;
  int64_t ldc;
  int* dbg_levels;   // bring-up: raw int32 level sums, [S][m][n] (NULL in production)
};

struct OzSmemTail {
  uint64_t full[OZ_MAX_RING];
  uint64_t empty[OZ_MAX_RING];
  uint64_t acc_full;
  uint32_t tmem_base;
  double col_scale[OZ_BN];
};

constexpr size_t OZ_SMEM_BYTES = (size_t)OZ_RING_BYTES + sizeof(OzSmemTail) + 1024;

// slice visited at position idx of a k-block: 1, S, 2, S-1, ...  (1-based slice numbers)
__device__ __forceinline__ int oz_order(int idx, int S) { return (idx & 1) ? S - (idx >> 1) : 1 + (idx >> 1); }
// position of slice p in that order
__device__ __forceinline__ int oz_pos(int p, int S) { return (2 * p <= S + 1) ? 2 * (p - 1) : 2 * (S - p) + 1; }

template <int BK>
__global__ void __launch_bounds__(192, 1) k_ozaki_gemm(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const OzArgs p) {
  using Cfg = OzCfg<BK>;
  constexpr int OZ_A_BYTES = Cfg::A_BYTES, OZ_UNIT_BYTES = Cfg::UNIT_BYTES, OZ_BK = BK;
  extern __shared__ unsigned char oz_raw[];
  // 1024-byte alignment for the swizzled tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(oz_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S, R = oz_ring_slots(S, Cfg::MAX_SLOTS);
  OzSmemTail* tail = reinterpret_cast<OzSmemTail*>(smem + (size_t)OZ_RING_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int64_t m0 = (int64_t)blockIdx.y * OZ_BM, n0 = (int64_t)blockIdx.x * OZ_BN;
  if (m0 >= p.m || n0 >= p.n) return;
  if (p.tri && n0 > m0 + OZ_BM - 1) return;  // the tile lies entirely above the diagonal

  if (tid == 0) {
    for (int i = 0; i < R; ++i) {
      mbar_init(&tail->full[i], 1);
      mbar_init(&tail->empty[i], 1);
    }
    mbar_init(&tail->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {  // one warp allocates the tensor memory (and frees it at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)),
                 "n"(OZ_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid >= 64 && tid < 64 + OZ_BN) {  // column scales 2^(eb_j) of this tile
    const int j = tid - 64;
    tail->col_scale[j] = (n0 + j < p.n) ? ldexp(1.0, p.eb[n0 + j]) : 0.0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;
  const int KB = (int)(p.kp / OZ_BK);

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int64_t u = 0;
      for (int kb = 0; kb < KB; ++kb) {
        for (int idx = 0; idx < S; ++idx, ++u) {
          const int slot = (int)(u % R);
          const uint32_t round = (uint32_t)(u / R);
          if (round > 0) mbar_wait(&tail->empty[slot], (round - 1) & 1);  // the slot's previous unit is dead
          unsigned char* base = smem + (size_t)slot * OZ_UNIT_BYTES;
          const int sl = oz_order(idx, S) - 1;
          mbar_arrive_expect_tx(&tail->full[slot], (uint32_t)OZ_UNIT_BYTES);
          tma_load_3d(base, &tmA, kb * OZ_BK, (int)m0, sl, &tail->full[slot]);
          tma_load_3d(base + OZ_A_BYTES, &tmB, kb * OZ_BK, (int)n0, sl, &tail->full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc_s8(OZ_BM, OZ_BN);
      uint32_t level_started = 0;  // bit L: accumulator of level L holds data
      for (int kb = 0; kb < KB; ++kb) {
        const int64_t ub = (int64_t)kb * S;
        // group r = 1 needs every slice of this k-block
        for (int idx = 0; idx < S; ++idx) {
          const int64_t u = ub + idx;
          mbar_wait(&tail->full[u % R], (uint32_t)(u / R) & 1);
        }
        tc_fence_after();
        for (int r = 1; 2 * r <= S + 1; ++r) {
          // pairs with min(pa, pb) = r and pa + pb <= S + 1
          for (int t = r; t <= S + 1 - r; ++t) {
            for (int side = 0; side < 2; ++side) {
              if (side == 1 && t == r) continue;  // (r, r) only once
              const int pa = side == 0 ? r : t, pb = side == 0 ? t : r;
              const int level = pa + pb;  // 2 .. S + 1
              const int64_t ua = ub + oz_pos(pa, S), ubb = ub + oz_pos(pb, S);
              const uint32_t a_addr = smem_u32(smem + (size_t)(ua % R) * OZ_UNIT_BYTES);
              const uint32_t b_addr = smem_u32(smem + (size_t)(ubb % R) * OZ_UNIT_BYTES + OZ_A_BYTES);
              const uint32_t d_addr = tmem + (uint32_t)(level - 2) * OZ_BN;
#pragma unroll
              for (int ks = 0; ks < OZ_BK / OZ_UMMA_K; ++ks) {
                // advancing along K inside the swizzle row: +32 bytes on the start address
                const uint64_t da = umma_desc_kmajor<BK>(a_addr + ks * OZ_UMMA_K);
                const uint64_t db = umma_desc_kmajor<BK>(b_addr + ks * OZ_UMMA_K);
                tc_mma_i8(d_addr, da, db, IDESC, (level_started >> level) & 1u);
                level_started |= 1u << level;
              }
            }
          }
          // slices r and S + 1 - r are dead for this k-block: hand their slots back
          tc_commit(&tail->empty[(ub + oz_pos(r, S)) % R]);
          if (S + 1 - r != r) tc_commit(&tail->empty[(ub + oz_pos(S + 1 - r, S)) % R]);
        }
      }
      tc_commit(&tail->acc_full);  // every accumulator is final
    }
  } else {
    // ===================================================== epilogue (warps 2..5 = 128 threads)
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;        // row of the tile owned by this thread
    const int64_t gr = m0 + row;
    mbar_wait(&tail->acc_full, 0);
    tc_fence_after();
    const double row_scale = (gr < p.m) ? p.alpha * ldexp(1.0, p.ea[gr]) : 0.0;
#pragma unroll 1
    for (int half = 0; half < OZ_BN / 32; ++half) {
      double acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.0;
      for (int level = S + 1; level >= 2; --level) {  // smallest contributions first
        int v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((level - 2) * OZ_BN + half * 32), v);
        const double w = ldexp(1.0, -OZ_BITS * level);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma((double)v[j], w, acc[j]);
        if (p.dbg_levels != nullptr && gr < p.m) {
          for (int j = 0; j < 32; ++j) {
            const int64_t gc = n0 + half * 32 + j;
            if (gc < p.n) p.dbg_levels[((int64_t)(level - 2) * p.m + gr) * p.n + gc] = v[j];
          }
        }
      }
      if (gr < p.m) {
        double* crow = p.C + gr * p.ldc + n0 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int64_t gc = n0 + half * 32 + j;
          if (gc < p.n) crow[j] += acc[j] * row_scale * tail->col_scale[half * 32 + j];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(OZ_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled oz_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

// planes [S][rows_pad][kp] int8 -> 3-D map, box = 128 bytes of k x box_rows rows x 1 slice, 128-byte swizzle
static int oz_make_map(CUtensorMap* tm, const int8_t* planes, int S, int64_t rows_pad, int64_t kp, int box_rows, int bk) {
  PFN_tmapEncodeTiled enc = oz_tmap_encoder();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from this driver");
    return SGDML_B200_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {(cuuint64_t)kp, (cuuint64_t)rows_pad, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)kp, (cuuint64_t)(rows_pad * kp)};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (int8 planes) failed with CUresult %d", (int)r);
    set_last_error(buf);
    return SGDML_B200_ERR_ARG;
  }
  return 0;
}

struct OzOperand {
  int8_t* planes = nullptr;
  int* exps = nullptr;
  int64_t rows_pad = 0, kp = 0;
};

static size_t oz_plane_bytes(int64_t rows, int64_t k, int S) {
  const int64_t rows_pad = (rows + OZ_BM - 1) / OZ_BM * OZ_BM, kp = (k + OZ_KPAD - 1) / OZ_KPAD * OZ_KPAD;
  return (size_t)S * rows_pad * kp;
}

// slices X (rows x k) into caller-provided device memory
static int oz_split_into(const double* X, int64_t rows, int64_t k, int64_t ldx, int S, int8_t* planes, int* exps,
                         OzOperand* o, cudaStream_t s) {
  o->rows_pad = (rows + OZ_BM - 1) / OZ_BM * OZ_BM;
  o->kp = (k + OZ_KPAD - 1) / OZ_KPAD * OZ_KPAD;
  o->planes = planes;
  o->exps = exps;
  k_ozaki_split<<<ceil_div(o->rows_pad, 8), 256, 0, s>>>(X, rows, k, ldx, S, o->rows_pad, o->kp, o->planes, o->exps);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_GEMM);
  return 0;
}

static int oz_launch(const OzOperand& oa, const OzOperand& ob, int64_t m, int64_t n, double alpha, double* C,
                     int64_t ldc, int S, int tri, cudaStream_t s, int* dbg_levels = nullptr) {
  // unit width: 64 bytes (default, deep ring) or 128 bytes (SGDML_B200_OZAKI_BK=128, the plain 128-byte swizzle)
  const char* bke = getenv("SGDML_B200_OZAKI_BK");
  const int bk = (bke != nullptr && atoi(bke) == 128) ? 128 : 64;
  CUtensorMap tmA, tmB;
  SG_TRY(oz_make_map(&tmA, oa.planes, S, oa.rows_pad, oa.kp, OZ_BM, bk));
  SG_TRY(oz_make_map(&tmB, ob.planes, S, ob.rows_pad, ob.kp, OZ_BN, bk));
  static bool configured[64] = {false};
  int dev = 0;
  SG_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    SG_CUDA(cudaFuncSetAttribute(k_ozaki_gemm<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM_BYTES));
    SG_CUDA(cudaFuncSetAttribute(k_ozaki_gemm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM_BYTES));
    configured[dev] = true;
  }
  OzArgs a;
  a.m = m;
  a.n = n;
  a.kp = oa.kp;
  a.S = S;
  a.tri = tri;
  a.alpha = alpha;
  a.ea = oa.exps;
  a.eb = ob.exps;
  a.C = C;
  a.ldc = ldc;
  a.dbg_levels = dbg_levels;
  dim3 grid((unsigned)ceil_div(n, OZ_BN), (unsigned)ceil_div(m, OZ_BM));
  SG_ARG(grid.y <= 65535);
  ProfScope ps(KID_GEMM, s);
  if (bk == 128)
    k_ozaki_gemm<128><<<grid, 192, OZ_SMEM_BYTES, s>>>(tmA, tmB, a);
  else
    k_ozaki_gemm<64><<<grid, 192, OZ_SMEM_BYTES, s>>>(tmA, tmB, a);
  SG_CUDA(cudaGetLastError());
  count_launch(KID_GEMM);
  return 0;
}

// Workspace of the symmetric update used by potrf: allocated once per factorisation (a cudaMalloc /
// cudaFree pair costs ~20 ms in a process that holds tens of GB -- see csrc/core.cu), reused by every
// outer step, no host synchronisation in between.
int ozaki_syrk_workspace_bytes(int64_t max_rows, int64_t max_k, int S, size_t* plane_bytes, size_t* exp_bytes) {
  *plane_bytes = oz_plane_bytes(max_rows, max_k, S);
  *exp_bytes = sizeof(int) * (size_t)((max_rows + OZ_BM - 1) / OZ_BM * OZ_BM);
  return 0;
}

// C (n x n, lower-triangle tiles) += alpha X X^T with X (n x k): stream-ordered, no allocation
int ozaki_syrk_device(int64_t n, int64_t k, double alpha, const double* X, int64_t ldx, double* C, int64_t ldc, int S,
                      int8_t* planes, int* exps, cudaStream_t s) {
  SG_ARG(S >= 2 && S <= OZ_MAX_S && n >= 1 && k >= 1 && k <= (1 << 14));
  OzOperand o;
  SG_TRY(oz_split_into(X, n, k, ldx, S, planes, exps, &o, s));
  return oz_launch(o, o, n, n, alpha, C, ldc, S, 1, s);
}

// C (m x n, ldc) += alpha * A (m x k, lda) * B (n x k, ldb)^T through S int8 slices per operand.
// tri != 0: m == n and only tiles touching the lower triangle are updated.  All pointers on the device.
// Self-contained form (allocates and frees its slice planes, synchronises the stream): tests and the
// predictor experiment; the Cholesky uses ozaki_syrk_device.
int ozaki_gemm_nt_device(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double* C, int64_t ldc, int S, int tri, cudaStream_t s) {
  SG_ARG(S >= 2 && S <= OZ_MAX_S && m >= 1 && n >= 1 && k >= 1);
  SG_ARG(k <= (1 << 14));  // int32 accumulation stays exact: 64^2 * k * S < 2^31
  const bool same = (A == B && m == n && lda == ldb);
  int8_t *pa = nullptr, *pb = nullptr;
  int *xa = nullptr, *xb = nullptr;
  auto cleanup = [&]() {
    cudaFree(pa);
    cudaFree(xa);
    cudaFree(pb);
    cudaFree(xb);
  };
  auto body = [&]() -> int {
    OzOperand oa, ob;
    SG_CUDA(cudaMalloc(&pa, oz_plane_bytes(m, k, S)));
    SG_CUDA(cudaMalloc(&xa, sizeof(int) * (size_t)((m + OZ_BM - 1) / OZ_BM * OZ_BM)));
    SG_TRY(oz_split_into(A, m, k, lda, S, pa, xa, &oa, s));
    if (same) {
      ob = oa;
    } else {
      SG_CUDA(cudaMalloc(&pb, oz_plane_bytes(n, k, S)));
      SG_CUDA(cudaMalloc(&xb, sizeof(int) * (size_t)((n + OZ_BM - 1) / OZ_BM * OZ_BM)));
      SG_TRY(oz_split_into(B, n, k, ldb, S, pb, xb, &ob, s));
    }
    SG_TRY(oz_launch(oa, ob, m, n, alpha, C, ldc, S, tri, s));
    SG_CUDA(cudaStreamSynchronize(s));  // the planes are freed below
    return 0;
  };
  int rc = body();
  cleanup();
  return rc;
}

// Bring-up aid: runs the split and the int8 products and hands back every intermediate.
//   planes_a [S][m_pad][kp] int8, exps_a [m_pad], planes_b / exps_b likewise, levels [S][m][n] int32
// (all device pointers; any of them may be NULL).  C receives C + A B^T as usual.
int ozaki_debug_device(int64_t m, int64_t n, int64_t k, const double* A, int64_t lda, const double* B, int64_t ldb,
                       double* C, int64_t ldc, int S, int8_t* planes_a, int* exps_a, int8_t* planes_b, int* exps_b,
                       int* levels, cudaStream_t s) {
  SG_ARG(S >= 2 && S <= OZ_MAX_S && m >= 1 && n >= 1 && k >= 1 && k <= (1 << 14));
  int8_t *pa = nullptr, *pb = nullptr;
  int *xa = nullptr, *xb = nullptr;
  auto cleanup = [&]() {
    cudaFree(pa);
    cudaFree(xa);
    cudaFree(pb);
    cudaFree(xb);
  };
  auto body = [&]() -> int {
    OzOperand oa, ob;
    const size_t ba = oz_plane_bytes(m, k, S), bb = oz_plane_bytes(n, k, S);
    const size_t ea = sizeof(int) * (size_t)((m + OZ_BM - 1) / OZ_BM * OZ_BM), eb = sizeof(int) * (size_t)((n + OZ_BM - 1) / OZ_BM * OZ_BM);
    SG_CUDA(cudaMalloc(&pa, ba));
    SG_CUDA(cudaMalloc(&xa, ea));
    SG_CUDA(cudaMalloc(&pb, bb));
    SG_CUDA(cudaMalloc(&xb, eb));
    SG_TRY(oz_split_into(A, m, k, lda, S, pa, xa, &oa, s));
    SG_TRY(oz_split_into(B, n, k, ldb, S, pb, xb, &ob, s));
    if (planes_a) SG_CUDA(cudaMemcpyAsync(planes_a, pa, ba, cudaMemcpyDeviceToDevice, s));
    if (exps_a) SG_CUDA(cudaMemcpyAsync(exps_a, xa, ea, cudaMemcpyDeviceToDevice, s));
    if (planes_b) SG_CUDA(cudaMemcpyAsync(planes_b, pb, bb, cudaMemcpyDeviceToDevice, s));
    if (exps_b) SG_CUDA(cudaMemcpyAsync(exps_b, xb, eb, cudaMemcpyDeviceToDevice, s));
    if (C != nullptr) SG_TRY(oz_launch(oa, ob, m, n, 1.0, C, ldc, S, 0, s, levels));
    SG_CUDA(cudaStreamSynchronize(s));
    return 0;
  };
  int rc = body();
  cleanup();
  return rc;
}

}  // namespace sgdml

using namespace sgdml;

extern "C" int sgdml_b200_ozaki_debug(int64_t m, int64_t n, int64_t k, const double* A, int64_t lda, const double* B,
                                      int64_t ldb, double* C, int64_t ldc, int n_slices, int8_t* planes_a, int* exps_a,
                                      int8_t* planes_b, int* exps_b, int* levels, void* stream) {
  SG_TRY(require_device());
  SG_ARG(A != nullptr && B != nullptr && lda >= k && ldb >= k);
  return ozaki_debug_device(m, n, k, A, lda, B, ldb, C, ldc, n_slices, planes_a, exps_a, planes_b, exps_b, levels,
                            (cudaStream_t)stream);
}

extern "C" int sgdml_b200_ozaki_gemm_nt(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                                        const double* B, int64_t ldb, double* C, int64_t ldc, int n_slices, int tri,
                                        void* stream) {
  SG_TRY(require_device());
  SG_ARG(A != nullptr && B != nullptr && C != nullptr && lda >= k && ldb >= k && ldc >= n);
  SG_ARG(is_device_ptr(A) && is_device_ptr(B) && is_device_ptr(C));
  if (tri) SG_ARG(m == n);
  return ozaki_gemm_nt_device(m, n, k, alpha, A, lda, B, ldb, C, ldc, n_slices, tri, (cudaStream_t)stream);
}

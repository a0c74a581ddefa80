// Query x gallery Euclidean distance (reid_dataset_evaluator.py:244-272) on the 5th-gen
// tensor cores: the one dense contraction of the path, a . b^T, as TMA-fed tcgen05.mma tiles
// with fp32 accumulators in TMEM; the epilogue fuses |a|^2 + |b|^2 - 2ab, the clamp at 0 and
// the sqrt (reid_dataset_evaluator.py:269-271) and writes the [m1, m2] fp32 block.
//
// fp32-accurate path: both operands arrive as bf16 residual planes (split_prep.cu). The
// product is the sum of plane-pair terms (p0.p0 + p0.p1 + p1.p0 for BF16X3), all accumulated
// into the same TMEM tile, smallest terms first — i.e. one GEMM whose K loop walks
// (term, k-block) pairs.  Both operands are K-major, so A and B tiles are plain TMA boxes of
// 64 bf16 (128 B, SWIZZLE_128B) x {128, 256} rows and feed UMMA without any transpose.
//
// Kernel shape: persistent, one CTA per SM, 128 x 256 output tile, BK = 64.
//   warp 0   : TMA producer (4-stage ring, 48 KB per stage)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per stage)
//   warps 2-5: epilogue, 128 threads = 128 TMEM lanes (rows); double-buffered accumulators
//              (2 x 256 TMEM columns) so the epilogue of tile t overlaps the mainloop of t+1.
#include "common.cuh"

#include <mutex>

namespace pps {

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;                              // bf16 elements = 128 bytes = one swizzle row
constexpr int kGemmStages = 4;
constexpr int kABytes = kBM * kBK * 2;               // 16 KB
constexpr int kBBytes = kBN * kBK * 2;               // 32 KB
constexpr int kStageBytesG = kABytes + kBBytes;      // 48 KB
constexpr int kGemmThreads = 192;
constexpr int kMaxTerms = 6;
constexpr size_t kGemmSmem = 1024 /*align slack*/ + (size_t)kGemmStages * kStageBytesG + 256 /*barriers*/;

struct GemmArgs {
  long long m1, m2;
  int kblocks;                // ceil(K / 64)
  int nterms;
  int term_a[kMaxTerms];      // plane of A used by term t
  int term_b[kMaxTerms];
  uint32_t idesc;
  const float* a_sqnorm;
  const float* b_sqnorm;
  float* out;
  long long ldo;
  int flags;
  int m_tiles, n_tiles;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
dist_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmArgs g) {
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kGemmStages * kStageBytesG);
  uint64_t* empty = full + kGemmStages;
  uint64_t* tfull = empty + kGemmStages;     // [2] accumulator ready for the epilogue
  uint64_t* tempty = tfull + 2;              // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles = (long long)g.m_tiles * g.n_tiles;
  const int kiters = g.nterms * g.kblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * kBN);   // 512 columns: two 128x256 fp32 accumulators
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (int)(t % g.m_tiles) * kBM;     // m fastest: concurrent CTAs share B tiles in L2
        const int n0 = (int)(t / g.m_tiles) * kBN;
        for (int term = 0; term < g.nterms; ++term) {
          const int pa = g.term_a[term], pb = g.term_b[term];
          for (int kb = 0; kb < g.kblocks; ++kb, ++it) {
            const uint32_t s = it % kGemmStages, ph = (it / kGemmStages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_expect_tx(&full[s], kStageBytesG);
            unsigned char* sa = smem + (size_t)s * kStageBytesG;
            tma_load_3d(sa, &tmA, &full[s], kb * kBK, m0, pa);
            tma_load_3d(sa + kABytes, &tmB, &full[s], kb * kBK, n0, pb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++acc_it) {
        const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(&tempty[as], aph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kBN;
        for (int ki = 0; ki < kiters; ++ki, ++it) {
          const uint32_t s = it % kGemmStages, ph = (it / kGemmStages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytesG);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in the >>4 address field
            tc_mma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, g.idesc, (ki | k) ? 1u : 0u);
          }
          tc_commit(&empty[s]);                 // smem slot reusable once these MMAs retire
        }
        tc_commit(&tfull[as]);                  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int lane_grp = warp & 3;              // TMEM lanes 32*lane_grp .. +31 are this warp's
    const int row = lane_grp * 32 + lane;
    uint32_t acc_it = 0;
    const bool want_sq = (g.flags & PPS_DIST_SQUARED) != 0;
    const bool want_dot = (g.flags & PPS_DIST_DOT) != 0;
    const bool vec_ok = ((g.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.out) & 15u) == 0);
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++acc_it) {
      const int m0 = (int)(t % g.m_tiles) * kBM;
      const int n0 = (int)(t / g.m_tiles) * kBN;
      const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const long long gi = (long long)m0 + row;
      const bool row_ok = gi < g.m1;
      const float an = (row_ok && !want_dot) ? __ldg(g.a_sqnorm + gi) : 0.f;
      float* orow = g.out + (row_ok ? gi : 0) * g.ldo;
      const uint32_t tbase = tmem_base + as * kBN + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kBN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tbase + c, r);
        tmem_ld_wait();
        const long long gj0 = (long long)n0 + c;
        if (row_ok && gj0 < g.m2) {
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float dot = __uint_as_float(r[e]);
            if (want_dot) {
              v[e] = dot;
            } else {
              const long long gj = gj0 + e;
              const float bn = gj < g.m2 ? __ldg(g.b_sqnorm + gj) : 0.f;
              // same association as the reference: (-2*ab + |a|^2) + |b|^2
              float d2 = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), an), bn);
              d2 = d2 < 0.f ? 0.f : d2;
              v[e] = want_sq ? d2 : __fsqrt_rn(d2);
            }
          }
          if (vec_ok && gj0 + 32 <= g.m2) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(orow + gj0 + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (gj0 + e < g.m2) orow[gj0 + e] = v[e];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kBN);
  }
}

// ------------------------------------------------------------------------------------
// CUDA-core fp32 distance (PPS_PREC_FP32): exact-product FMA path on the original rows.
// 64x64 tile per CTA, 16x16 threads, 4x4 outputs per thread, K tile 16.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dist_fp32_kernel(const float* __restrict__ a, long long lda,
                                                        const float* __restrict__ an, long long m1,
                                                        const float* __restrict__ b, long long ldb,
                                                        const float* __restrict__ bn, long long m2, int dim, int flags,
                                                        float* __restrict__ out, long long ldo) {
  __shared__ float sa[16][64 + 4];
  __shared__ float sb[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = (long long)blockIdx.y * 64, n0 = (long long)blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < dim; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      sa[k][r] = (m0 + r < m1 && k0 + k < dim) ? a[(m0 + r) * lda + k0 + k] : 0.f;
      sb[k][r] = (n0 + r < m2 && k0 + k < dim) ? b[(n0 + r) * ldb + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; bv[i] = sb[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool want_sq = (flags & PPS_DIST_SQUARED) != 0, want_dot = (flags & PPS_DIST_DOT) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gi = m0 + ty * 4 + i;
    if (gi >= m1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long gj = n0 + tx * 4 + j;
      if (gj >= m2) continue;
      float v = acc[i][j];
      if (!want_dot) {
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, v), an[gi]), bn[gj]);
        d2 = d2 < 0.f ? 0.f : d2;
        v = want_sq ? d2 : __fsqrt_rn(d2);
      }
      out[gi * ldo + gj] = v;
    }
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// planes: [n_planes][rows][kpad] 16-bit elements, K-major
static int make_operand_map(CUtensorMap* tm, const void* planes, long long rows, long long plane_rows, int kpad,
                            int n_planes, int box_rows, bool f16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cuda_fail(cudaErrorUnknown, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)kpad, (cuuint64_t)rows, (cuuint64_t)n_planes};
  cuuint64_t strides[2] = {(cuuint64_t)kpad * 2, (cuuint64_t)plane_rows * (cuuint64_t)kpad * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed");
  return PPS_OK;
}

}  // namespace pps

using namespace pps;

extern "C" int pps_dist_tc(const void* a_planes, const float* a_sqnorm, long long m1, int a_planes_n,
                           long long a_plane_rows, const void* b_planes, const float* b_sqnorm, long long m2,
                           int b_planes_n, long long b_plane_rows, int dim, int precision, int flags, float* dist,
                           long long ldd, void* stream) {
  if (m1 < 0 || m2 < 0 || dim <= 0 || ldd < m2) return PPS_ERR_INVALID_ARG;
  if (a_plane_rows == 0) a_plane_rows = m1;
  if (b_plane_rows == 0) b_plane_rows = m2;
  if (a_plane_rows < m1 || b_plane_rows < m2) return PPS_ERR_INVALID_ARG;
  if (m1 == 0 || m2 == 0) return PPS_OK;
  if (!a_planes || !b_planes || !dist) return PPS_ERR_INVALID_ARG;
  if (!(flags & PPS_DIST_DOT) && (!a_sqnorm || !b_sqnorm)) return PPS_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(a_planes) & 15u) || (reinterpret_cast<uintptr_t>(b_planes) & 15u))
    return PPS_ERR_ALIGN;
  if (m1 > 0x7fffff00LL || m2 > 0x7fffff00LL) return PPS_ERR_UNSUPPORTED;   // TMA coordinates are int32

  GemmArgs g;
  int need = 1;
  bool f16 = false;
  switch (precision) {
    case PPS_PREC_BF16X1:
      g.nterms = 1; g.term_a[0] = 0; g.term_b[0] = 0; need = 1; break;
    case PPS_PREC_BF16X3: {
      const int ta[3] = {1, 0, 0}, tb[3] = {0, 1, 0};   // small terms first
      g.nterms = 3; for (int i = 0; i < 3; ++i) { g.term_a[i] = ta[i]; g.term_b[i] = tb[i]; }
      need = 2; break;
    }
    case PPS_PREC_BF16X6: {
      const int ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
      g.nterms = 6; for (int i = 0; i < 6; ++i) { g.term_a[i] = ta[i]; g.term_b[i] = tb[i]; }
      need = 3; break;
    }
    case PPS_PREC_F16X1:
      g.nterms = 1; g.term_a[0] = 0; g.term_b[0] = 0; need = 1; f16 = true; break;
    default: return PPS_ERR_INVALID_ARG;
  }
  if (a_planes_n < need || b_planes_n < need) return PPS_ERR_INVALID_ARG;

  const int kpad = pps_kpad(dim);
  g.m1 = m1; g.m2 = m2;
  g.kblocks = kpad / kBK;
  const uint32_t fmt = f16 ? 0u : 1u;   // F16F32Format: F16 = 0, BF16 = 1
  g.idesc = (1u << 4) /*D = f32*/ | (fmt << 7) | (fmt << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  g.a_sqnorm = a_sqnorm; g.b_sqnorm = b_sqnorm;
  g.out = dist; g.ldo = ldd; g.flags = flags;
  g.m_tiles = (int)((m1 + kBM - 1) / kBM);
  g.n_tiles = (int)((m2 + kBN - 1) / kBN);

  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, a_planes, m1, a_plane_rows, kpad, a_planes_n, kBM, f16);
  if (rc) return rc;
  rc = make_operand_map(&tmB, b_planes, m2, b_plane_rows, kpad, b_planes_n, kBN, f16);
  if (rc) return rc;

  static thread_local int configured_dev = -1;
  int dev = 0;
  PPS_CUDA_TRY(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    PPS_CUDA_TRY(cudaFuncSetAttribute(dist_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    configured_dev = dev;
  }
  const long long tiles = (long long)g.m_tiles * g.n_tiles;
  const int sms = sm_count();
  const int grid = (int)(tiles < sms ? tiles : sms);
  dist_tc_kernel<<<grid, kGemmThreads, kGemmSmem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, g);
  PPS_LAUNCH_CHECK("dist_tc_kernel");
  return PPS_OK;
}

extern "C" int pps_dist_fp32(const float* a, long long lda, const float* a_sqnorm, long long m1, const float* b,
                             long long ldb, const float* b_sqnorm, long long m2, int dim, int flags, float* dist,
                             long long ldd, void* stream) {
  if (m1 < 0 || m2 < 0 || dim <= 0 || lda < dim || ldb < dim || ldd < m2) return PPS_ERR_INVALID_ARG;
  if (m1 == 0 || m2 == 0) return PPS_OK;
  if (!a || !b || !dist) return PPS_ERR_INVALID_ARG;
  if (!(flags & PPS_DIST_DOT) && (!a_sqnorm || !b_sqnorm)) return PPS_ERR_INVALID_ARG;
  const long long gx = (m2 + 63) / 64, gy = (m1 + 63) / 64;
  if (gy > 65535 || gx > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
  dist_fp32_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, lda, a_sqnorm, m1, b, ldb, b_sqnorm, m2, dim, flags, dist, ldd);
  PPS_LAUNCH_CHECK("dist_fp32_kernel");
  return PPS_OK;
}

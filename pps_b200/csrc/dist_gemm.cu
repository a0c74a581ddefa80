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
// 2-CTA kernel (default): a cluster of two CTAs (one SM pair) owns a 256 x 256 output tile and
// issues tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA, N = 256).  Per 64-wide k-block each
// CTA loads ONLY its 128 rows of every A plane and its 128-row half of every B plane, and all
// plane-pair terms of the split product reuse those tiles: for bf16x3 that is 4 x 16 KB per CTA
// for 12 MMAs (5.3 KB of L2->SM traffic per 128-cycle MMA slot and SM, against 12 KB for the
// 1-CTA kernel above, which sits at the chip's TMA/L2 delivery limit rather than at the tensor
// pipe).  Barriers: full[] lives in the leader (rank 0) and collects both CTAs' TMA bytes;
// empty[] / tfull[] exist in both CTAs and are signalled by multicast tcgen05.commit; tempty[]
// lives in the leader and takes the arrivals of both CTAs' epilogue warps.
// ------------------------------------------------------------------------------------
constexpr int kT2Rows = 128;
constexpr int kTile2Bytes = kT2Rows * kBK * 2;      // 16 KB: one plane of one operand for one k-block
constexpr int kRing2Bytes = 192 * 1024;
constexpr int kMaxStages2 = 6;
constexpr int kOutChunkBytes = 32 * 32 * 4;                 // one warp's 32 x 32 fp32 staging tile
constexpr int kOutStageBytes = 4 * 2 * kOutChunkBytes;      // 4 epilogue warps, double-buffered
constexpr size_t kGemm2Smem = (size_t)kRing2Bytes + kOutStageBytes + 2 * 256 * 4 /*|b|^2 or alpha/beta*/ + 256 /*barriers*/;

struct Gemm2Args {
  GemmArgs g;
  int planes;       // planes of each operand loaded per k-block (1..3)
  int stages;       // ring depth = kRing2Bytes / stage bytes
  // grouped form (embedding head): `groups` independent products; group grp uses A rows [grp*a_group_rows, ...),
  // B rows [grp*b_group_rows, ...) and writes output columns [grp*out_group_cols, ...).  Distance: groups = 1.
  int groups;
  long long a_group_rows, b_group_rows, out_group_cols;
};

constexpr int EPI_DIST = 0;      // |a|^2 + |b|^2 - 2ab, clamp, sqrt (or squared / raw dot by flags)
constexpr int EPI_AFFINE_RELU = 1;   // max(0, dot * alpha[col] + beta[col])   (a_sqnorm = alpha, b_sqnorm = beta)

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
dist_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ Gemm2Args ga) {
  extern __shared__ __align__(1024) unsigned char smem[];   // SWIZZLE_128B tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* stage_out = smem + kRing2Bytes;                                   // [4 warps][2][4 KB]
  float* bn_s = reinterpret_cast<float*>(stage_out + kOutStageBytes);             // [2][256] (or [2][2][128])
  constexpr int kBTile = (BN / 2) * kBK * 2;                 // this CTA's half of a B plane tile
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(bn_s) + 2 * 256 * 4);
  uint64_t* empty = full + kMaxStages2;
  uint64_t* tfull = empty + kMaxStages2;      // [2]
  uint64_t* tempty = tfull + 2;               // [2] (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const GemmArgs& g = ga.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long tiles_per_group = (long long)g.m_tiles * g.n_tiles;   // 256 x BN tiles
  const long long tiles = tiles_per_group * ga.groups;
  const long long pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int planes = ga.planes, stages = ga.stages;
  const uint32_t stage_bytes = (uint32_t)planes * (kTile2Bytes + kBTile);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kMaxStages2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);   // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long t = pair; t < tiles; t += npairs) {
        const long long grp = t / tiles_per_group, tt = t % tiles_per_group;
        const int m0 = (int)(grp * ga.a_group_rows) + (int)(tt % g.m_tiles) * 256 + (int)rank * kT2Rows;
        const int n0 = (int)(grp * ga.b_group_rows) + (int)(tt / g.m_tiles) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < g.kblocks; ++kb, ++it) {
          const uint32_t s = it % stages, ph = (it / stages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          if (rank == 0) mbar_expect_tx(&full[s], 2u * stage_bytes);
          const uint32_t full0 = map_to_cta(smem_u32(&full[s]), 0);
          unsigned char* slot = smem + (size_t)s * stage_bytes;
          for (int p = 0; p < planes; ++p) tma_load_3d_2cta(slot + p * kTile2Bytes, &tmA, full0, kb * kBK, m0, p);
          for (int p = 0; p < planes; ++p)
            tma_load_3d_2cta(slot + planes * kTile2Bytes + p * kBTile, &tmB, full0, kb * kBK, n0, p);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (long long t = pair; t < tiles; t += npairs, ++acc_it) {
        const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        mbar_wait(&tempty[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < g.kblocks; ++kb, ++it) {
          const uint32_t s = it % stages, ph = (it / stages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t slot = smem_u32(smem + (size_t)s * stage_bytes);
          for (int term = 0; term < g.nterms; ++term) {
            const uint64_t adesc = umma_desc_k_sw128(slot + g.term_a[term] * kTile2Bytes);
            const uint64_t bdesc = umma_desc_k_sw128(slot + planes * kTile2Bytes + g.term_b[term] * kBTile);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc_mma_f16_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, g.idesc, (kb | term | k) ? 1u : 0u);
          }
          tc_commit_2cta(&empty[s], 3);          // both CTAs' slots reusable once these MMAs retire
        }
        tc_commit_2cta(&tfull[as], 3);           // accumulator complete in both CTAs
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    // TMEM -> registers (thread = row) -> |a|^2 + |b|^2 - 2ab, clamp, sqrt -> 128B-swizzled staging tile in
    // shared memory -> one TMA store per 32 x 32 chunk (coalesced, clipped at the matrix edges by the tensor map).
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;
    const int etid = (int)threadIdx.x - 64;     // 0..127 among the epilogue threads
    uint32_t acc_it = 0, chunk_it = 0;
    const bool want_sq = (g.flags & PPS_DIST_SQUARED) != 0;
    const bool want_dot = (g.flags & PPS_DIST_DOT) != 0;
    unsigned char* my_stage = stage_out + (size_t)lane_grp * (2 * kOutChunkBytes);
    const uint32_t swz = (uint32_t)(lane & 7);
    for (long long t = pair; t < tiles; t += npairs, ++acc_it) {
      const long long grp = t / tiles_per_group, tt = t % tiles_per_group;
      const int m0 = (int)(tt % g.m_tiles) * 256 + (int)rank * kT2Rows;             // output row (within the group's rows)
      const int n0 = (int)(grp * ga.out_group_cols) + (int)(tt / g.m_tiles) * BN;   // output column
      const int n_end = EPI == EPI_DIST ? (int)g.m2 : (int)((grp + 1) * ga.out_group_cols);   // first column not ours
      const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
      // per-column operands of the tile -> shared (double-buffered by accumulator stage):
      // |b|^2 of the gallery rows, or alpha / beta of the output channels
      float* bn = bn_s + as * 256;
      if (EPI == EPI_AFFINE_RELU) {
        const long long gj = (long long)n0 + etid;            // BN == 128 == number of epilogue threads
        bn[etid] = gj < n_end ? __ldg(g.a_sqnorm + gj) : 0.f;
        bn[etid + 128] = gj < n_end ? __ldg(g.b_sqnorm + gj) : 0.f;
      } else if (!want_dot) {
        const long long gj = (long long)n0 + etid;
        bn[etid] = gj < g.m2 ? __ldg(g.b_sqnorm + gj) : 0.f;
        if (BN > 128) bn[etid + 128] = gj + 128 < g.m2 ? __ldg(g.b_sqnorm + gj + 128) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long gi = (long long)m0 + row;
      const float an = (EPI == EPI_DIST && gi < g.m1 && !want_dot) ? __ldg(g.a_sqnorm + gi) : 0.f;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + as * BN + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32, ++chunk_it) {
        uint32_t r[32];
        tmem_ld_32x32(tbase + c, r);
        unsigned char* buf = my_stage + (chunk_it & 1u) * kOutChunkBytes;
        bulk_wait_group_read<1>();               // the store that last read this buffer has drained it
        __syncwarp();
        tmem_ld_wait();
        const float4* bn4 = reinterpret_cast<const float4*>(bn + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[4];
          const float4 b4 = bn4[j];
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          float cc[4] = {0.f, 0.f, 0.f, 0.f};
          if (EPI == EPI_AFFINE_RELU) {
            const float4 c4 = reinterpret_cast<const float4*>(bn + 128 + c)[j];
            cc[0] = c4.x; cc[1] = c4.y; cc[2] = c4.z; cc[3] = c4.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float dot = __uint_as_float(r[4 * j + e]);
            if (EPI == EPI_AFFINE_RELU) {
              v[e] = fmaxf(fmaf(dot, bb[e], cc[e]), 0.f);
            } else if (want_dot) {
              v[e] = dot;
            } else {
              // same association as the reference: (-2*ab + |a|^2) + |b|^2   (-2*ab is exact, so the FMA rounds once)
              float d2 = __fadd_rn(__fmaf_rn(-2.f, dot, an), bb[e]);
              d2 = fmaxf(d2, 0.f);
              v[e] = want_sq ? d2 : sqrt_approx(d2);
            }
          }
          *reinterpret_cast<float4*>(buf + lane * 128 + (((uint32_t)j ^ swz) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (m0 + lane_grp * 32 < g.m1 && n0 + c < n_end) tma_store_2d(&tmO, buf, n0 + c, m0 + lane_grp * 32);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty[as]);
        else mbar_arrive_cluster(map_to_cta(smem_u32(&tempty[as]), 0));
      }
    }
    bulk_wait_group_read<0>();
  }

  tc_fence_before();
  cluster_sync_all();      // no CTA of the pair leaves (or frees TMEM) while its peer still signals it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
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

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0;
  PPS_CUDA_TRY(cudaGetDevice(&dev));
  const int sms = sm_count();
  CUtensorMap tmA, tmB;

  const bool out_tma_ok = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(dist) & 15u) == 0);
  if (!(flags & PPS_DIST_KERNEL_1CTA) && sms >= 2 && out_tma_ok) {
    // ---- 2-CTA kernel: 256 x 256 tiles, 128-row boxes for both operands ----
    Gemm2Args ga;
    ga.g = g;
    ga.g.m_tiles = (int)((m1 + 255) / 256);
    ga.g.n_tiles = (int)((m2 + 255) / 256);
    ga.g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    ga.planes = need;
    ga.stages = kRing2Bytes / (2 * need * kTile2Bytes);
    if (ga.stages > kMaxStages2) ga.stages = kMaxStages2;
    ga.groups = 1; ga.a_group_rows = 0; ga.b_group_rows = 0; ga.out_group_cols = 0;
    int rc = make_operand_map(&tmA, a_planes, m1, a_plane_rows, kpad, a_planes_n, kT2Rows, f16);
    if (rc) return rc;
    rc = make_operand_map(&tmB, b_planes, m2, b_plane_rows, kpad, b_planes_n, kT2Rows, f16);
    if (rc) return rc;
    CUtensorMap tmO;
    {
      EncodeTiledFn fn = encode_fn();
      if (!fn) return cuda_fail(cudaErrorUnknown, "cuTensorMapEncodeTiled entry point not found");
      cuuint64_t dims[2] = {(cuuint64_t)m2, (cuuint64_t)m1};
      cuuint64_t strides[1] = {(cuuint64_t)ldd * 4};
      cuuint32_t box[2] = {32, 32};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = fn(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dist, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(output) failed");
    }
    static thread_local int configured2_dev = -1;
    if (configured2_dev != dev) {
      PPS_CUDA_TRY(cudaFuncSetAttribute(dist_tc2_kernel<256, EPI_DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kGemm2Smem));
      configured2_dev = dev;
    }
    const long long tiles2 = (long long)ga.g.m_tiles * ga.g.n_tiles;
    long long slots = sms / 2;
    if ((flags & PPS_DIST_RESERVE_SM_PAIR) && slots > 8) slots -= 1;   // leave one SM pair to concurrent small kernels
    const long long pairs = tiles2 < slots ? tiles2 : slots;
    dist_tc2_kernel<256, EPI_DIST><<<(unsigned)(2 * pairs), kGemmThreads, kGemm2Smem, st>>>(tmA, tmB, tmO, ga);
    PPS_LAUNCH_CHECK("dist_tc2_kernel");
    return PPS_OK;
  }

  int rc = make_operand_map(&tmA, a_planes, m1, a_plane_rows, kpad, a_planes_n, kBM, f16);
  if (rc) return rc;
  rc = make_operand_map(&tmB, b_planes, m2, b_plane_rows, kpad, b_planes_n, kBN, f16);
  if (rc) return rc;
  static thread_local int configured_dev = -1;
  if (configured_dev != dev) {
    PPS_CUDA_TRY(cudaFuncSetAttribute(dist_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    configured_dev = dev;
  }
  const long long tiles = (long long)g.m_tiles * g.n_tiles;
  const int grid = (int)(tiles < sms ? tiles : sms);
  dist_tc_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(tmA, tmB, g);
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

// ------------------------------------------------------------------------------------
// Next row: the per-combination embedding between pooling and distance
// (detectron/modeling/reid_heads.py:34-76 at test time): for every combination k
//   out[n, k*E + e] = max(0, alpha[k*E+e] * (sum_c pooled[k, n, c] * W[k, e, c]) + beta[k*E+e])
// i.e. Conv1x1(C -> E) + bias + SpatialBN(test: an affine map) + ReLU with the conv bias and the BN statistics
// folded into alpha / beta by the caller, written straight into the concatenated [N, K*E] feature
// (Concat(axis=1), :95-101).  K independent [N, C] x [C, E] products = one grouped launch of the 2-CTA kernel
// (256 x 128 tiles, tcgen05.mma.cta_group::2 with N = 128), same split-precision operands as the distance.
// ------------------------------------------------------------------------------------
extern "C" int pps_embed_tc(const void* x_planes /*[planes][K*N][kpad]*/, int x_planes_n, long long N,
                            const void* w_planes /*[planes][K*E][kpad]*/, int w_planes_n, int E, int K, int C,
                            const float* alpha, const float* beta, int precision, float* out /*[N, ldo >= K*E]*/,
                            long long ldo, void* stream) {
  if (N < 0 || K <= 0 || C <= 0 || E <= 0 || ldo < (long long)K * E) return PPS_ERR_INVALID_ARG;
  if (N == 0) return PPS_OK;
  if (E != 128) return PPS_ERR_UNSUPPORTED;             // cfg.REID.BPM_DIM of every shipped PPS config
  if (!x_planes || !w_planes || !alpha || !beta || !out) return PPS_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(x_planes) & 15u) || (reinterpret_cast<uintptr_t>(w_planes) & 15u)) return PPS_ERR_ALIGN;
  if ((ldo & 3) || (reinterpret_cast<uintptr_t>(out) & 15u)) return PPS_ERR_ALIGN;
  if ((long long)K * N > 0x7fffff00LL) return PPS_ERR_UNSUPPORTED;
  Gemm2Args ga;
  GemmArgs& g = ga.g;
  int need = 1;
  switch (precision) {
    case PPS_PREC_BF16X1: g.nterms = 1; g.term_a[0] = 0; g.term_b[0] = 0; need = 1; break;
    case PPS_PREC_BF16X3: {
      const int ta[3] = {1, 0, 0}, tb[3] = {0, 1, 0};
      g.nterms = 3; for (int i = 0; i < 3; ++i) { g.term_a[i] = ta[i]; g.term_b[i] = tb[i]; }
      need = 2; break;
    }
    case PPS_PREC_BF16X6: {
      const int ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
      g.nterms = 6; for (int i = 0; i < 6; ++i) { g.term_a[i] = ta[i]; g.term_b[i] = tb[i]; }
      need = 3; break;
    }
    default: return PPS_ERR_INVALID_ARG;
  }
  if (x_planes_n < need || w_planes_n < need) return PPS_ERR_INVALID_ARG;
  const int kpad = pps_kpad(C);
  g.m1 = N; g.m2 = (long long)K * E;
  g.kblocks = kpad / kBK;
  g.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  g.a_sqnorm = alpha; g.b_sqnorm = beta;
  g.out = out; g.ldo = ldo; g.flags = 0;
  g.m_tiles = (int)((N + 255) / 256);
  g.n_tiles = 1;
  ga.planes = need;
  ga.stages = kRing2Bytes / (need * (kTile2Bytes + 64 * kBK * 2));
  if (ga.stages > kMaxStages2) ga.stages = kMaxStages2;
  ga.groups = K; ga.a_group_rows = N; ga.b_group_rows = E; ga.out_group_cols = E;

  CUtensorMap tmA, tmB, tmO;
  int rc = make_operand_map(&tmA, x_planes, (long long)K * N, (long long)K * N, kpad, x_planes_n, kT2Rows, false);
  if (rc) return rc;
  rc = make_operand_map(&tmB, w_planes, (long long)K * E, (long long)K * E, kpad, w_planes_n, 64, false);
  if (rc) return rc;
  {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cuda_fail(cudaErrorUnknown, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {(cuuint64_t)((long long)K * E), (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)ldo * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(output) failed");
  }
  static thread_local int configured_dev = -1;
  int dev = 0;
  PPS_CUDA_TRY(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    PPS_CUDA_TRY(cudaFuncSetAttribute(dist_tc2_kernel<128, EPI_AFFINE_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kGemm2Smem));
    configured_dev = dev;
  }
  const int sms = sm_count();
  const long long tiles = (long long)g.m_tiles * K;
  const long long pairs = tiles < sms / 2 ? tiles : sms / 2;
  dist_tc2_kernel<128, EPI_AFFINE_RELU><<<(unsigned)(2 * pairs), kGemmThreads, kGemm2Smem,
                                          static_cast<cudaStream_t>(stream)>>>(tmA, tmB, tmO, ga);
  PPS_LAUNCH_CHECK("dist_tc2_kernel<embed>");
  return PPS_OK;
}

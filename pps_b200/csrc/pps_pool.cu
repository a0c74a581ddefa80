// Part-power-set pooling, fused: strip average/max pooling + all subset combinations.
//
// Replaces the Caffe2 sub-graph the reference emits per image
//   Split(axis=2) -> n x {AveragePool, MaxPool}(global)          bpm_heads.py:41-55
//   for m in 1..2^n-1: Mean(avg_j, j in S_m) / Max(max_j) / Add     pps_heads.py:47-76
// (~200 tiny launches for n = 6) with one persistent kernel that reads every conv5 plane
// once and writes every combination once: 4*C*H*W bytes in + 4*K*C bytes out per image.
//
// Data movement (HBM-bound kernel, no tensor cores):
//   * a "unit" is one image x 32 consecutive channels: its planes are one contiguous run of
//     32*H*W floats in NCHW, so a producer thread fetches them with 1-D bulk TMA copies
//     (cp.async.bulk, completion on an mbarrier) into a 4-stage shared-memory ring;
//   * 8 consumer warps reduce (plane, strip) pairs out of shared memory with 128-bit loads
//     (lane-rotated start so a quarter-warp touches 32 distinct banks), leave avg/max in a
//     double-buffered [strip][channel] table, and
//   * combine: lane = channel, warp = subset mask, so every output row is one coalesced
//     128-byte streaming store.
#include "common.cuh"

#include <cfloat>

namespace pps {

constexpr int kCB = 32;                  // channels per unit
constexpr int kPoolStages = 3;           // shared-memory ring depth (2 CTAs per SM -> 6 slots in flight per SM)
constexpr int kStageBytes = 32 * 1024;   // bytes per ring slot
constexpr int kConsumerWarps = 8;
constexpr int kPoolThreads = 32 * (1 + kConsumerWarps);
constexpr int kMaxComboList = 512;
constexpr int kPoolPlanesOut = 2;       // template MODE bit: write bf16 operand planes instead of fp32 values

struct PoolArgs {
  const float* x;
  float* y;
  int N, C, H, W;
  int n_parts, mode;
  int n_out;            // combinations emitted
  int use_list;         // 1: masks come from combos[]
  int planes_per_stage;
  long long ysn, ysk;
  // operand-plane output (pps_pool_planes_fwd): instead of fp32 values, write the bf16 residual planes the tensor-core
  // embedding consumes (plane p = bf16(v - sum of the planes before it), as split_rows_kernel does), same element
  // offsets, plane p at y_planes + p * plane_stride
  __nv_bfloat16* y_planes;
  int out_planes;
  long long plane_stride;
  int row0[PPS_POOL_MAX_PARTS + 1];
  float inv_cnt[PPS_POOL_MAX_PARTS + 1];   // 1.0f / k  (Caffe2 Mean scales the sum by 1.0f/InputSize())
  int combos[kMaxComboList];
};

// All output rows of one unit.  lane = channel: the thread keeps its channel's NP strip averages / maxes in
// registers and walks this warp's share of the subset masks with a fully unrolled, predicated bit loop
// (no per-bit shared-memory reads, no data-dependent branches).  Sums run over ascending part index, like
// Caffe2 Mean (sum in input order, then one multiply by 1.0f / count).
template <int NP, int MODE>
__device__ __forceinline__ void combine_unit(const PoolArgs& a, const float* pavg, const float* pmax, int cw,
                                             int nwarps, int lane, int nch, long long yoff) {
  float* ybase = a.y + yoff;
  float av[NP], mv[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    av[j] = pavg[j * kCB + lane];
    mv[j] = (MODE & 1) == PPS_POOL_MAX_AVE ? pmax[j * kCB + lane] : 0.f;
  }
  for (int idx = cw; idx < a.n_out; idx += nwarps) {
    const int m = a.use_list ? a.combos[idx] : idx + 1;
    float val;
    if ((MODE & 1) == PPS_POOL_MAX_AVE) {
      float s = 0.f, mx = -FLT_MAX;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const bool on = (m >> j) & 1;
        s = on ? s + av[j] : s;
        mx = on ? fmaxf(mx, mv[j]) : mx;
      }
      const int cnt = __popc(m);
      val = (cnt > 1 ? s * a.inv_cnt[cnt] : s) + mx;
    } else {
      float mx = -FLT_MAX;
#pragma unroll
      for (int j = 0; j < NP; ++j) mx = ((m >> j) & 1) ? fmaxf(mx, av[j]) : mx;
      val = mx;
    }
    if (!(MODE & kPoolPlanesOut)) {
      if (lane < nch) st_stream_f32(ybase + (long long)idx * a.ysk + lane, val);
    } else if (lane < nch) {
      float r = val;
      __nv_bfloat16* dst = a.y_planes + yoff + (long long)idx * a.ysk + lane;
      for (int p = 0; p < a.out_planes; ++p) {
        const __nv_bfloat16 h = __float2bfloat16_rn(r);
        dst[(long long)p * a.plane_stride] = h;
        r -= __bfloat162float(h);                          // exact in fp32
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// fast path: W % 4 == 0, 16-byte aligned x, one plane fits a ring slot
// ------------------------------------------------------------------------------------
template <int NP, int MODE>
__global__ void __launch_bounds__(kPoolThreads, 2) pool_tma_kernel(const __grid_constant__ PoolArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);                               // [stages][kStageBytes/4]
  float* pavg = reinterpret_cast<float*>(smem_raw + kPoolStages * kStageBytes);    // [2][MAXP][kCB]
  float* pmax = pavg + 2 * PPS_POOL_MAX_PARTS * kCB;                               // [2][MAXP][kCB]
  uint64_t* full = reinterpret_cast<uint64_t*>(pmax + 2 * PPS_POOL_MAX_PARTS * kCB);
  uint64_t* empty = full + kPoolStages;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W;
  const int cblocks = (a.C + kCB - 1) / kCB;
  const long long units = (long long)a.N * cblocks;
  const int P = a.planes_per_stage;

  if (tid == 0) {
    for (int s = 0; s < kPoolStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ===== producer: one thread streams the planes of this CTA's units into the ring =====
    if (lane == 0) {
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const long long n = u / cblocks;
        const int c0 = (int)(u % cblocks) * kCB;
        const int nch = min(kCB, a.C - c0);
        const float* src = a.x + (n * a.C + c0) * (long long)HW;
        for (int p0 = 0; p0 < nch; p0 += P, ++it) {
          const int npl = min(P, nch - p0);
          const uint32_t s = it % kPoolStages, ph = (it / kPoolStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const uint32_t bytes = (uint32_t)npl * (uint32_t)HW * 4u;
          mbar_expect_tx(&full[s], bytes);
          bulk_load_1d(stage + (size_t)s * (kStageBytes / 4), src + (long long)p0 * HW, bytes, &full[s]);
        }
      }
    }
  } else {
    // ===== consumers =====
    const int ctid = tid - 32;
    const int cw = warp - 1;
    uint32_t it = 0;
    int ubuf = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const long long n = u / cblocks;
      const int c0 = (int)(u % cblocks) * kCB;
      const int nch = min(kCB, a.C - c0);
      float* ua = pavg + ubuf * PPS_POOL_MAX_PARTS * kCB;
      float* um = pmax + ubuf * PPS_POOL_MAX_PARTS * kCB;
      for (int p0 = 0; p0 < nch; p0 += P, ++it) {
        const int npl = min(P, nch - p0);
        const uint32_t s = it % kPoolStages, ph = (it / kPoolStages) & 1u;
        mbar_wait(&full[s], ph);
        const float* sbase = stage + (size_t)s * (kStageBytes / 4);
        const int tasks = npl * NP;
        for (int t = ctid; t < tasks; t += 32 * kConsumerWarps) {
          const int pl = t / NP, j = t - pl * NP;
          const int r0 = a.row0[j], r1 = a.row0[j + 1];
          const int L4 = ((r1 - r0) * a.W) >> 2;
          const float4* b4 = reinterpret_cast<const float4*>(sbase + pl * HW + r0 * a.W);
          int idx = lane % L4;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll 4
          for (int q = 0; q < L4; ++q) {
            const float4 v = b4[idx];
            idx = (idx + 1 == L4) ? 0 : idx + 1;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
          }
          const float sum = (acc.x + acc.y) + (acc.z + acc.w);
          ua[j * kCB + p0 + pl] = __fdiv_rn(sum, (float)((r1 - r0) * a.W));
          if ((MODE & 1) == PPS_POOL_MAX_AVE) um[j * kCB + p0 + pl] = fmaxf(fmaxf(mx.x, mx.y), fmaxf(mx.z, mx.w));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);   // slot may be refilled
      }
      // all strip results of this unit are in ua/um
      asm volatile("bar.sync 1, %0;" ::"r"(32 * kConsumerWarps) : "memory");
      combine_unit<NP, MODE>(a, ua, um, cw, kConsumerWarps, lane, nch, n * a.ysn + c0);
      ubuf ^= 1;
    }
  }
}

// ------------------------------------------------------------------------------------
// generic path: any W / alignment / plane size. One CTA per unit, plain global loads.
// ------------------------------------------------------------------------------------
template <int NP, int MODE>
__global__ void __launch_bounds__(256) pool_generic_kernel(const __grid_constant__ PoolArgs a) {
  __shared__ float pavg[PPS_POOL_MAX_PARTS * kCB];
  __shared__ float pmax[PPS_POOL_MAX_PARTS * kCB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = a.H * a.W;
  const int cblocks = (a.C + kCB - 1) / kCB;
  const long long u = blockIdx.x;
  const long long n = u / cblocks;
  const int c0 = (int)(u % cblocks) * kCB;
  const int nch = min(kCB, a.C - c0);
  for (int pl = warp; pl < nch; pl += 8) {
    const float* plane = a.x + (n * a.C + c0 + pl) * (long long)HW;
    for (int j = 0; j < NP; ++j) {
      const int e0 = a.row0[j] * a.W, e1 = a.row0[j + 1] * a.W;
      float s = 0.f, mx = -FLT_MAX;
      for (int e = e0 + lane; e < e1; e += 32) {
        const float v = __ldg(plane + e);
        s += v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (lane == 0) {
        pavg[j * kCB + pl] = __fdiv_rn(s, (float)(e1 - e0));
        pmax[j * kCB + pl] = mx;
      }
    }
  }
  __syncthreads();
  combine_unit<NP, MODE>(a, pavg, pmax, warp, 8, lane, nch, n * a.ysn + c0);
}

// ------------------------------------------------------------------------------------
// Backward of the fused pooling (the reference's multi-scale branch is train-only, pps_heads.py:88-142; a custom op
// registers its gradient like detectron/ops/pairwise_distance_op.cc:14-24).  Gradients of the stock Caffe2 operators
// the sub-graph is made of:
//   Mean(avg_j, j in S)      d avg_j += dY_S / |S|
//   Max(v_j, j in S)         d v_j   += dY_S where v_j == max          (Caffe2 MaxGradient: EVERY tied input gets dY)
//   Add                      both inputs get dY
//   AveragePool(global)      dX[e] = d avg_j / (h_j W) for every element of strip j
//   MaxPool(global)          dX[arg max] = d max_j  (first maximal element in row-major order on ties)
// One CTA per unit (image x 32 channels): (1) strip average / max / arg-max of every plane from x, (2) lane = channel walks
// the combinations and accumulates d avg_j, d max_j from dY, (3) dX written once, coalesced.  HBM traffic per image:
// x read once + dY read once + dX written once = 2 * 4 C H W + 4 K C bytes.
// ------------------------------------------------------------------------------------
struct PoolBwdArgs {
  const float* dy;
  float* dx;
  long long dysn, dysk;
};

template <int NP, int MODE>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const __grid_constant__ PoolArgs a, const __grid_constant__ PoolBwdArgs b) {
  __shared__ float pavg[NP * kCB];
  __shared__ float pmax[NP * kCB];
  __shared__ int parg[NP * kCB];
  __shared__ float d_avg[NP * kCB];      // gradient w.r.t. the strip average, already divided by the strip size
  __shared__ float d_max[NP * kCB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = a.H * a.W;
  const int cblocks = (a.C + kCB - 1) / kCB;
  const long long u = blockIdx.x;
  const long long n = u / cblocks;
  const int c0 = (int)(u % cblocks) * kCB;
  const int nch = min(kCB, a.C - c0);
  for (int i = threadIdx.x; i < NP * kCB; i += 256) { d_avg[i] = 0.f; d_max[i] = 0.f; }
  // (1) forward statistics of every plane
  for (int pl = warp; pl < nch; pl += 8) {
    const float* plane = a.x + (n * a.C + c0 + pl) * (long long)HW;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int e0 = a.row0[j] * a.W, e1 = a.row0[j + 1] * a.W;
      float s = 0.f, mx = -FLT_MAX;
      int arg = e1;
      for (int e = e0 + lane; e < e1; e += 32) {
        const float v = __ldg(plane + e);
        s += v;
        if (v > mx) { mx = v; arg = e; }        // ascending e per lane: the first maximum of the lane
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
        if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
      }
      if (lane == 0) {
        pavg[j * kCB + pl] = __fdiv_rn(s, (float)(e1 - e0));
        pmax[j * kCB + pl] = mx;
        parg[j * kCB + pl] = arg;
      }
    }
  }
  __syncthreads();
  // (2) lane = channel: gradients of the strip statistics from the combinations' gradients
  if (lane < nch) {
    float av[NP], mv[NP], ga[NP], gm[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      av[j] = pavg[j * kCB + lane];
      mv[j] = pmax[j * kCB + lane];
      ga[j] = 0.f; gm[j] = 0.f;
    }
    const float* dyb = b.dy + n * b.dysn + c0 + lane;
    for (int idx = warp; idx < a.n_out; idx += 8) {
      const int m = a.use_list ? a.combos[idx] : idx + 1;
      const float g = __ldg(dyb + (long long)idx * b.dysk);
      if (MODE == PPS_POOL_MAX_AVE) {
        float mx = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < NP; ++j) mx = ((m >> j) & 1) ? fmaxf(mx, mv[j]) : mx;
        const int cnt = __popc(m);
        const float gs = cnt > 1 ? g * a.inv_cnt[cnt] : g;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const bool on = (m >> j) & 1;
          ga[j] += on ? gs : 0.f;
          gm[j] += (on && mv[j] == mx) ? g : 0.f;
        }
      } else {
        float mx = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < NP; ++j) mx = ((m >> j) & 1) ? fmaxf(mx, av[j]) : mx;
#pragma unroll
        for (int j = 0; j < NP; ++j) ga[j] += (((m >> j) & 1) && av[j] == mx) ? g : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float inv = 1.0f / (float)((a.row0[j + 1] - a.row0[j]) * a.W);
      atomicAdd(&d_avg[j * kCB + lane], ga[j] * inv);
      if (MODE == PPS_POOL_MAX_AVE) atomicAdd(&d_max[j * kCB + lane], gm[j]);
    }
  }
  __syncthreads();
  // (3) dX
  for (int pl = warp; pl < nch; pl += 8) {
    float* dplane = b.dx + (n * a.C + c0 + pl) * (long long)HW;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int e0 = a.row0[j] * a.W, e1 = a.row0[j + 1] * a.W;
      const float base = d_avg[j * kCB + pl];
      const float gmx = MODE == PPS_POOL_MAX_AVE ? d_max[j * kCB + pl] : 0.f;
      const int arg = parg[j * kCB + pl];
      for (int e = e0 + lane; e < e1; e += 32) st_stream_f32(dplane + e, base + (e == arg ? gmx : 0.f));
    }
  }
}

template <int NP>
static int launch_pool_bwd(const PoolArgs& a, const PoolBwdArgs& b, long long units, cudaStream_t st) {
  if (a.mode == PPS_POOL_MAX_AVE) pool_bwd_kernel<NP, PPS_POOL_MAX_AVE><<<(int)units, 256, 0, st>>>(a, b);
  else pool_bwd_kernel<NP, PPS_POOL_AVG_MAX><<<(int)units, 256, 0, st>>>(a, b);
  PPS_LAUNCH_CHECK("pool_bwd_kernel");
  return PPS_OK;
}

template <int NP, int MODE>
static int launch_pool(const PoolArgs& a, bool fast, long long units, size_t smem, cudaStream_t st) {
  if (fast) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    PPS_CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      PPS_CUDA_TRY(cudaFuncSetAttribute(pool_tma_kernel<NP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured_dev = dev;
    }
    const long long slots = 2LL * sm_count();   // two resident CTAs per SM
    const int grid = (int)(units < slots ? units : slots);
    pool_tma_kernel<NP, MODE><<<grid, kPoolThreads, smem, st>>>(a);
    PPS_LAUNCH_CHECK("pool_tma_kernel");
  } else {
    pool_generic_kernel<NP, MODE><<<(int)units, 256, 0, st>>>(a);
    PPS_LAUNCH_CHECK("pool_generic_kernel");
  }
  return PPS_OK;
}

template <int MODE>
static int dispatch_parts(const PoolArgs& a, bool fast, long long units, size_t smem, cudaStream_t st) {
  switch (a.n_parts) {
    case 1: return launch_pool<1, MODE>(a, fast, units, smem, st);
    case 2: return launch_pool<2, MODE>(a, fast, units, smem, st);
    case 3: return launch_pool<3, MODE>(a, fast, units, smem, st);
    case 4: return launch_pool<4, MODE>(a, fast, units, smem, st);
    case 5: return launch_pool<5, MODE>(a, fast, units, smem, st);
    case 6: return launch_pool<6, MODE>(a, fast, units, smem, st);
    case 7: return launch_pool<7, MODE>(a, fast, units, smem, st);
    case 8: return launch_pool<8, MODE>(a, fast, units, smem, st);
    case 9: return launch_pool<9, MODE>(a, fast, units, smem, st);
    case 10: return launch_pool<10, MODE>(a, fast, units, smem, st);
    default: return PPS_ERR_SHAPE;
  }
}

}  // namespace pps

using namespace pps;

// shape checks + strip table + combination list shared by the forward and the backward
static int pool_setup(PoolArgs& a, int N, int C, int H, int W, int n_parts, const int* split, int mode, const int* combos,
                      int n_combos) {
  if (N < 0 || C <= 0 || H <= 0 || W <= 0 || !split) return PPS_ERR_INVALID_ARG;
  if (mode != PPS_POOL_AVG_MAX && mode != PPS_POOL_MAX_AVE) return PPS_ERR_INVALID_ARG;
  if (n_parts < 1 || n_parts > PPS_POOL_MAX_PARTS) return PPS_ERR_SHAPE;
  a.row0[0] = 0;
  for (int j = 0; j < n_parts; ++j) {
    if (split[j] <= 0) return PPS_ERR_SHAPE;
    a.row0[j + 1] = a.row0[j] + split[j];
  }
  if (a.row0[n_parts] != H) return PPS_ERR_SHAPE;   // Caffe2 Split enforces sum(split) == dim
  for (int j = n_parts + 1; j <= PPS_POOL_MAX_PARTS; ++j) a.row0[j] = H;
  a.inv_cnt[0] = 0.f;
  for (int k = 1; k <= PPS_POOL_MAX_PARTS; ++k) a.inv_cnt[k] = 1.0f / (float)k;
  const int full_mask = (1 << n_parts) - 1;
  if (combos) {
    if (n_combos < 1 || n_combos > kMaxComboList) return PPS_ERR_SHAPE;
    for (int i = 0; i < n_combos; ++i) {
      if (combos[i] < 1 || combos[i] > full_mask) return PPS_ERR_SHAPE;
      a.combos[i] = combos[i];
    }
    a.use_list = 1;
    a.n_out = n_combos;
  } else {
    a.use_list = 0;
    a.n_out = full_mask;
  }
  a.N = N; a.C = C; a.H = H; a.W = W;
  a.n_parts = n_parts; a.mode = mode;
  a.x = nullptr; a.y = nullptr; a.y_planes = nullptr; a.out_planes = 0; a.plane_stride = 0; a.ysn = 0; a.ysk = 0;
  a.planes_per_stage = 0;
  return PPS_OK;
}

static int pool_launch(const float* x, int N, int C, int H, int W, int n_parts, const int* split, int mode,
                       const int* combos, int n_combos, float* y, void* y_planes, int out_planes, long long plane_stride,
                       long long y_stride_n, long long y_stride_k, void* stream) {
  PoolArgs a;
  const int rc = pool_setup(a, N, C, H, W, n_parts, split, mode, combos, n_combos);
  if (rc != PPS_OK) return rc;
  if (N == 0) return PPS_OK;
  if (!x || (!y && !y_planes)) return PPS_ERR_INVALID_ARG;
  a.x = x; a.y = y;
  a.y_planes = static_cast<__nv_bfloat16*>(y_planes); a.out_planes = y_planes ? out_planes : 0; a.plane_stride = plane_stride;
  a.ysn = y_stride_n; a.ysk = y_stride_k;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const long long plane_bytes = (long long)H * W * 4;
  const int cblocks = (C + kCB - 1) / kCB;
  const long long units = (long long)N * cblocks;
  const bool fast = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && plane_bytes <= kStageBytes;
  if (units > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
  a.planes_per_stage = fast ? (int)((kStageBytes / plane_bytes) < kCB ? (kStageBytes / plane_bytes) : kCB) : 0;
  const size_t smem = (size_t)kPoolStages * kStageBytes + 4 * PPS_POOL_MAX_PARTS * kCB * sizeof(float) +
                      2 * kPoolStages * sizeof(uint64_t);
  if (a.out_planes)
    return mode == PPS_POOL_MAX_AVE ? dispatch_parts<PPS_POOL_MAX_AVE | kPoolPlanesOut>(a, fast, units, smem, st)
                                    : dispatch_parts<PPS_POOL_AVG_MAX | kPoolPlanesOut>(a, fast, units, smem, st);
  return mode == PPS_POOL_MAX_AVE ? dispatch_parts<PPS_POOL_MAX_AVE>(a, fast, units, smem, st)
                                  : dispatch_parts<PPS_POOL_AVG_MAX>(a, fast, units, smem, st);
}

extern "C" int pps_pool_fwd(const float* x, int N, int C, int H, int W, int n_parts, const int* split, int mode,
                            const int* combos, int n_combos, float* y, long long y_stride_n, long long y_stride_k,
                            void* stream) {
  return pool_launch(x, N, C, H, W, n_parts, split, mode, combos, n_combos, y, nullptr, 0, 0, y_stride_n, y_stride_k, stream);
}

// Pooling straight into the operand planes of the tensor-core embedding (pps_embed_tc): output k of image n is row
// k * N + n of a [planes][K * N][kpad] bf16 buffer, i.e. the [K, N, C] layout already split the way pps_split_rows
// would split it - the fp32 pooled intermediate and its split pass never touch HBM.
extern "C" int pps_pool_planes_fwd(const float* x, int N, int C, int H, int W, int n_parts, const int* split, int mode,
                                   const int* combos, int n_combos, void* out_planes, int planes, void* stream) {
  if (planes < 1 || planes > 3) return PPS_ERR_INVALID_ARG;
  if (N > 0 && !out_planes) return PPS_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(out_planes) & 15u) return PPS_ERR_ALIGN;
  if (n_parts < 1 || n_parts > PPS_POOL_MAX_PARTS) return PPS_ERR_SHAPE;
  const int kpad = pps_kpad(C);
  const long long K = combos ? n_combos : ((1LL << n_parts) - 1);
  const long long rows = K * (long long)N;
  if (kpad != C && rows > 0)       // K padding of every row must read as zero
    PPS_CUDA_TRY(cudaMemsetAsync(out_planes, 0, (size_t)planes * rows * kpad * 2, static_cast<cudaStream_t>(stream)));
  return pool_launch(x, N, C, H, W, n_parts, split, mode, combos, n_combos, nullptr, out_planes, planes, rows * kpad,
                     (long long)kpad, (long long)N * kpad, stream);
}

// dX of pps_pool_fwd (see pool_bwd_kernel): x the forward input, dy the gradient of the forward output addressed like y
// (element (n, k, c) at dy[n*dy_stride_n + k*dy_stride_k + c]), dx [N, C, H, W] written in full.
extern "C" int pps_pool_bwd(const float* x, const float* dy, int N, int C, int H, int W, int n_parts, const int* split,
                            int mode, const int* combos, int n_combos, long long dy_stride_n, long long dy_stride_k,
                            float* dx, void* stream) {
  PoolArgs a;
  const int rc = pool_setup(a, N, C, H, W, n_parts, split, mode, combos, n_combos);
  if (rc != PPS_OK) return rc;
  if (N == 0) return PPS_OK;
  if (!x || !dy || !dx) return PPS_ERR_INVALID_ARG;
  a.x = x;
  PoolBwdArgs b;
  b.dy = dy; b.dx = dx; b.dysn = dy_stride_n; b.dysk = dy_stride_k;
  const long long units = (long long)N * ((C + kCB - 1) / kCB);
  if (units > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (n_parts) {
    case 1: return launch_pool_bwd<1>(a, b, units, st);
    case 2: return launch_pool_bwd<2>(a, b, units, st);
    case 3: return launch_pool_bwd<3>(a, b, units, st);
    case 4: return launch_pool_bwd<4>(a, b, units, st);
    case 5: return launch_pool_bwd<5>(a, b, units, st);
    case 6: return launch_pool_bwd<6>(a, b, units, st);
    case 7: return launch_pool_bwd<7>(a, b, units, st);
    case 8: return launch_pool_bwd<8>(a, b, units, st);
    case 9: return launch_pool_bwd<9>(a, b, units, st);
    case 10: return launch_pool_bwd<10>(a, b, units, st);
    default: return PPS_ERR_SHAPE;
  }
}

// Training-side triplet mining ops of the reference, the only custom CUDA on its re-ID path:
//   PairWiseDistance / PairWiseDistanceGradient   detectron/ops/pairwise_distance_op.cu:9-22, :78-91
//   BatchHard / BatchHardGradient                 detectron/ops/batch_hard_op.cc:9-59, :62-123
//                                                 (GPU = GPUFallbackOp: D2H -> CPU loop -> H2D, batch_hard_op.cu:9-10)
// Shapes are tiny (N = 64..256 anchors, D = 128), so these are latency-bound: the reference spends its time in
// atomics (gradient: 2*N*N*D fp32 atomicAdds) and in the host round trip of BatchHard.  Here:
//   * forward: 32 x 32 output tile per CTA from shared-memory row tiles, exact (x_p - x_q)^2 like the reference;
//   * gradient: dX[n,d] = 2 * sum_q (x_n[d] - x_q[d]) * (dZ[n,q] + dZ[q,n])  - one thread per (n,d), no atomics,
//     deterministic;
//   * BatchHard on the device, one warp per anchor (first index wins ties, as the reference's strict compares do);
//   * a fused forward that mines the hardest positive / negative straight from the features without ever
//     writing the [N, N] matrix.
#include "common.cuh"

#include <cfloat>

namespace pps {

constexpr int kTT = 32;

__global__ void __launch_bounds__(kTT * 8) pairwise_sqdist_kernel(const float* __restrict__ x, int N, int D,
                                                                  float* __restrict__ z) {
  __shared__ float sp[kTT][kTT + 1];
  __shared__ float sq[kTT][kTT + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8 threads, 4 rows each
  const int p0 = blockIdx.y * kTT, q0 = blockIdx.x * kTT;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int d0 = 0; d0 < D; d0 += kTT) {
    for (int r = ty; r < kTT; r += 8) {
      const int d = d0 + tx;
      sp[r][tx] = (p0 + r < N && d < D) ? x[(long long)(p0 + r) * D + d] : 0.f;
      sq[r][tx] = (q0 + r < N && d < D) ? x[(long long)(q0 + r) * D + d] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < kTT; ++d) {
      const float b = sq[tx][d];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float sub = sp[ty * 4 + i][d] - b;
        acc[i] = fmaf(sub, sub, acc[i]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty * 4 + i, q = q0 + tx;
    if (p < N && q < N) z[(long long)p * N + q] = acc[i];
  }
}

__global__ void pairwise_sqdist_grad_kernel(const float* __restrict__ x, const float* __restrict__ dz, int N, int D,
                                            float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * D) return;
  const int n = (int)(i / D), d = (int)(i % D);
  const float xn = x[i];
  float acc = 0.f;
  for (int q = 0; q < N; ++q) {
    const float w = dz[(long long)n * N + q] + dz[(long long)q * N + n];
    acc = fmaf(xn - x[(long long)q * D + d], w, acc);
  }
  dx[i] = 2.f * acc;
}

// one warp per anchor; `row(j)` yields the distance of anchor a to item j
template <class RowFn>
__device__ __forceinline__ void hardest_of_row(RowFn row, const int* __restrict__ labels, int a, int N, int lane,
                                               float* ap, float* an, int* idx_p, int* idx_n) {
  const int la = labels[a];
  float best_p = 0.f, best_n = FLT_MAX;          // initial values of batch_hard_op.cc:33,45
  int ip = -1, in = -1;
  for (int j = lane; j < N; j += 32) {
    const float v = row(j);
    if (labels[j] == la) {
      if (best_p < v) { best_p = v; ip = j; }
    } else {
      if (best_n > v) { best_n = v; in = j; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float op = __shfl_xor_sync(0xffffffffu, best_p, o), on = __shfl_xor_sync(0xffffffffu, best_n, o);
    const int oip = __shfl_xor_sync(0xffffffffu, ip, o), oin = __shfl_xor_sync(0xffffffffu, in, o);
    // strict compares in index order == the lowest index among equal extrema wins
    if (oip >= 0 && (op > best_p || (op == best_p && (ip < 0 || oip < ip)))) { best_p = op; ip = oip; }
    if (oin >= 0 && (on < best_n || (on == best_n && (in < 0 || oin < in)))) { best_n = on; in = oin; }
  }
  if (lane == 0) {
    ap[a] = best_p;
    an[a] = best_n;
    if (idx_p) idx_p[a] = ip;
    if (idx_n) idx_n[a] = in;
  }
}

__global__ void batch_hard_kernel(const float* __restrict__ xd, const int* __restrict__ labels, int N,
                                  float* __restrict__ ap, float* __restrict__ an, int* __restrict__ idx_p,
                                  int* __restrict__ idx_n) {
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= N) return;
  const float* r = xd + (long long)a * N;
  hardest_of_row([r](int j) { return r[j]; }, labels, a, N, threadIdx.x & 31, ap, an, idx_p, idx_n);
}

// fused: squared distances of 8 anchors (one per warp) to every item straight from the features, no [N, N] matrix.
// The CTA stages a 128-wide slice of its 8 anchor rows and of 32 item rows in shared memory (coalesced loads, rows
// padded by one word: conflict-free), thread (warp = anchor, lane = item) accumulates its pair over d in ascending
// order - the same fmaf chain as pairwise_sqdist_kernel, so the mined values equal the two-step result.
constexpr int kFA = 8;          // anchors per CTA
constexpr int kFD = 128;        // feature slice

__global__ void __launch_bounds__(32 * kFA) batch_hard_fused_kernel(const float* __restrict__ x, const int* __restrict__ labels,
                                                                     int N, int D, float* __restrict__ ap,
                                                                     float* __restrict__ an, int* __restrict__ idx_p,
                                                                     int* __restrict__ idx_n) {
  __shared__ float sa[kFA][kFD + 1];
  __shared__ float sj[32][kFD + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a = blockIdx.x * kFA + w;
  const int la = a < N ? labels[a] : 0;
  float best_p = 0.f, best_n = FLT_MAX;          // initial values of batch_hard_op.cc:33,45
  int ip = -1, in = -1;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    float acc = 0.f;
    for (int d0 = 0; d0 < D; d0 += kFD) {
      const int dn = min(kFD, D - d0);
      __syncthreads();
      for (int i = threadIdx.x; i < kFA * kFD; i += 32 * kFA) {
        const int r = i / kFD, c = i % kFD, row = blockIdx.x * kFA + r;
        sa[r][c] = (row < N && c < dn) ? x[(long long)row * D + d0 + c] : 0.f;
      }
      for (int i = threadIdx.x; i < 32 * kFD; i += 32 * kFA) {
        const int r = i / kFD, c = i % kFD, row = j0 + r;
        sj[r][c] = (row < N && c < dn) ? x[(long long)row * D + d0 + c] : 0.f;
      }
      __syncthreads();
      for (int c = 0; c < dn; ++c) {
        const float sub = sa[w][c] - sj[lane][c];
        acc = fmaf(sub, sub, acc);
      }
    }
    if (a < N && j < N) {
      if (labels[j] == la) {
        if (best_p < acc) { best_p = acc; ip = j; }
      } else {
        if (best_n > acc) { best_n = acc; in = j; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float op = __shfl_xor_sync(0xffffffffu, best_p, o), on = __shfl_xor_sync(0xffffffffu, best_n, o);
    const int oip = __shfl_xor_sync(0xffffffffu, ip, o), oin = __shfl_xor_sync(0xffffffffu, in, o);
    // strict compares in index order == the lowest index among equal extrema wins
    if (oip >= 0 && (op > best_p || (op == best_p && (ip < 0 || oip < ip)))) { best_p = op; ip = oip; }
    if (oin >= 0 && (on < best_n || (on == best_n && (in < 0 || oin < in)))) { best_n = on; in = oin; }
  }
  if (lane == 0 && a < N) {
    ap[a] = best_p;
    an[a] = best_n;
    if (idx_p) idx_p[a] = ip;
    if (idx_n) idx_n[a] = in;
  }
}

__global__ void batch_hard_grad_kernel(const int* __restrict__ idx_p, const int* __restrict__ idx_n,
                                       const float* __restrict__ dap, const float* __restrict__ dan, int N,
                                       float* __restrict__ dx) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  // the reference writes dX[a*N + idx] even when idx == -1 (no candidate), i.e. out of the row; skipped here
  if (idx_p[a] >= 0) dx[(long long)a * N + idx_p[a]] = dap[a];
  if (idx_n[a] >= 0) dx[(long long)a * N + idx_n[a]] = dan[a];
}

}  // namespace pps

using namespace pps;

extern "C" int pps_pairwise_distance_fwd(const float* x, int N, int D, float* z, void* stream) {
  if (N < 0 || D <= 0) return PPS_ERR_SHAPE;
  if (N == 0) return PPS_OK;
  if (!x || !z) return PPS_ERR_INVALID_ARG;
  const unsigned t = (unsigned)((N + kTT - 1) / kTT);
  pairwise_sqdist_kernel<<<dim3(t, t), kTT * 8, 0, static_cast<cudaStream_t>(stream)>>>(x, N, D, z);
  PPS_LAUNCH_CHECK("pairwise_sqdist_kernel");
  return PPS_OK;
}

extern "C" int pps_pairwise_distance_bwd(const float* x, const float* dz, int N, int D, float* dx, void* stream) {
  if (N < 0 || D <= 0) return PPS_ERR_SHAPE;
  if (N == 0) return PPS_OK;
  if (!x || !dz || !dx) return PPS_ERR_INVALID_ARG;
  const long long n = (long long)N * D;
  pairwise_sqdist_grad_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, dz, N, D, dx);
  PPS_LAUNCH_CHECK("pairwise_sqdist_grad_kernel");
  return PPS_OK;
}

extern "C" int pps_batch_hard_fwd(const float* xdist, const int32_t* labels, int N, float* ap, float* an,
                                  int32_t* idx_p, int32_t* idx_n, void* stream) {
  if (N < 0) return PPS_ERR_SHAPE;
  if (N == 0) return PPS_OK;
  if (!xdist || !labels || !ap || !an) return PPS_ERR_INVALID_ARG;
  batch_hard_kernel<<<(unsigned)((N + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(xdist, labels, N, ap, an, idx_p,
                                                                                          idx_n);
  PPS_LAUNCH_CHECK("batch_hard_kernel");
  return PPS_OK;
}

extern "C" int pps_batch_hard_fused_fwd(const float* x, const int32_t* labels, int N, int D, float* ap, float* an,
                                        int32_t* idx_p, int32_t* idx_n, void* stream) {
  if (N < 0 || D <= 0) return PPS_ERR_SHAPE;
  if (N == 0) return PPS_OK;
  if (!x || !labels || !ap || !an) return PPS_ERR_INVALID_ARG;
  batch_hard_fused_kernel<<<(unsigned)((N + kFA - 1) / kFA), 32 * kFA, 0, static_cast<cudaStream_t>(stream)>>>(x, labels, N, D, ap,
                                                                                                             an, idx_p, idx_n);
  PPS_LAUNCH_CHECK("batch_hard_fused_kernel");
  return PPS_OK;
}

extern "C" int pps_batch_hard_bwd(const int32_t* idx_p, const int32_t* idx_n, const float* dap, const float* dan, int N,
                                  float* dx, void* stream) {
  if (N < 0) return PPS_ERR_SHAPE;
  if (N == 0) return PPS_OK;
  if (!idx_p || !idx_n || !dap || !dan || !dx) return PPS_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PPS_CUDA_TRY(cudaMemsetAsync(dx, 0, (size_t)N * N * sizeof(float), st));
  batch_hard_grad_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(idx_p, idx_n, dap, dan, N, dx);
  PPS_LAUNCH_CHECK("batch_hard_grad_kernel");
  return PPS_OK;
}

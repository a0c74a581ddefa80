// k-reciprocal re-ranking (Zhong et al., CVPR 2017) as the reference runs it after the plain ranking:
// detectron/datasets/reid_dataset_evaluator.py:442-519 `re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6,
// lambda_value=0.3)`, called from evaluate() at :161-207 (cfg.REID.RERANK defaults to True, config.py:1022).
// SURVEY §8f row 2.  The reference loops over all N = nq + ng images in Python on dense [N, N] float32 arrays; here
// every step is a kernel over sparse rows:
//   rerank_normalize_kernel   OD[i][j] = M[j][i]^2 / max_r M[r][i]^2                            (:447-454)
//   (pps_topk_*)              initial_rank[i][:k1+1]: the k1+1 nearest columns of every row      (:456)
//   rerank_krecip_kernel      k-reciprocal set, its 1/2-k1 expansion, V[i] = softmax-like weights  (:462-487)
//   rerank_expand_kernel      V[i] <- mean of V over the k2 nearest rows (query expansion)       (:489-494)
//   rerank_inv_*              inverted index of the gallery rows of V by column                  (:496-498)
//   rerank_jaccard_kernel     Jaccard distance of every query to every gallery row + the blend   (:500-513)
// Sparse rows are ELL: [N][cap] column indices (ascending) and values, with a per-row count.
#include "common.cuh"

#include <cfloat>

namespace pps {

// ------------------------------------------------------------------------------------
// OD = transpose(M^2 / colmax(M^2)): a tiled transpose with the column maxima taken in a first pass.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rerank_colmax_kernel(const float* __restrict__ m, long long ld, int n,
                                                            float* __restrict__ colmax) {
  // one thread per column, rows strided over blockIdx.y; float max of squares via atomicMax on the bits (values >= 0)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float mx = 0.f;
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    const float v = m[(long long)r * ld + c];
    mx = fmaxf(mx, __fmul_rn(v, v));
  }
  atomicMax(reinterpret_cast<int*>(colmax) + c, __float_as_int(mx));
}

__global__ void __launch_bounds__(256) rerank_normalize_kernel(const float* __restrict__ m, long long ld, int n,
                                                               const float* __restrict__ colmax, float* __restrict__ od,
                                                               long long ldo) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;          // tile of M: rows by.., cols bx..
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = by + k, c = bx + tx;
    float v = 0.f;
    if (r < n && c < n) {
      const float x = m[(long long)r * ld + c];
      v = __fdiv_rn(__fmul_rn(x, x), colmax[c]);                  // np.power(.., 2) then / np.max(axis=0): float32 ops
    }
    tile[k][tx] = v;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int orow = bx + k, ocol = by + tx;                      // OD[c][r]
    if (orow < n && ocol < n) od[(long long)orow * ldo + ocol] = tile[tx][k];
  }
}

// ------------------------------------------------------------------------------------
// k-reciprocal neighbours and V rows.  One warp per image i.  R = initial_rank [n][rk] (rk >= k1 + 1).
// ------------------------------------------------------------------------------------
constexpr int kVCap = 256;            // >= (k1 + 1) + (k1 + 1) * (k1/2 + 1) = 252 for k1 = 20

__device__ __forceinline__ void warp_bitonic_sort_256(int* a, int lane) {
  for (int k = 2; k <= kVCap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < kVCap; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(128) rerank_krecip_kernel(const int32_t* __restrict__ R, int rk, int n, int k1,
                                                            const float* __restrict__ od, long long ldo,
                                                            int32_t* __restrict__ v_idx, float* __restrict__ v_val,
                                                            int32_t* __restrict__ v_cnt) {
  __shared__ int s_exp[4][kVCap];
  __shared__ int s_kr[4][32];
  __shared__ int s_ck[4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  int* ex = s_exp[warp];
  int* kr = s_kr[warp];
  int* ck = s_ck[warp];
  const int kf = k1 + 1;                                       // forward list length            (:463)
  const int kh = (int)rintf(k1 / 2.f) + 1;                     // int(np.around(k1 / 2.)) + 1    (:471-473); rint = half-to-even
  // --- k-reciprocal neighbours of i: forward neighbours f with i among f's forward neighbours (:463-466) ---
  int f = -1;
  bool rec = false;
  if (lane < kf) {
    f = R[(long long)i * rk + lane];
    for (int t = 0; t < kf; ++t) rec |= (R[(long long)f * rk + t] == i);
  }
  unsigned bal = __ballot_sync(0xffffffffu, rec);
  const int nkr = __popc(bal);
  if (rec) kr[__popc(bal & ((1u << lane) - 1u))] = f;
  for (int t = lane; t < kVCap; t += 32) ex[t] = 0x7fffffff;
  __syncwarp();
  if (lane < nkr) ex[lane] = kr[lane];
  int nex = nkr;
  __syncwarp();
  // --- expansion by the candidates' own (k1/2)-reciprocal sets (:468-483) ---
  for (int a = 0; a < nkr; ++a) {
    const int cand = kr[a];
    int cf = -1;
    bool crec = false;
    if (lane < kh) {
      cf = R[(long long)cand * rk + lane];
      for (int t = 0; t < kh; ++t) crec |= (R[(long long)cf * rk + t] == cand);
    }
    const unsigned cb = __ballot_sync(0xffffffffu, crec);
    const int m = __popc(cb);
    if (crec) ck[__popc(cb & ((1u << lane) - 1u))] = cf;
    __syncwarp();
    bool common = false;
    if (lane < m) {
      const int x = ck[lane];
      for (int t = 0; t < nkr; ++t) common |= (kr[t] == x);
    }
    const int inter = __popc(__ballot_sync(0xffffffffu, common));
    if ((double)inter > 2.0 / 3.0 * (double)m) {               // len(intersect1d) > 2./3 * len(candidate set)
      if (lane < m) ex[nex + lane] = ck[lane];
      nex += m;
    }
    __syncwarp();
  }
  // --- np.unique: sort, drop duplicates (:485) ---
  warp_bitonic_sort_256(ex, lane);
  int out = 0;
  int32_t* oi = v_idx + (long long)i * kVCap;
  float* ov = v_val + (long long)i * kVCap;
  for (int base = 0; base < kVCap; base += 32) {
    const int x = ex[base + lane];
    const int prev = (base + lane) > 0 ? ex[base + lane - 1] : -1;
    const bool keep = x != 0x7fffffff && x != prev;
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    if (keep) oi[out + __popc(kb & ((1u << lane) - 1u))] = x;
    out += __popc(kb);
  }
  __syncwarp();
  // --- weight = exp(-OD[i, idx]); V[i, idx] = weight / sum(weight)   (:486-487) ---
  double part = 0.0;
  for (int t = lane; t < out; t += 32) {
    const float w = expf(-od[(long long)i * ldo + oi[t]]);
    ov[t] = w;
    part += (double)w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  const float tot = (float)part;
  __syncwarp();
  for (int t = lane; t < out; t += 32) ov[t] = __fdiv_rn(ov[t], tot);
  if (lane == 0) v_cnt[i] = out;
}

// ------------------------------------------------------------------------------------
// Query expansion (:489-494): V_qe[i] = mean over the k2 nearest rows of V (float32, rows added in rank order,
// then one division by k2).  One CTA per image: gather <= k2 * 256 (column, rank, value) entries, sort by
// (column, rank), sum every column's run in order.
// ------------------------------------------------------------------------------------
constexpr int kQeThreads = 128;
constexpr int kQeBuf = 2048;          // >= k2 * kVCap for k2 <= 8

__device__ void bitonic_sort_u64(unsigned long long* a, int n, int tid, int nthreads) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n; i += nthreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kQeThreads) rerank_expand_kernel(const int32_t* __restrict__ R, int rk, int n, int k2,
                                                                    const int32_t* __restrict__ v_idx,
                                                                    const float* __restrict__ v_val,
                                                                    const int32_t* __restrict__ v_cnt, int qe_cap,
                                                                    int32_t* __restrict__ q_idx, float* __restrict__ q_val,
                                                                    int32_t* __restrict__ q_cnt) {
  __shared__ unsigned long long buf[kQeBuf];       // ((column << 3 | rank) << 32) | value bits
  __shared__ int s_n, s_out;
  const int i = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { s_n = 0; s_out = 0; }
  for (int t = tid; t < kQeBuf; t += kQeThreads) buf[t] = ~0ull;
  __syncthreads();
  for (int t = 0; t < k2; ++t) {
    const int r = R[(long long)i * rk + t];
    const int c = v_cnt[r];
    const int base = s_n;                           // (uniform: read after the barrier below / above)
    for (int e = tid; e < c; e += kQeThreads) {
      const unsigned long long col = (unsigned long long)(uint32_t)v_idx[(long long)r * kVCap + e];
      buf[base + e] = (((col << 3) | (unsigned long long)t) << 32) | (unsigned long long)__float_as_uint(v_val[(long long)r * kVCap + e]);
    }
    __syncthreads();
    if (tid == 0) s_n = base + c;
    __syncthreads();
  }
  const int total = s_n;
  int npow = 64;
  while (npow < total) npow <<= 1;
  bitonic_sort_u64(buf, npow, tid, kQeThreads);
  // heads of runs of equal column: sum the run (<= k2 entries, rank order) in float32, divide by k2
  const float inv_k = (float)k2;
  for (int base = 0; base < total; base += kQeThreads) {
    const int e = base + tid;
    bool head = false;
    uint32_t col = 0;
    float sum = 0.f;
    if (e < total) {
      col = (uint32_t)(buf[e] >> 35);
      head = (e == 0) || ((uint32_t)(buf[e - 1] >> 35) != col);
      if (head) {
        for (int x = e; x < total && (uint32_t)(buf[x] >> 35) == col; ++x)
          sum = __fadd_rn(sum, __uint_as_float((uint32_t)(buf[x] & 0xffffffffull)));
      }
    }
    // ordered compaction of the heads of this slice
    const unsigned hb = __ballot_sync(0xffffffffu, head);
    __shared__ int wcnt[kQeThreads / 32];
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) wcnt[warp] = __popc(hb);
    __syncthreads();
    int before = __popc(hb & ((1u << lane) - 1u));
    int slice = 0;
    for (int w = 0; w < kQeThreads / 32; ++w) {
      if (w < warp) before += wcnt[w];
      slice += wcnt[w];
    }
    const int o0 = s_out;
    if (head && o0 + before < qe_cap) {
      q_idx[(long long)i * qe_cap + o0 + before] = (int32_t)col;
      q_val[(long long)i * qe_cap + o0 + before] = __fdiv_rn(sum, inv_k);
    }
    __syncthreads();
    if (tid == 0) s_out = o0 + slice;
    __syncthreads();
  }
  if (tid == 0) q_cnt[i] = s_out < qe_cap ? s_out : qe_cap;
}

// ------------------------------------------------------------------------------------
// Inverted index of the GALLERY rows (images nq .. n-1) by column: count -> scan -> fill.
// Entries of a column hold distinct rows, so the Jaccard kernel can update them without atomics.
// ------------------------------------------------------------------------------------
__global__ void rerank_inv_count_kernel(const int32_t* __restrict__ q_idx, const int32_t* __restrict__ q_cnt, int qe_cap,
                                        int row0, int n, int32_t* __restrict__ col_cnt) {
  const int r = row0 + blockIdx.x;
  if (r >= n) return;
  const int c = q_cnt[r];
  for (int e = threadIdx.x; e < c; e += blockDim.x) atomicAdd(&col_cnt[q_idx[(long long)r * qe_cap + e]], 1);
}

__global__ void __launch_bounds__(1024) rerank_inv_scan_kernel(const int32_t* __restrict__ col_cnt, int n,
                                                               int32_t* __restrict__ col_off, int32_t* __restrict__ cursor) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int v = i < n ? col_cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < n) {
      const int off = carry + (warp ? warp_sum[warp - 1] : 0) + incl - v;
      col_off[i] = off;
      cursor[i] = off;
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) col_off[n] = carry_s;
}

__global__ void rerank_inv_fill_kernel(const int32_t* __restrict__ q_idx, const float* __restrict__ q_val,
                                       const int32_t* __restrict__ q_cnt, int qe_cap, int row0, int n,
                                       int32_t* __restrict__ cursor, int32_t* __restrict__ inv_row,
                                       float* __restrict__ inv_val) {
  const int r = row0 + blockIdx.x;
  if (r >= n) return;
  const int c = q_cnt[r];
  for (int e = threadIdx.x; e < c; e += blockDim.x) {
    const int col = q_idx[(long long)r * qe_cap + e];
    const int p = atomicAdd(&cursor[col], 1);
    inv_row[p] = r - row0;                       // gallery-relative row
    inv_val[p] = q_val[(long long)r * qe_cap + e];
  }
}

// ------------------------------------------------------------------------------------
// Jaccard distance + blend (:500-513).  One CTA per query i; its non-zero columns are visited in ascending order
// (as the reference does) and every column's inverted list updates temp_min[row] in parallel: rows inside a list
// are distinct, lists are separated by a barrier, so the float32 sums are formed in the reference's order with no
// atomics.  final[i][g] = (1 - temp/(2 - temp)) * (1 - lambda) + OD[i][nq + g] * lambda.
// temp_min lives in shared memory (ng floats) or, for galleries beyond that, in the output row itself.
// ------------------------------------------------------------------------------------
constexpr int kJacThreads = 256;

__global__ void __launch_bounds__(kJacThreads) rerank_jaccard_kernel(const int32_t* __restrict__ q_idx,
                                                                      const float* __restrict__ q_val,
                                                                      const int32_t* __restrict__ q_cnt, int qe_cap,
                                                                      const int32_t* __restrict__ col_off,
                                                                      const int32_t* __restrict__ inv_row,
                                                                      const float* __restrict__ inv_val, int nq, int ng,
                                                                      const float* __restrict__ od, long long ldo,
                                                                      float lambda_value, int use_smem,
                                                                      float* __restrict__ out, long long ld_out) {
  extern __shared__ float s_temp[];
  const int i = blockIdx.x, tid = threadIdx.x;
  float* temp = use_smem ? s_temp : out + (long long)i * ld_out;
  for (int g = tid; g < ng; g += kJacThreads) temp[g] = 0.f;
  __syncthreads();
  const int c = q_cnt[i];
  for (int e = 0; e < c; ++e) {
    const int col = q_idx[(long long)i * qe_cap + e];
    const float vi = q_val[(long long)i * qe_cap + e];
    const int p0 = col_off[col], p1 = col_off[col + 1];
    for (int p = p0 + tid; p < p1; p += kJacThreads) {
      const int g = inv_row[p];
      temp[g] = __fadd_rn(temp[g], fminf(vi, inv_val[p]));
    }
    __syncthreads();
  }
  const float one_minus = 1.0f - lambda_value;   // the reference's Python floats become float32 when they meet the arrays
  for (int g = tid; g < ng; g += kJacThreads) {
    const float t = temp[g];
    const float jac = __fsub_rn(1.0f, __fdiv_rn(t, __fsub_rn(2.0f, t)));
    const float o = od[(long long)i * ldo + nq + g];
    out[(long long)i * ld_out + g] = __fadd_rn(__fmul_rn(jac, one_minus), __fmul_rn(o, lambda_value));
  }
}

}  // namespace pps

using namespace pps;

extern "C" int pps_rerank_vcap(void) { return kVCap; }

extern "C" int pps_rerank_normalize(const float* m, long long ld, long long n, float* colmax, float* od, long long ldo,
                                    void* stream) {
  if (n < 0 || ld < n || ldo < n) return PPS_ERR_INVALID_ARG;
  if (n == 0) return PPS_OK;
  if (!m || !colmax || !od) return PPS_ERR_INVALID_ARG;
  if (n > 0x7fffffffLL / 2) return PPS_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PPS_CUDA_TRY(cudaMemsetAsync(colmax, 0, (size_t)n * 4, st));
  const int ysplit = (int)(n < 64 ? 1 : 64);
  rerank_colmax_kernel<<<dim3((unsigned)((n + 255) / 256), ysplit), 256, 0, st>>>(m, ld, (int)n, colmax);
  PPS_LAUNCH_CHECK("rerank_colmax_kernel");
  const unsigned t = (unsigned)((n + 31) / 32);
  rerank_normalize_kernel<<<dim3(t, t), 256, 0, st>>>(m, ld, (int)n, colmax, od, ldo);
  PPS_LAUNCH_CHECK("rerank_normalize_kernel");
  return PPS_OK;
}

extern "C" int pps_rerank_krecip(const int32_t* initial_rank, int rank_cols, long long n, int k1, const float* od,
                                 long long ldo, int32_t* v_idx, float* v_val, int32_t* v_cnt, void* stream) {
  if (n < 0 || k1 < 1 || rank_cols < k1 + 1 || ldo < n) return PPS_ERR_INVALID_ARG;
  if (k1 + 1 > 32 || (k1 + 1) + (k1 + 1) * ((int)rintf(k1 / 2.f) + 1) > kVCap) return PPS_ERR_UNSUPPORTED;
  if (n == 0) return PPS_OK;
  if (!initial_rank || !od || !v_idx || !v_val || !v_cnt) return PPS_ERR_INVALID_ARG;
  rerank_krecip_kernel<<<(unsigned)((n + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      initial_rank, rank_cols, (int)n, k1, od, ldo, v_idx, v_val, v_cnt);
  PPS_LAUNCH_CHECK("rerank_krecip_kernel");
  return PPS_OK;
}

extern "C" int pps_rerank_expand(const int32_t* initial_rank, int rank_cols, long long n, int k2, const int32_t* v_idx,
                                 const float* v_val, const int32_t* v_cnt, int qe_cap, int32_t* q_idx, float* q_val,
                                 int32_t* q_cnt, void* stream) {
  if (n < 0 || k2 < 1 || k2 > 8 || rank_cols < k2 || qe_cap < 1) return PPS_ERR_INVALID_ARG;
  if (k2 * kVCap > kQeBuf) return PPS_ERR_UNSUPPORTED;
  if (n == 0) return PPS_OK;
  if (!initial_rank || !v_idx || !v_val || !v_cnt || !q_idx || !q_val || !q_cnt) return PPS_ERR_INVALID_ARG;
  rerank_expand_kernel<<<(unsigned)n, kQeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      initial_rank, rank_cols, (int)n, k2, v_idx, v_val, v_cnt, qe_cap, q_idx, q_val, q_cnt);
  PPS_LAUNCH_CHECK("rerank_expand_kernel");
  return PPS_OK;
}

// col_cnt / cursor: [n] scratch; col_off: [n + 1]; inv_row / inv_val: capacity >= total non-zeros of the gallery rows
// (<= ng * qe_cap).  Images nq .. n-1 are the gallery.
extern "C" int pps_rerank_invert(const int32_t* q_idx, const float* q_val, const int32_t* q_cnt, int qe_cap, long long nq,
                                 long long n, int32_t* col_cnt, int32_t* col_off, int32_t* cursor, int32_t* inv_row,
                                 float* inv_val, void* stream) {
  if (n < 0 || nq < 0 || nq > n || qe_cap < 1) return PPS_ERR_INVALID_ARG;
  if (n == 0) return PPS_OK;
  if (!q_idx || !q_val || !q_cnt || !col_cnt || !col_off || !cursor || !inv_row || !inv_val) return PPS_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PPS_CUDA_TRY(cudaMemsetAsync(col_cnt, 0, (size_t)n * 4, st));
  const long long ng = n - nq;
  if (ng > 0) {
    rerank_inv_count_kernel<<<(unsigned)ng, 128, 0, st>>>(q_idx, q_cnt, qe_cap, (int)nq, (int)n, col_cnt);
    PPS_LAUNCH_CHECK("rerank_inv_count_kernel");
  }
  rerank_inv_scan_kernel<<<1, 1024, 0, st>>>(col_cnt, (int)n, col_off, cursor);
  PPS_LAUNCH_CHECK("rerank_inv_scan_kernel");
  if (ng > 0) {
    rerank_inv_fill_kernel<<<(unsigned)ng, 128, 0, st>>>(q_idx, q_val, q_cnt, qe_cap, (int)nq, (int)n, cursor, inv_row,
                                                        inv_val);
    PPS_LAUNCH_CHECK("rerank_inv_fill_kernel");
  }
  return PPS_OK;
}

extern "C" int pps_rerank_jaccard(const int32_t* q_idx, const float* q_val, const int32_t* q_cnt, int qe_cap,
                                  const int32_t* col_off, const int32_t* inv_row, const float* inv_val, long long nq,
                                  long long ng, const float* od, long long ldo, float lambda_value, float* out,
                                  long long ld_out, void* stream) {
  if (nq < 0 || ng < 0 || ldo < nq + ng || ld_out < ng || qe_cap < 1) return PPS_ERR_INVALID_ARG;
  if (nq == 0 || ng == 0) return PPS_OK;
  if (!q_idx || !q_val || !q_cnt || !col_off || !inv_row || !inv_val || !od || !out) return PPS_ERR_INVALID_ARG;
  const size_t smem = (size_t)ng * 4;
  const int use_smem = smem <= 200 * 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  PPS_CUDA_TRY(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    PPS_CUDA_TRY(cudaFuncSetAttribute(rerank_jaccard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dev = dev;
  }
  rerank_jaccard_kernel<<<(unsigned)nq, kJacThreads, use_smem ? smem : 0, static_cast<cudaStream_t>(stream)>>>(
      q_idx, q_val, q_cnt, qe_cap, col_off, inv_row, inv_val, (int)nq, (int)ng, od, ldo, lambda_value, use_smem, out, ld_out);
  PPS_LAUNCH_CHECK("rerank_jaccard_kernel");
  return PPS_OK;
}

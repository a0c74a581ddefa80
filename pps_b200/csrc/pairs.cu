// Same-id pair lists built on the device (the junk mask and the matches of
// reid_dataset_evaluator.py:320,327-328,421,427-428 without the [m1, m2] boolean matrices).
//
// For query i the gallery items g with gallery_ids[g] == query_ids[i] are its pairs: positives
// (camera differs) and junk (same camera).  Output is the same CSR the host builder
// (pps_pairs_fill, c_api.cu) produces — pair_off[nq+1], pair_q[E], pair_g[E] ascending inside a
// query, pair_pos[E] — so every rank of a sharded run builds identical lists and the later
// all-reduces line up element by element.
//
// Brute force on purpose: nq x ng 64-bit compares are ~2000x cheaper than the nq x ng x D
// contraction they accompany, need no sort and no hash table, and the order is deterministic:
//   count : CTA (query block of 32, gallery segment) -> 8 warps x 4 queries; a warp streams the
//           segment's ids once (coalesced), 4 ballots per 32 ids, popc-accumulates per query.
//   offsets: per-query totals of this gallery block -> (sharded: all-gather over ranks ->) one CTA
//           scans over queries -> pair_off, this block's first slot per query, {n_pairs, max/query}.
//   fill  : same sweep as count; slot = my_base[q] + hits in earlier segments + ballot prefix
//           -> ascending g.  It also zero-fills the per-pair accumulators of the rank sweeps.
// A gallery sharded over ranks sweeps only its own block (nq x ng_local compares per rank); the
// blocks' lists interleave into one global CSR because blocks are contiguous row ranges.
#include "common.cuh"

namespace pps {

constexpr int kPairQPerWarp = 4;
constexpr int kPairWarps = 8;
constexpr int kPairQPerCta = kPairQPerWarp * kPairWarps;   // 32

template <bool FILL>
__global__ void __launch_bounds__(32 * kPairWarps)
pairs_sweep_kernel(const int64_t* __restrict__ qid, const int64_t* __restrict__ qcam, int nq,
                   const int64_t* __restrict__ gid, const int64_t* __restrict__ gcam, long long ng, long long seg,
                   int nseg, int32_t* __restrict__ seg_cnt /*[nq][nseg] hits of query q in segment s*/,
                   const int32_t* __restrict__ my_base /*[nq] first global slot of this block's hits of query q*/,
                   long long g_offset /*global index of local gallery row 0*/, int32_t* __restrict__ pair_q,
                   int32_t* __restrict__ pair_g, uint8_t* __restrict__ pair_pos, int32_t* __restrict__ pair_pos32,
                   float* __restrict__ zero_f32, uint32_t* __restrict__ zero_u32,
                   uint32_t* __restrict__ zero_per_query, long long capacity,
                   const int32_t* __restrict__ ng_dev /*optional: rows actually present (<= ng), read on the device*/) {
  if (ng_dev) ng = min(ng, (long long)*ng_dev);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * kPairQPerCta + warp * kPairQPerWarp;
  const int s = blockIdx.y;
  if (q0 >= nq) return;
  int64_t ids[kPairQPerWarp], cams[kPairQPerWarp];
  int acc[kPairQPerWarp];
  bool live[kPairQPerWarp];
#pragma unroll
  for (int u = 0; u < kPairQPerWarp; ++u) {
    live[u] = q0 + u < nq;
    ids[u] = live[u] ? qid[q0 + u] : 0;
    cams[u] = (FILL && live[u]) ? qcam[q0 + u] : 0;
    acc[u] = 0;
    if (FILL && live[u]) {
      // first slot of (query, segment) = my_base[q] + hits of q in the earlier segments of this block
      int part = 0;
      for (int sp = lane; sp < s; sp += 32) part += seg_cnt[(long long)(q0 + u) * nseg + sp];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      acc[u] = my_base[q0 + u] + part;
      if (zero_per_query && s == 0 && lane == 0) zero_per_query[q0 + u] = 0u;
    }
  }
  const long long g_begin = (long long)s * seg;
  const long long g_end = min(ng, g_begin + seg);
  const unsigned lt = (1u << lane) - 1u;
  constexpr int kB = 8;                                     // batches of 32 ids in flight per warp
  for (long long g0 = g_begin; g0 < g_end; g0 += 32 * kB) {
    int64_t v[kB];
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      const long long g = g0 + k * 32 + lane;
      v[k] = g < g_end ? gid[g] : 0;
    }
#pragma unroll
    for (int k = 0; k < kB; ++k) {
      const long long g = g0 + k * 32 + lane;
      const bool in = g < g_end;
      // fast path: matches are rare (a few dozen per query in the whole gallery), so first ask whether ANY lane
      // matches ANY of the warp's queries on the low 32 bits; only then do the exact 64-bit compares and ballots
      const uint32_t vlo = (uint32_t)v[k];
      bool maybe = false;
#pragma unroll
      for (int u = 0; u < kPairQPerWarp; ++u) maybe |= (vlo == (uint32_t)ids[u]);
      if (!__any_sync(0xffffffffu, maybe && in)) continue;
#pragma unroll
      for (int u = 0; u < kPairQPerWarp; ++u) {
        const bool hit = in && live[u] && v[k] == ids[u];
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (FILL) {
          if (hit) {
            const long long slot = (long long)acc[u] + __popc(b & lt);
            if (slot < capacity) {
              pair_q[slot] = q0 + u;
              pair_g[slot] = (int32_t)(g + g_offset);
              const int pos = gcam[g] != cams[u] ? 1 : 0;
              if (pair_pos) pair_pos[slot] = (uint8_t)pos;
              if (pair_pos32) pair_pos32[slot] = pos;
              if (zero_f32) zero_f32[slot] = 0.f;
              if (zero_u32) zero_u32[slot] = 0u;
            }
          }
        }
        acc[u] += __popc(b);
      }
    }
  }
  if (!FILL && lane == 0) {
#pragma unroll
    for (int u = 0; u < kPairQPerWarp; ++u)
      if (live[u]) seg_cnt[(long long)(q0 + u) * nseg + s] = acc[u];
  }
}

// local_cnt[q] = hits of query q in this gallery block (sum over its segments)
__global__ void pairs_rowsum_kernel(const int32_t* __restrict__ seg_cnt, int nq, int nseg, int32_t* __restrict__ local_cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int32_t* row = seg_cnt + (long long)q * nseg;
  int sum = 0;
  for (int sgi = 0; sgi < nseg; ++sgi) sum += row[sgi];
  local_cnt[q] = sum;
}

// Global CSR offsets from the per-block counts cnt_all[world][nq] (world = 1: the local counts):
//   pair_off[q] = sum over q' < q and all blocks;  my_base[q] = pair_off[q] + sum over blocks r' < rank of cnt_all[r'][q]
//   (blocks are contiguous gallery ranges in rank order, so block-major order inside a query is ascending g);
//   totals = {n_pairs, max pairs of one query}.  One CTA of 1024 threads, 1024 queries per iteration.
__global__ void __launch_bounds__(1024) pairs_offsets_kernel(const int32_t* __restrict__ cnt_all, int world, int rank,
                                                             int nq, int32_t* __restrict__ pair_off,
                                                             int32_t* __restrict__ my_base, int32_t* __restrict__ totals) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  __shared__ int max_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_s = 0; max_s = 0; }
  __syncthreads();
  int mx = 0;
  for (int base = 0; base < nq; base += 1024) {
    const int q = base + tid;
    int sum = 0, before = 0;
    if (q < nq) {
      for (int r = 0; r < world; ++r) {
        const int v = cnt_all[(long long)r * nq + q];
        sum += v;
        before += r < rank ? v : 0;
      }
    }
    mx = max(mx, sum);
    int incl = sum;
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
      warp_sum[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    if (q < nq) {
      const int off = carry + (warp ? warp_sum[warp - 1] : 0) + incl - sum;
      pair_off[q] = off;
      my_base[q] = off + before;
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) atomicMax(&max_s, mx);
  __syncthreads();
  if (tid == 0) { pair_off[nq] = carry_s; totals[0] = carry_s; totals[1] = max_s; }
}

__global__ void pairs_unpack_pos_kernel(const int32_t* __restrict__ pos32, long long n, uint8_t* __restrict__ pos8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos8[i] = (uint8_t)(pos32[i] != 0);
}

static void pair_segments(long long ng, long long* seg, int* nseg) {
  long long s = 1024;
  while ((ng + s - 1) / s > 512) s *= 2;
  *seg = s;
  *nseg = (int)((ng + s - 1) / s);
  if (*nseg < 1) *nseg = 1;
}

// ------------------------------------------------------------------------------------
// Compacted same-id gallery ("G'") for the threshold pass of galleries that do not fit one
// distance block.  The thresholds of the ranking are the distances of the same-id pairs only, so
// they come from a dense product of the queries with JUST the gallery rows that appear in some pair
// (a few 10^4 rows however large the gallery is) instead of a first sweep over the whole gallery.
// All queries of one id share the same gallery list, so G' is the concatenation of the lists of
// one representative query per id (the first query carrying it), restricted to the row window
// [row_lo, row_hi) of this gallery shard; pair e of query q then sits at column
// gp_off[rep(q)] + (e - pair_off[q]) - lo(q) of the product.
// ------------------------------------------------------------------------------------
// The representative of a query = the lowest query index carrying the same id.  All queries of one id share one gallery
// list, so the id is identified by the list's first gallery row g0: an open-addressing table keyed by g0 (claimed with one
// atomicCAS on the key itself) collects the minimum query index per id - O(nq) instead of the O(nq^2) scan over earlier
// queries this kernel started with (270 us at 3 368 queries, ncu launch list profiles/r02_e_launches.md).
__global__ void compact_hash_insert_kernel(int nq, const int32_t* __restrict__ pair_off, const int32_t* __restrict__ pair_g,
                                           int32_t* __restrict__ keys, int32_t* __restrict__ vals, uint32_t mask) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int e0 = pair_off[q];
  if (pair_off[q + 1] == e0) return;                      // no pair at all: the query is its own representative
  const int32_t g0 = pair_g[e0];
  uint32_t s = ((uint32_t)g0 * 0x9E3779B1u) & mask;
  while (true) {
    const int32_t old = atomicCAS(&keys[s], -1, g0);
    if (old == -1 || old == g0) { atomicMin(&vals[s], q); return; }
    s = (s + 1) & mask;
  }
}

__global__ void compact_rep_kernel(const int64_t* __restrict__ qid, int nq, const int32_t* __restrict__ pair_off,
                                   const int32_t* __restrict__ pair_g, int row_lo, int row_hi,
                                   const int32_t* __restrict__ keys, const int32_t* __restrict__ vals, uint32_t mask,
                                   int32_t* __restrict__ rep, int32_t* __restrict__ lo_out, int32_t* __restrict__ cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  int r = q;
  if (pair_off[q + 1] > pair_off[q]) {
    const int32_t g0 = pair_g[pair_off[q]];
    uint32_t s = ((uint32_t)g0 * 0x9E3779B1u) & mask;
    while (keys[s] != g0) s = (s + 1) & mask;             // present: inserted by compact_hash_insert_kernel
    r = vals[s];
  }
  rep[q] = r;
  (void)qid;
  // the pair list is ascending in the gallery index: the rows of the window are one sub-range
  const int e0 = pair_off[q], e1 = pair_off[q + 1];
  int a = e0, b = e1;
  while (a < b) { const int m = (a + b) >> 1; if (pair_g[m] < row_lo) a = m + 1; else b = m; }
  const int lo = a;
  b = e1;
  while (a < b) { const int m = (a + b) >> 1; if (pair_g[m] < row_hi) a = m + 1; else b = m; }
  lo_out[q] = lo - e0;
  cnt[q] = a - lo;              // rows of q's list inside the window (whether or not q is the representative)
}

// one CTA: exclusive scan of the representatives' counts -> gp_off[q] (meaningful where rep[q] == q)
__global__ void __launch_bounds__(1024) compact_scan_kernel(const int32_t* __restrict__ rep, const int32_t* __restrict__ cnt,
                                                            int nq, int32_t* __restrict__ gp_off,
                                                            int32_t* __restrict__ n_rows) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nq; base += 1024) {
    const int q = base + tid;
    const int v = (q < nq && rep[q] == q) ? cnt[q] : 0;
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
    if (q < nq) gp_off[q] = carry + (warp ? warp_sum[warp - 1] : 0) + incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) *n_rows = carry_s;
}

__global__ void compact_fill_kernel(const int32_t* __restrict__ pair_off, const int32_t* __restrict__ pair_q,
                                    const int32_t* __restrict__ pair_g, long long n_pairs,
                                    const int32_t* __restrict__ rep, const int32_t* __restrict__ lo,
                                    const int32_t* __restrict__ cnt, const int32_t* __restrict__ gp_off,
                                    int32_t* __restrict__ gp_rows, int32_t* __restrict__ pair_col,
                                    const int32_t* __restrict__ n_dev) {
  if (n_dev) n_pairs = min(n_pairs, (long long)*n_dev);
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pairs) return;
  const int q = pair_q[e];
  int col = -1;
  if (q >= 0) {
    const int i = (int)(e - pair_off[q]) - lo[q];
    if (i >= 0 && i < cnt[q]) {
      const int r = rep[q];
      col = gp_off[r] + i;
      if (r == q) gp_rows[col] = pair_g[e];
    }
  }
  pair_col[e] = col;
}

// ------------------------------------------------------------------------------------
// Candidate pre-filter for very large galleries.  The pair sweeps compare every query id with every gallery id; with
// millions of distractor rows almost all of that is wasted, so first keep only the gallery rows whose id is the id
// of SOME query: a hash set of the query ids (open addressing, built once), one membership probe per gallery row,
// and an ORDERED compaction (block counts -> scan -> fill) so that the candidate list stays ascending in the
// gallery index.  The sweeps then run on the compact (id, camera) arrays and pps_pairs_remap turns their pair_g
// (an index into the candidate list) back into gallery rows.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_id(int64_t id, uint32_t mask) {
  uint64_t x = (uint64_t)id * 0x9E3779B97F4A7C15ull;
  return (uint32_t)(x >> 32) & mask;
}

__global__ void qset_build_kernel(const int64_t* __restrict__ qid, int nq, int64_t* __restrict__ keys,
                                  int32_t* __restrict__ occ, uint32_t mask) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int64_t id = qid[q];
  uint32_t s = hash_id(id, mask);
  while (true) {
    const int old = atomicCAS(&occ[s], 0, 1);
    if (old == 0) { keys[s] = id; return; }      // claimed an empty slot (duplicates of an id may take several: harmless)
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ bool qset_has(int64_t id, const int64_t* __restrict__ keys, const int32_t* __restrict__ occ,
                                         uint32_t mask) {
  uint32_t s = hash_id(id, mask);
  while (occ[s]) {
    if (keys[s] == id) return true;
    s = (s + 1) & mask;
  }
  return false;
}

constexpr int kFilterRows = 2048;      // gallery rows per CTA (256 threads x 8)

template <bool FILL>
__global__ void __launch_bounds__(256) prefilter_kernel(const int64_t* __restrict__ gid, const int64_t* __restrict__ gcam,
                                                        long long ng, const int64_t* __restrict__ keys,
                                                        const int32_t* __restrict__ occ, uint32_t mask,
                                                        int32_t* __restrict__ blk_cnt /*count: out; fill: exclusive offsets*/,
                                                        int32_t* __restrict__ cand_rows, int64_t* __restrict__ cand_gid,
                                                        int64_t* __restrict__ cand_gcam) {
  __shared__ int warp_cnt[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row0 = (long long)blockIdx.x * kFilterRows;
  // thread t owns 8 CONSECUTIVE rows, so ranks inside the block follow the gallery order
  const long long r0 = row0 + (long long)tid * 8;
  int64_t id[8];
  unsigned hit = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    id[e] = 0;
    if (r0 + e < ng) {
      id[e] = gid[r0 + e];
      if (qset_has(id[e], keys, occ, mask)) hit |= 1u << e;
    }
  }
  const int mine = __popc(hit);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_cnt[warp] = incl;
  __syncthreads();
  int before = incl - mine, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < warp) before += warp_cnt[w];
    total += warp_cnt[w];
  }
  if (!FILL) {
    if (tid == 0) blk_cnt[blockIdx.x] = total;
    return;
  }
  int pos = blk_cnt[blockIdx.x] + before;
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (hit & (1u << e)) {
      cand_rows[pos] = (int32_t)(r0 + e);
      cand_gid[pos] = id[e];
      cand_gcam[pos] = gcam[r0 + e];
      ++pos;
    }
}

// one CTA: in-place exclusive scan of n block counts, total -> *n_out
__global__ void __launch_bounds__(1024) prefilter_scan_kernel(int32_t* __restrict__ cnt, int n, int32_t* __restrict__ n_out) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int v = i < n ? cnt[i] : 0;
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
    if (i < n) cnt[i] = carry + (warp ? warp_sum[warp - 1] : 0) + incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) *n_out = carry_s;
}

__global__ void pairs_remap_kernel(int32_t* __restrict__ pair_g, long long n, const int32_t* __restrict__ cand_rows,
                                   long long offset, const int32_t* __restrict__ n_dev) {
  if (n_dev) n = min(n, (long long)*n_dev);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pair_g[i] = (int32_t)(cand_rows[pair_g[i]] + offset);
}

// rows [*n_rows, cap) of the compacted row list := a valid row (the product over `cap` columns then reads real memory; the
// columns beyond *n_rows are never gathered)
__global__ void compact_pad_kernel(int32_t* __restrict__ gp_rows, const int32_t* __restrict__ n_rows, long long cap, int32_t pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap && i >= (long long)*n_rows) gp_rows[i] = pad;
}

}  // namespace pps

using namespace pps;

// workspace: [nq][nseg] segment counts | [nq] local counts | [nq] my_base
static long long ws_ints(long long nq, long long ng, int* nseg_out) {
  long long seg; int nseg;
  pair_segments(ng, &seg, &nseg);
  if (nseg_out) *nseg_out = nseg;
  return (nq > 0 ? nq : 1) * ((long long)nseg + 2);
}

extern "C" long long pps_pairs_workspace_bytes(long long nq, long long ng) {
  if (nq < 0 || ng < 0) return PPS_ERR_INVALID_ARG;
  return ws_ints(nq, ng, nullptr) * 4;
}

namespace pps {
int pairs_local_count_ex(const int64_t* query_ids, long long nq, const int64_t* gallery_ids, long long ng, void* workspace,
                         int32_t** local_cnt, void* stream, const int32_t* ng_dev) {
  if (nq < 0 || ng < 0 || nq > 0x7fffffffLL || ng > 0x7fffffffLL) return PPS_ERR_INVALID_ARG;
  if (!workspace) return PPS_ERR_INVALID_ARG;
  if (nq > 0 && ng > 0 && (!query_ids || !gallery_ids)) return PPS_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long seg; int nseg;
  pair_segments(ng, &seg, &nseg);
  int32_t* seg_cnt = static_cast<int32_t*>(workspace);
  int32_t* lc = seg_cnt + (nq > 0 ? nq : 1) * (long long)nseg;
  if (local_cnt) *local_cnt = lc;
  if ((long long)nq * nseg > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
  if (nq == 0) return PPS_OK;
  if (ng == 0) {
    PPS_CUDA_TRY(cudaMemsetAsync(seg_cnt, 0, (size_t)nq * (nseg + 1) * 4, st));
    return PPS_OK;
  }
  const dim3 grid((unsigned)((nq + kPairQPerCta - 1) / kPairQPerCta), (unsigned)nseg);
  pairs_sweep_kernel<false><<<grid, 32 * kPairWarps, 0, st>>>(query_ids, nullptr, (int)nq, gallery_ids, nullptr, ng, seg,
                                                              nseg, seg_cnt, nullptr, 0, nullptr, nullptr, nullptr,
                                                              nullptr, nullptr, nullptr, nullptr, 0, ng_dev);
  PPS_LAUNCH_CHECK("pairs_sweep_kernel<count>");
  pairs_rowsum_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(seg_cnt, (int)nq, nseg, lc);
  PPS_LAUNCH_CHECK("pairs_rowsum_kernel");
  return PPS_OK;
}
}  // namespace pps

extern "C" int pps_pairs_local_count(const int64_t* query_ids, long long nq, const int64_t* gallery_ids, long long ng,
                                     void* workspace, int32_t** local_cnt, void* stream) {
  return pairs_local_count_ex(query_ids, nq, gallery_ids, ng, workspace, local_cnt, stream, nullptr);
}

extern "C" int pps_pairs_offsets(const int32_t* cnt_all, int world, int rank, long long nq, long long ng_local,
                                 void* workspace, int32_t* pair_off, int32_t* totals, void* stream) {
  if (nq < 0 || world < 1 || rank < 0 || rank >= world || nq > 0x7fffffffLL) return PPS_ERR_INVALID_ARG;
  if (!cnt_all || !workspace || !pair_off || !totals) return PPS_ERR_INVALID_ARG;
  int nseg;
  ws_ints(nq, ng_local, &nseg);
  int32_t* my_base = static_cast<int32_t*>(workspace) + (nq > 0 ? nq : 1) * ((long long)nseg + 1);
  pairs_offsets_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(cnt_all, world, rank, (int)nq, pair_off, my_base,
                                                                          totals);
  PPS_LAUNCH_CHECK("pairs_offsets_kernel");
  return PPS_OK;
}

namespace pps {
int pairs_fill_local_ex(const int64_t* query_ids, const int64_t* query_cams, long long nq, const int64_t* gallery_ids,
                        const int64_t* gallery_cams, long long ng, long long gallery_offset, const void* workspace,
                        int32_t* pair_q, int32_t* pair_g, uint8_t* pair_pos, int32_t* pair_pos32, float* zero_f32,
                        uint32_t* zero_u32, uint32_t* zero_per_query, long long capacity, void* stream, const int32_t* ng_dev) {
  if (nq < 0 || ng < 0 || nq > 0x7fffffffLL || ng > 0x7fffffffLL || capacity < 0 || gallery_offset < 0 ||
      gallery_offset + ng > 0x7fffffffLL)
    return PPS_ERR_INVALID_ARG;
  if (nq == 0) return PPS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ng == 0 || capacity == 0) {     // no local pairs: only the per-query counters need their zero
    if (zero_per_query) PPS_CUDA_TRY(cudaMemsetAsync(zero_per_query, 0, (size_t)nq * 4, st));
    return PPS_OK;
  }
  if (!query_ids || !query_cams || !gallery_ids || !gallery_cams || !workspace || !pair_q || !pair_g ||
      (!pair_pos && !pair_pos32))
    return PPS_ERR_INVALID_ARG;
  long long seg; int nseg;
  pair_segments(ng, &seg, &nseg);
  int32_t* seg_cnt = const_cast<int32_t*>(static_cast<const int32_t*>(workspace));
  const int32_t* my_base = seg_cnt + nq * ((long long)nseg + 1);
  const dim3 grid((unsigned)((nq + kPairQPerCta - 1) / kPairQPerCta), (unsigned)nseg);
  pairs_sweep_kernel<true><<<grid, 32 * kPairWarps, 0, st>>>(query_ids, query_cams, (int)nq, gallery_ids, gallery_cams, ng,
                                                             seg, nseg, seg_cnt, my_base, gallery_offset, pair_q, pair_g,
                                                             pair_pos, pair_pos32, zero_f32, zero_u32, zero_per_query,
                                                             capacity, ng_dev);
  PPS_LAUNCH_CHECK("pairs_sweep_kernel<fill>");
  return PPS_OK;
}

// the device-count forms used by the speculative (sync-free) pass: sizes are host upper bounds, the actual counts are read
// on the device
int pairs_remap_ex(int32_t* pair_g, long long n_cap, const int32_t* cand_rows, long long offset, void* stream,
                   const int32_t* n_dev) {
  if (n_cap <= 0) return PPS_OK;
  pairs_remap_kernel<<<(unsigned)((n_cap + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pair_g, n_cap, cand_rows,
                                                                                                     offset, n_dev);
  PPS_LAUNCH_CHECK("pairs_remap_kernel");
  return PPS_OK;
}
}  // namespace pps

extern "C" int pps_pairs_fill_local(const int64_t* query_ids, const int64_t* query_cams, long long nq,
                                    const int64_t* gallery_ids, const int64_t* gallery_cams, long long ng,
                                    long long gallery_offset, const void* workspace, int32_t* pair_q, int32_t* pair_g,
                                    uint8_t* pair_pos, int32_t* pair_pos32, float* zero_f32, uint32_t* zero_u32,
                                    uint32_t* zero_per_query, long long capacity, void* stream) {
  return pairs_fill_local_ex(query_ids, query_cams, nq, gallery_ids, gallery_cams, ng, gallery_offset, workspace, pair_q, pair_g,
                             pair_pos, pair_pos32, zero_f32, zero_u32, zero_per_query, capacity, stream, nullptr);
}

extern "C" int pps_pairs_unpack_pos(const int32_t* pair_pos32, long long n_pairs, uint8_t* pair_pos, void* stream) {
  if (n_pairs < 0) return PPS_ERR_INVALID_ARG;
  if (n_pairs == 0) return PPS_OK;
  if (!pair_pos32 || !pair_pos) return PPS_ERR_INVALID_ARG;
  pairs_unpack_pos_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pair_pos32, n_pairs, pair_pos);
  PPS_LAUNCH_CHECK("pairs_unpack_pos_kernel");
  return PPS_OK;
}

namespace pps {
int pairs_count_device_ex(const int64_t* query_ids, long long nq, const int64_t* gallery_ids, long long ng_cap, void* workspace,
                          int32_t* pair_off, int32_t* totals, void* stream, const int32_t* ng_dev) {
  if (!pair_off || !totals) return PPS_ERR_INVALID_ARG;
  int32_t* lc = nullptr;
  int rc = pairs_local_count_ex(query_ids, nq, gallery_ids, ng_cap, workspace, &lc, stream, ng_dev);
  if (rc != PPS_OK) return rc;
  return pps_pairs_offsets(lc, 1, 0, nq, ng_cap, workspace, pair_off, totals, stream);
}
int pairs_fill_device_ex(const int64_t* query_ids, const int64_t* query_cams, long long nq, const int64_t* gallery_ids,
                         const int64_t* gallery_cams, long long ng_cap, const void* workspace, int32_t* pair_q, int32_t* pair_g,
                         uint8_t* pair_pos, float* zero_f32, uint32_t* zero_u32, uint32_t* zero_per_query, long long capacity,
                         void* stream, const int32_t* ng_dev) {
  return pairs_fill_local_ex(query_ids, query_cams, nq, gallery_ids, gallery_cams, ng_cap, 0, workspace, pair_q, pair_g,
                             pair_pos, nullptr, zero_f32, zero_u32, zero_per_query, capacity, stream, ng_dev);
}
}  // namespace pps

// ---- single-block convenience forms (whole gallery on one device) ----
extern "C" int pps_pairs_count_device(const int64_t* query_ids, long long nq, const int64_t* gallery_ids, long long ng,
                                      void* workspace, int32_t* pair_off, int32_t* totals, void* stream) {
  if (!pair_off || !totals) return PPS_ERR_INVALID_ARG;
  int32_t* lc = nullptr;
  int rc = pps_pairs_local_count(query_ids, nq, gallery_ids, ng, workspace, &lc, stream);
  if (rc != PPS_OK) return rc;
  return pps_pairs_offsets(lc, 1, 0, nq, ng, workspace, pair_off, totals, stream);
}

extern "C" int pps_pairs_fill_device(const int64_t* query_ids, const int64_t* query_cams, long long nq,
                                     const int64_t* gallery_ids, const int64_t* gallery_cams, long long ng,
                                     const void* workspace, int32_t* pair_q, int32_t* pair_g, uint8_t* pair_pos,
                                     float* zero_f32, uint32_t* zero_u32, uint32_t* zero_per_query, long long capacity,
                                     void* stream) {
  return pps_pairs_fill_local(query_ids, query_cams, nq, gallery_ids, gallery_cams, ng, 0, workspace, pair_q, pair_g,
                              pair_pos, nullptr, zero_f32, zero_u32, zero_per_query, capacity, stream);
}

// ---- compacted same-id gallery for the threshold pass (see compact_rep_kernel) ----
static uint32_t compact_slots(long long nq) {
  uint32_t n = 64;
  while ((long long)n < 2 * nq) n <<= 1;
  return n;
}

extern "C" long long pps_pairs_compact_workspace_bytes(long long nq) {
  if (nq < 0 || nq > 0x3fffffffLL) return PPS_ERR_INVALID_ARG;
  return (4 * (nq > 0 ? nq : 1) + 4) * 4 + (long long)compact_slots(nq) * 8;     // rep, lo, cnt, gp_off | hash keys, values
}

namespace pps {
int pairs_compact_rows_ex(const int64_t* query_ids, long long nq, const int32_t* pair_off, const int32_t* pair_q,
                          const int32_t* pair_g, long long n_pairs, long long row_lo, long long row_hi, void* workspace,
                          int32_t* gp_rows, int32_t* pair_col, int32_t* n_rows, void* stream, const int32_t* n_pairs_dev,
                          long long rows_cap /*> 0: pad gp_rows [*n_rows, rows_cap) with row_lo*/) {
  if (nq < 0 || n_pairs < 0 || row_lo < 0 || row_hi < row_lo || nq > 0x7fffffffLL || row_hi > 0x7fffffffLL)
    return PPS_ERR_INVALID_ARG;
  if (!n_rows) return PPS_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nq == 0 || n_pairs == 0) {
    PPS_CUDA_TRY(cudaMemsetAsync(n_rows, 0, 4, st));
    return PPS_OK;
  }
  if (!query_ids || !pair_off || !pair_q || !pair_g || !workspace || !gp_rows || !pair_col) return PPS_ERR_INVALID_ARG;
  int32_t* rep = static_cast<int32_t*>(workspace);
  int32_t* lo = rep + nq;
  int32_t* cnt = lo + nq;
  int32_t* gp_off = cnt + nq;
  const uint32_t slots = compact_slots(nq);
  int32_t* keys = gp_off + nq + 4;
  int32_t* vals = keys + slots;
  PPS_CUDA_TRY(cudaMemsetAsync(keys, 0xFF, (size_t)slots * 4, st));      // -1: empty
  PPS_CUDA_TRY(cudaMemsetAsync(vals, 0x7F, (size_t)slots * 4, st));      // 0x7f7f7f7f: above every query index
  compact_hash_insert_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>((int)nq, pair_off, pair_g, keys, vals, slots - 1);
  PPS_LAUNCH_CHECK("compact_hash_insert_kernel");
  compact_rep_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(query_ids, (int)nq, pair_off, pair_g, (int)row_lo,
                                                                   (int)row_hi, keys, vals, slots - 1, rep, lo, cnt);
  PPS_LAUNCH_CHECK("compact_rep_kernel");
  compact_scan_kernel<<<1, 1024, 0, st>>>(rep, cnt, (int)nq, gp_off, n_rows);
  PPS_LAUNCH_CHECK("compact_scan_kernel");
  compact_fill_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(pair_off, pair_q, pair_g, n_pairs, rep, lo, cnt,
                                                                         gp_off, gp_rows, pair_col, n_pairs_dev);
  PPS_LAUNCH_CHECK("compact_fill_kernel");
  if (rows_cap > 0) {
    compact_pad_kernel<<<(unsigned)((rows_cap + 255) / 256), 256, 0, st>>>(gp_rows, n_rows, rows_cap, (int32_t)row_lo);
    PPS_LAUNCH_CHECK("compact_pad_kernel");
  }
  return PPS_OK;
}
}  // namespace pps

extern "C" int pps_pairs_compact_rows(const int64_t* query_ids, long long nq, const int32_t* pair_off,
                                      const int32_t* pair_q, const int32_t* pair_g, long long n_pairs, long long row_lo,
                                      long long row_hi, void* workspace, int32_t* gp_rows, int32_t* pair_col,
                                      int32_t* n_rows, void* stream) {
  return pairs_compact_rows_ex(query_ids, nq, pair_off, pair_q, pair_g, n_pairs, row_lo, row_hi, workspace, gp_rows, pair_col,
                               n_rows, stream, nullptr, 0);
}

// ---- candidate pre-filter (see qset_build_kernel) ----
static uint32_t qset_slots(long long nq) {
  uint32_t n = 64;
  while ((long long)n < 2 * nq) n <<= 1;
  return n;
}

extern "C" long long pps_pairs_prefilter_workspace_bytes(long long nq, long long ng) {
  if (nq < 0 || ng < 0) return PPS_ERR_INVALID_ARG;
  const long long slots = qset_slots(nq);
  const long long nblk = (ng + kFilterRows - 1) / kFilterRows;
  return slots * 12 + (nblk + 4) * 4 + 64;
}

extern "C" int pps_pairs_prefilter(const int64_t* query_ids, long long nq, const int64_t* gallery_ids,
                                   const int64_t* gallery_cams, long long ng, void* workspace, int32_t* cand_rows,
                                   int64_t* cand_gid, int64_t* cand_gcam, int32_t* n_cand, void* stream) {
  if (nq < 0 || ng < 0 || nq > 0x3fffffffLL || ng > 0x7fffffffLL) return PPS_ERR_INVALID_ARG;
  if (!n_cand) return PPS_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nq == 0 || ng == 0) {
    PPS_CUDA_TRY(cudaMemsetAsync(n_cand, 0, 4, st));
    return PPS_OK;
  }
  if (!query_ids || !gallery_ids || !gallery_cams || !workspace || !cand_rows || !cand_gid || !cand_gcam)
    return PPS_ERR_INVALID_ARG;
  const uint32_t slots = qset_slots(nq);
  int64_t* keys = static_cast<int64_t*>(workspace);
  int32_t* occ = reinterpret_cast<int32_t*>(keys + slots);
  int32_t* blk = occ + slots;
  const long long nblk = (ng + kFilterRows - 1) / kFilterRows;
  PPS_CUDA_TRY(cudaMemsetAsync(occ, 0, (size_t)slots * 4, st));
  qset_build_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(query_ids, (int)nq, keys, occ, slots - 1);
  PPS_LAUNCH_CHECK("qset_build_kernel");
  prefilter_kernel<false><<<(unsigned)nblk, 256, 0, st>>>(gallery_ids, gallery_cams, ng, keys, occ, slots - 1, blk, nullptr,
                                                          nullptr, nullptr);
  PPS_LAUNCH_CHECK("prefilter_kernel<count>");
  prefilter_scan_kernel<<<1, 1024, 0, st>>>(blk, (int)nblk, n_cand);
  PPS_LAUNCH_CHECK("prefilter_scan_kernel");
  prefilter_kernel<true><<<(unsigned)nblk, 256, 0, st>>>(gallery_ids, gallery_cams, ng, keys, occ, slots - 1, blk, cand_rows,
                                                         cand_gid, cand_gcam);
  PPS_LAUNCH_CHECK("prefilter_kernel<fill>");
  return PPS_OK;
}

extern "C" int pps_pairs_remap(int32_t* pair_g, long long n_pairs, const int32_t* cand_rows, long long offset,
                               void* stream) {
  if (n_pairs < 0) return PPS_ERR_INVALID_ARG;
  if (n_pairs == 0) return PPS_OK;
  if (!pair_g || !cand_rows) return PPS_ERR_INVALID_ARG;
  pairs_remap_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pair_g, n_pairs,
                                                                                                       cand_rows, offset, nullptr);
  PPS_LAUNCH_CHECK("pairs_remap_kernel");
  return PPS_OK;
}

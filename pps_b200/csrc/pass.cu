// One ranking pass over a gallery of any size - several distance blocks per device, the gallery optionally sharded over
// several devices - driven from C: what evaluator.RankEngine did launch by launch from Python in round 1.
//
//   pps_pass_begin   pair lists from the GLOBAL id / camera vectors (every rank builds the same lists: hash pre-filter of
//                    the gallery rows that can match at all -> brute-force sweeps over those -> CSR), operand split of the
//                    queries, distance of the FIRST local block (enqueued before the pair lists are waited for, so the host
//                    round trips that size the lists hide under it), thresholds = distances of the same-id pairs that live
//                    in this shard (one block: gathered from it; several: one product with the compacted same-id rows).
//                                                   -> exchange 1: all-reduce(SUM, int32) of *d_x1 [n_x1 words]
//   pps_pass_count   per block: split, tcgen05 distance (+ top-k admission in its epilogue after the first block), counting
//                    sweep, candidate merge.        -> exchange 2: all-gather of *d_x2 [x2_bytes per rank]
//   pps_pass_end     reduce the gathered counters / merge the gathered top-k keys, finalize, results to the host.
// With world == 1 there is nothing to exchange and the three calls just run back to back.
#include "ctx.cuh"

#include <chrono>
#include <cmath>

namespace pps {

constexpr int kTkCap = 2048;            // top-k candidates per query and block the distance epilogue may append
constexpr long long kPrefilterMinRows = 32768;
constexpr long long kThreshCols = 32768;   // columns of the threshold product per launch

// counters / flags of the R gathered buffers summed into this rank's own buffer
__global__ void pass_reduce_counters_kernel(const unsigned char* __restrict__ gathered, size_t stride_bytes, size_t off_bytes,
                                            int world, long long n_words, uint32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  uint32_t s = 0;
  for (int r = 0; r < world; ++r) s += reinterpret_cast<const uint32_t*>(gathered + (size_t)r * stride_bytes + off_bytes)[i];
  out[i] = s;
}

// k smallest of the world x k gathered keys of a query (every rank's list is sorted; keys are unique: they carry the
// global gallery index).  One CTA per query, bitonic sort of <= 4096 keys in shared memory.
__global__ void __launch_bounds__(256) pass_merge_keys_kernel(const unsigned char* __restrict__ gathered, size_t stride_bytes,
                                                              int world, int k, int n_pow2,
                                                              unsigned long long* __restrict__ out) {
  extern __shared__ unsigned long long keys_s[];
  const int q = blockIdx.x, tid = threadIdx.x;
  const int n = world * k;
  for (int i = tid; i < n_pow2; i += 256) {
    unsigned long long v = ~0ull;
    if (i < n) {
      const int r = i / k, j = i - r * k;
      v = reinterpret_cast<const unsigned long long*>(gathered + (size_t)r * stride_bytes)[(long long)q * k + j];
    }
    keys_s[i] = v;
  }
  __syncthreads();
  for (int kk = 2; kk <= n_pow2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_pow2; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = keys_s[i], y = keys_s[ixj];
          const bool up = (i & kk) == 0;
          if ((x > y) == up) { keys_s[i] = y; keys_s[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += 256) out[(long long)q * k + i] = keys_s[i];
}

__global__ void pass_or_flag_kernel(const int32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && *src) *dst = 1u;
}

__global__ void pass_copy_flags_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  if (threadIdx.x < 2 && blockIdx.x == 0) dst[threadIdx.x] = src[threadIdx.x];
}

__global__ void pass_clamp_offsets_kernel(int32_t* __restrict__ pair_off, long long n, int32_t cap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && pair_off[i] > cap) pair_off[i] = cap;
}

// speculative pass: did every actual count stay inside the bound the pass was laid out for?  flags[1] != 0 -> repeat
__global__ void pass_check_bounds_kernel(const int32_t* __restrict__ tot, int32_t cap_pairs, int32_t cap_maxp, int32_t cap_cand,
                                         int32_t cap_rows, uint32_t* __restrict__ flag) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const bool bad = tot[0] > cap_pairs || tot[1] > cap_maxp || (cap_cand >= 0 && tot[2] > cap_cand) ||
                     (cap_rows >= 0 && tot[3] > cap_rows);
    if (bad) *flag = 1u;
  }
}

}  // namespace pps

using namespace pps;

namespace {

// gallery blocks of one pass: with top-k admission in the distance epilogue the FIRST block is kept short - it is the one
// that takes the one-read sweep, and all it has to do is establish an admission bound tight enough that a full block
// admits well under kTkCap candidates per query (rows in random order: ~k * R / F per query for a block of R rows after
// F swept rows).  Galleries that would fit one block are split the same way when they are long enough to profit.
int plan_blocks(PassState& p, bool short_first) {
  p.n_blocks = 0;
  const long long ngl = p.ngl, block_rows = p.chunk;
  if (ngl <= 0) return PPS_OK;
  long long first = std::min(block_rows, ngl);
  if (short_first && p.topk > 0) {
    long long want = (long long)(2.2 * p.topk * (double)std::min(block_rows, ngl) / kTkCap);   // planned for the default cap
    want = std::max<long long>(16384, (want + 255) / 256 * 256);   // (8 192: the weak bound floods the epilogue admission; 32 768: the sweep itself dominates a 65 k-row shard)
    if (ngl >= 4 * want) first = std::min(first, want);
  }
  long long r0 = 0, rows = first;
  while (r0 < ngl) {
    if (p.n_blocks >= kPassMaxBlocks) return PPS_ERR_UNSUPPORTED;
    p.blk_row0[p.n_blocks] = r0;
    p.blk_rows[p.n_blocks] = rows;
    ++p.n_blocks;
    r0 += rows;
    rows = std::min(block_rows, ngl - r0);
  }
  return PPS_OK;
}

// host time spent blocked in a wait (phase timing on): where the pass stalls the CPU
struct HostWait {
  pps_ctx* c; int slot; std::chrono::steady_clock::time_point t0;
  HostWait(pps_ctx* c_, int slot_) : c(c_), slot(slot_), t0(std::chrono::steady_clock::now()) {}
  ~HostWait() {
    if (c->timing) c->pass.host_wait_ms[slot] += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

struct TimedLaunch {          // CUDA events around one launch when phase timing is on
  pps_ctx* c; cudaStream_t s; int slot;
  TimedLaunch(pps_ctx* c_, cudaStream_t s_, int kind) : c(c_), s(s_), slot(-1) {
    PassState& p = c->pass;
    if (!c->timing || p.n_timed >= kPassTimedLaunches) return;
    slot = p.n_timed++;
    p.timed_kind[slot] = kind;
    for (int i = 0; i < 2; ++i)
      if (!p.ev_t[slot][i]) cudaEventCreate(&p.ev_t[slot][i]);
    cudaEventRecord(p.ev_t[slot][0], s);
  }
  ~TimedLaunch() { if (slot >= 0) cudaEventRecord(c->pass.ev_t[slot][1], s); }
};

// operand split of gallery rows [r0, r0 + rows) of the local shard -> (planes pointer, sqnorm pointer) for the GEMM
int split_block(pps_ctx* c, int blk, long long r0, long long rows, cudaStream_t cs, const void** planes_out) {
  PassState& p = c->pass;
  const size_t esz = p.dtype == PPS_DTYPE_F16 ? 2 : 4;
  const unsigned char* src = static_cast<const unsigned char*>(p.d_g) + (size_t)r0 * p.dim * esz;
  if (p.g_from_host && blk >= 0) PPS_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_slab[blk], 0));   // this block's rows have arrived
  TimedLaunch t(c, cs, 1);
  if (p.g_inplace) {         // fp16 rows that already are a K-major operand plane: only the row norms
    *planes_out = src;
    return pps_row_sqnorm(src, PPS_DTYPE_F16, rows, p.dim, p.dim, p.gn.as<float>(), cs);
  }
  *planes_out = p.gs.p;
  return pps_split_rows(src, p.dtype, rows, p.dim, p.dim, p.split_planes, p.gs.p, p.gn.as<float>(), cs);
}

int distance_block(pps_ctx* c, const void* g_planes, long long rows, long long col0, bool admit, cudaStream_t cs) {
  PassState& p = c->pass;
  TimedLaunch t(c, cs, 2);
  if (admit)
    return pps_dist_topk_tc(p.qs.p, p.qn.as<float>(), p.nq, p.planes, 0, g_planes, p.gn.as<float>(), rows, p.planes, 0, p.dim,
                            p.precision, 0, p.dist.as<float>(), p.ldd, col0, p.tk_bound.as<uint32_t>(),
                            p.tk_cnt.as<uint32_t>(), p.tk_cand.as<uint64_t>(), p.tk_cap, cs);
  return pps_dist_tc(p.qs.p, p.qn.as<float>(), p.nq, p.planes, 0, g_planes, p.gn.as<float>(), rows, p.planes, 0, p.dim,
                     p.precision, p.world > 1 ? PPS_DIST_RESERVE_SM_PAIR : 0, p.dist.as<float>(), p.ldd, cs);
}

}  // namespace

extern "C" int pps_pass_begin(pps_ctx* c, const void* d_q, long long nq, const void* d_g, long long ng_local, int dim,
                              int dtype, const int64_t* d_qid, const int64_t* d_qcam, const int64_t* d_gid,
                              const int64_t* d_gcam, long long ng_global, long long gallery_offset, int world, int rank,
                              int precision, int topk, long long max_block_bytes, int flags, void* stream,
                              int32_t** d_x1, long long* n_x1) {
  if (!c) return PPS_ERR_INVALID_ARG;
  PassState& p = c->pass;
  p.active = false;
  if (nq <= 0 || ng_local < 0 || dim <= 0 || ng_global < 0 || gallery_offset < 0 || gallery_offset + ng_local > ng_global ||
      world < 1 || rank < 0 || rank >= world || topk < 0 || topk > PPS_TOPK_MAX || max_block_bytes <= 0)
    return PPS_ERR_INVALID_ARG;
  if (nq > 0x7fffffffLL || ng_global > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
  if (!d_q || !d_qid || !d_qcam || (ng_local > 0 && !d_g) || (ng_global > 0 && (!d_gid || !d_gcam))) return PPS_ERR_INVALID_ARG;
  if (dtype != PPS_DTYPE_F32 && dtype != PPS_DTYPE_F16) return PPS_ERR_INVALID_ARG;
  if (dtype == PPS_DTYPE_F16) precision = PPS_PREC_F16X1;
  switch (precision) {
    case PPS_PREC_BF16X1: p.planes = 1; break;
    case PPS_PREC_BF16X3: p.planes = 2; break;
    case PPS_PREC_BF16X6: p.planes = 3; break;
    case PPS_PREC_F16X3: p.planes = 2; break;
    case PPS_PREC_F16X1: if (dtype != PPS_DTYPE_F16) return PPS_ERR_INVALID_ARG; p.planes = 1; break;
    default: return PPS_ERR_INVALID_ARG;
  }
  if (world > 1 && (long long)world * std::max(topk, 1) > 4096) return PPS_ERR_UNSUPPORTED;
  PPS_CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t cs = static_cast<cudaStream_t>(stream), ss = c->side_s;
  p.nq = nq; p.ngl = ng_local; p.ng_global = ng_global; p.offset = gallery_offset; p.dim = dim; p.kpad = pps_kpad(dim);
  p.dtype = dtype; p.precision = precision; p.topk = topk; p.world = world; p.rank = rank; p.flags = flags;
  p.split_planes = precision == PPS_PREC_F16X3 ? (2 | PPS_SPLIT_F16_SCALED) : p.planes;
  p.d_q = d_q; p.d_g = d_g; p.d_qid = d_qid; p.d_qcam = d_qcam; p.d_gid = d_gid; p.d_gcam = d_gcam;
  p.g_inplace = dtype == PPS_DTYPE_F16 && (dim % 64) == 0 && (reinterpret_cast<uintptr_t>(d_g) & 15u) == 0;
  p.n_timed = 0;
  for (float& w : p.host_wait_ms) w = 0.f;
  p.tk_cap = (flags >> 8) & 0xffff;                    // PPS_PASS_TKCAP(n): candidate buffer entries per query (tests)
  if (p.tk_cap <= 0 || p.tk_cap > kTkCap) p.tk_cap = kTkCap;
  if (!p.ev_a) PPS_CUDA_TRY(cudaEventCreateWithFlags(&p.ev_a, cudaEventDisableTiming));
  if (!p.ev_rows) PPS_CUDA_TRY(cudaEventCreateWithFlags(&p.ev_rows, cudaEventDisableTiming));

  // ---- block plan ----
  long long chunk = std::max<long long>(256, std::min<long long>(std::max<long long>(ng_local, 1), max_block_bytes / (4 * nq)));
  if (chunk < ng_local) chunk = std::max<long long>(256, chunk / 256 * 256);
  p.chunk = chunk;
  p.ldd = (std::min(chunk, std::max<long long>(ng_local, 1)) + 3) / 4 * 4;
  const bool want_epi = topk > 0 && !(flags & PPS_PASS_NO_EPILOGUE_TOPK);
  PPS_TRY(plan_blocks(p, want_epi));
  p.epi_topk = want_epi && p.n_blocks > 1;
  if (p.n_blocks > 0) {             // ldd / block buffer follow the largest block
    long long mx = 0;
    for (int b = 0; b < p.n_blocks; ++b) mx = std::max(mx, p.blk_rows[b]);
    p.ldd = (mx + 3) / 4 * 4;
  }

  // ---- buffers that do not depend on the pair count ----
  PPS_TRY(p.qs.ensure((size_t)pps_split_bytes(nq, dim, p.planes)));
  PPS_TRY(p.qn.ensure((size_t)nq * 8));
  if (!p.g_inplace) PPS_TRY(p.gs.ensure((size_t)pps_split_bytes(std::max<long long>(p.ldd, 1), dim, p.planes)));
  PPS_TRY(p.gn.ensure((size_t)std::max<long long>(p.ldd, 1) * 8));
  PPS_TRY(p.dist.ensure((size_t)nq * p.ldd * 4));
  PPS_TRY(p.pair_off.ensure(((size_t)nq + 1) * 4));
  PPS_TRY(p.totals.ensure(64));
  // per-query results: ONE device buffer [ap f64 | first i32 | valid u8 | pad | flags 2 x u32] mirrored by the pinned
  // staging, so that they come back in a single copy (each small D2H copy costs ~8 us of stream time)
  const size_t off_ap = 64, off_first = off_ap + (size_t)nq * 8, off_valid = off_first + (size_t)nq * 4;
  const size_t off_flags = (off_valid + (size_t)nq + 7) & ~(size_t)7;
  p.res_bytes = off_flags + 8 - off_ap;
  PPS_TRY(p.ap.ensure(p.res_bytes));
  PPS_TRY(c->h_small.ensure(off_flags + 8));
  unsigned char* hb = c->h_small.as<unsigned char>();
  c->st.totals = reinterpret_cast<int32_t*>(hb);
  c->st.ap = reinterpret_cast<double*>(hb + off_ap);
  c->st.first = reinterpret_cast<int32_t*>(hb + off_first);
  c->st.valid = hb + off_valid;
  p.res_first_off = off_first - off_ap; p.res_valid_off = off_valid - off_ap; p.res_flags_off = off_flags - off_ap;
  int32_t* h_tot = c->st.totals;           // [0] n_pairs [1] max_pairs [2] n_cand [3] n_rows [4] flags

  // ---- the ids are valid on `cs`; the pair-list work runs on the side stream ----
  PPS_CUDA_TRY(cudaEventRecord(c->ev_in, cs));
  PPS_CUDA_TRY(cudaStreamWaitEvent(ss, c->ev_in, 0));
  p.prefilter = ng_global >= kPrefilterMinRows;
  if (p.prefilter) {
    PPS_TRY(p.pf_ws.ensure((size_t)pps_pairs_prefilter_workspace_bytes(nq, ng_global)));
    PPS_TRY(p.cand_rows.ensure((size_t)ng_global * 4));
    PPS_TRY(p.cand_gid.ensure((size_t)ng_global * 8));
    PPS_TRY(p.cand_gcam.ensure((size_t)ng_global * 8));
    PPS_TRY(pps_pairs_prefilter(d_qid, nq, d_gid, d_gcam, ng_global, p.pf_ws.p, p.cand_rows.as<int32_t>(),
                                p.cand_gid.as<int64_t>(), p.cand_gcam.as<int64_t>(), p.totals.as<int32_t>() + 2, ss));
    PPS_CUDA_TRY(cudaMemcpyAsync(h_tot + 2, p.totals.as<int32_t>() + 2, 4, cudaMemcpyDeviceToHost, ss));
    PPS_CUDA_TRY(cudaEventRecord(p.ev_a, ss));
  }

  // ---- host input (pps_pass_set_host_input): d_q / d_g are staging buffers this call fills.  The gallery goes up on the
  // copy stream - one block after the other, or, for a shard that is one block, in ~8 row slabs whose split + distance
  // (each writing its column range of the block) overlap the copy of the next slab, as pps_evaluate_host_ctx does ----
  const size_t esz_in = dtype == PPS_DTYPE_F16 ? 2 : 4;
  const void* h_q = p.h_q;
  const void* h_g = p.h_g;
  p.h_q = p.h_g = nullptr;
  p.g_from_host = false;
  if (h_q) PPS_CUDA_TRY(cudaMemcpyAsync(const_cast<void*>(d_q), h_q, (size_t)nq * dim * esz_in, cudaMemcpyHostToDevice, cs));
  bool slabbed = false;
  if (h_g && ng_local > 0) {
    unsigned char* dg = static_cast<unsigned char*>(const_cast<void*>(d_g));
    const unsigned char* hg = static_cast<const unsigned char*>(h_g);
    PPS_CUDA_TRY(cudaStreamWaitEvent(c->copy_s, c->ev_in, 0));
    if (p.n_blocks == 1) {
      slabbed = true;
    } else if (p.n_blocks <= kMaxSlabs) {
      p.g_from_host = true;
      for (int b = 0; b < p.n_blocks; ++b) {
        const size_t off = (size_t)p.blk_row0[b] * dim * esz_in;
        PPS_CUDA_TRY(cudaMemcpyAsync(dg + off, hg + off, (size_t)p.blk_rows[b] * dim * esz_in, cudaMemcpyHostToDevice, c->copy_s));
        PPS_CUDA_TRY(cudaEventRecord(c->ev_slab[b], c->copy_s));
      }
    } else {
      PPS_CUDA_TRY(cudaMemcpyAsync(dg, hg, (size_t)ng_local * dim * esz_in, cudaMemcpyHostToDevice, cs));
    }
  }

  // ---- main stream: operand split of the queries and of the first block.  These are shared-memory-free streaming kernels,
  // so the pair-list kernels of the side stream run beside them.  The DISTANCE of the first block is enqueued only after
  // the pair lists: its persistent CTAs fill the shared memory of every SM, nothing else can start while it runs, and
  // the host round trips that size the lists would otherwise each wait for it (measured: 2.2 ms per pass at 520 k rows) ----
  {
    TimedLaunch t(c, cs, 1);
    PPS_TRY(pps_split_rows(d_q, dtype, nq, dim, dim, p.split_planes, p.qs.p, p.qn.as<float>(), cs));
  }
  const void* gp0 = nullptr;
  long long slab = 0;
  if (slabbed) {
    unsigned char* dg = static_cast<unsigned char*>(const_cast<void*>(d_g));
    const unsigned char* hg = static_cast<const unsigned char*>(h_g);
    slab = ((ng_local + 7) / 8 + 255) & ~255LL;
    if (slab < 1024) slab = 1024;
    while ((ng_local + slab - 1) / slab > kMaxSlabs) slab *= 2;
    int si = 0;
    for (long long r0 = 0; r0 < ng_local; r0 += slab, ++si) {          // all uploads are on their way before any wait
      const long long nr = std::min(slab, ng_local - r0);
      const size_t off = (size_t)r0 * dim * esz_in;
      PPS_CUDA_TRY(cudaMemcpyAsync(dg + off, hg + off, (size_t)nr * dim * esz_in, cudaMemcpyHostToDevice, c->copy_s));
      PPS_CUDA_TRY(cudaEventRecord(c->ev_slab[si], c->copy_s));
    }
  } else if (p.n_blocks > 0) {
    PPS_TRY(split_block(c, 0, p.blk_row0[0], p.blk_rows[0], cs, &gp0));
  }

  // ---- pair lists.  SIZING pass (the first one for a shape, or after an overflow): three 4-byte read-backs size the
  // lists.  SPECULATIVE pass (every later one): the sizes of the sizing pass are used as upper bounds, the actual counts
  // are read by the kernels on the device, nothing comes back to the host until the results do - the CPU runs ahead of
  // the GPU for the whole pass - and pps_pass_end checks the bounds (PPS_ERR_PASS_RESIZE -> the caller repeats the pass
  // as a sizing pass).  The lists are global, so every rank of a sharded run takes the same decision. ----
  PassState::Hints& hn = p.hints;
  p.speculative = !(flags & PPS_PASS_SIZING) && hn.valid && hn.nq == nq && hn.ng_global == ng_global && hn.ngl == ng_local &&
                  hn.offset == gallery_offset && hn.mbb == max_block_bytes && hn.dim == dim && hn.dtype == dtype &&
                  hn.topk == topk && hn.world == world && hn.precision == precision;
  const bool spec = p.speculative;
  int32_t* d_tot = p.totals.as<int32_t>();          // [0] n_pairs [1] max_pairs [2] n_cand [3] n_rows
  const int64_t* sweep_gid = d_gid;
  const int64_t* sweep_gcam = d_gcam;
  long long sweep_rows = ng_global;
  const int32_t* dn_cand = nullptr;
  if (p.prefilter) {
    if (spec) {
      p.cap_cand = std::min<long long>(ng_global, (hn.n_cand + 1023) / 1024 * 1024);
    } else {
      {
        HostWait hw(c, 0);
        PPS_CUDA_TRY(cudaEventSynchronize(p.ev_a));
      }
      p.n_cand = h_tot[2];
      p.cap_cand = p.n_cand;
    }
    sweep_gid = p.cand_gid.as<int64_t>(); sweep_gcam = p.cand_gcam.as<int64_t>(); sweep_rows = p.cap_cand;
    dn_cand = d_tot + 2;
  }
  PPS_TRY(p.pair_ws.ensure((size_t)std::max<long long>(pps_pairs_workspace_bytes(nq, sweep_rows), 16)));
  PPS_TRY(pairs_count_device_ex(d_qid, nq, sweep_gid, sweep_rows, p.pair_ws.p, p.pair_off.as<int32_t>(), d_tot, ss, dn_cand));
  if (spec) {
    p.n_pairs = hn.n_pairs > 0 ? (hn.n_pairs + hn.n_pairs / 8 + 1023) / 1024 * 1024 : 0;     // capacity (layout) of this pass
    p.max_pairs = hn.max_pairs > 0 ? (hn.max_pairs + 8 + 15) / 16 * 16 : 0;
    // lists that outgrew the capacity are cut (memory safety); pps_pass_end reports it and the pass is repeated
    pass_clamp_offsets_kernel<<<(unsigned)((nq + 1 + 255) / 256), 256, 0, ss>>>(p.pair_off.as<int32_t>(), nq + 1, (int32_t)p.n_pairs);
    PPS_LAUNCH_CHECK("pass_clamp_offsets_kernel");
  } else {
    PPS_CUDA_TRY(cudaMemcpyAsync(h_tot, d_tot, 8, cudaMemcpyDeviceToHost, ss));
    PPS_CUDA_TRY(cudaEventRecord(c->ev_totals, ss));
    {
      HostWait hw(c, 1);
      PPS_CUDA_TRY(cudaEventSynchronize(c->ev_totals));
    }
    p.n_pairs = h_tot[0];
    p.max_pairs = h_tot[1];
  }
  const int32_t* dn_pairs = spec ? d_tot : nullptr;
  const size_t np1 = (size_t)std::max<long long>(p.n_pairs, 1);
  // single-plane operands, several blocks: the blocks after the first take the counting epilogue (+ admission) and are never
  // written.  (With a 2-plane split the threshold tables would cost the operand ring a stage, and the 3-term mainloop
  // gains nothing from a lighter epilogue: those keep block + count kernel.)
  {
    const int need = spec ? hn.max_pairs + 8 : p.max_pairs;
    p.fused = p.n_blocks > 1 && p.planes == 1 && p.n_pairs > 0 && need <= 64 && !(flags & PPS_PASS_NO_FUSED_COUNT) &&
              (topk == 0 || p.epi_topk);
    p.p_cap = std::max(8, (need + 7) / 8 * 8);
  }
  PPS_TRY(p.small.ensure(16));
  PPS_CUDA_TRY(cudaMemsetAsync(p.small.p, 0, 16, ss));
  PPS_TRY(p.pair_q.ensure(np1 * 4));
  PPS_TRY(p.pair_g.ensure(np1 * 4));
  PPS_TRY(p.pair_pos.ensure(np1));
  PPS_TRY(p.pair_d.ensure(np1 * 4));
  p.packed_bytes = ((size_t)nq * topk * 8 + p.counter_words() * 4 + 15) & ~(size_t)15;
  PPS_TRY(p.packed.ensure(p.packed_bytes));
  // the fill zero-fills pair_d, cnt_le and cnt_first of the slots it writes (a sizing pass: all of them, the lists are
  // global; a speculative pass exchanges whole capacities, so they are cleared in full first)
  if (spec) {
    PPS_CUDA_TRY(cudaMemsetAsync(p.pair_d.p, 0, np1 * 4, ss));
    PPS_CUDA_TRY(cudaMemsetAsync(p.cnt_first(), 0, p.counter_words() * 4, ss));
    PPS_CUDA_TRY(cudaMemsetAsync(p.pair_q.p, 0xFF, np1 * 4, ss));          // unused slots: query -1 (skipped by the gathers)
    PPS_CUDA_TRY(cudaMemsetAsync(p.pair_pos.p, 0, np1, ss));
    PPS_CUDA_TRY(cudaMemsetAsync(p.pair_g.p, 0, np1 * 4, ss));
  } else {
    PPS_CUDA_TRY(cudaMemsetAsync(p.flags_dev(), 0, 8, ss));
    if (p.n_pairs == 0) PPS_CUDA_TRY(cudaMemsetAsync(p.cnt_first(), 0, (size_t)nq * 4, ss));
  }
  PPS_TRY(pairs_fill_device_ex(d_qid, d_qcam, nq, sweep_gid, sweep_gcam, sweep_rows, p.pair_ws.p, p.pair_q.as<int32_t>(),
                               p.pair_g.as<int32_t>(), p.pair_pos.as<uint8_t>(), p.pair_d.as<float>(), p.cnt_le(),
                               p.cnt_first(), p.n_pairs, ss, dn_cand));
  if (p.prefilter && p.n_pairs > 0)
    PPS_TRY(pairs_remap_ex(p.pair_g.as<int32_t>(), p.n_pairs, p.cand_rows.as<int32_t>(), 0, ss, dn_pairs));
  if (topk > 0) {
    PPS_TRY(pps_topk_init(reinterpret_cast<uint64_t*>(p.keys()), nq, topk, ss));
    if (p.epi_topk) {
      PPS_TRY(p.tk_bound.ensure((size_t)nq * 4));
      PPS_TRY(p.tk_cnt.ensure((size_t)nq * 4));
      PPS_TRY(p.tk_cand.ensure((size_t)nq * p.tk_cap * 8));
    }
  }
  // several blocks: the thresholds come from one product with the compacted same-id rows of this shard
  p.n_rows = 0;
  p.cap_rows = 0;
  PPS_CUDA_TRY(cudaMemsetAsync(d_tot + 3, 0, 4, ss));
  if (p.n_blocks > 1 && p.n_pairs > 0) {
    PPS_TRY(p.gp_ws.ensure((size_t)pps_pairs_compact_workspace_bytes(nq)));
    if (spec) p.cap_rows = hn.n_rows > 0 ? std::min<long long>((hn.n_rows + 255) / 256 * 256, std::max<long long>(p.ngl, 1)) : 0;
    PPS_TRY(p.gp_rows.ensure(std::max(np1, (size_t)p.cap_rows) * 4));
    PPS_TRY(p.pair_col.ensure(np1 * 4));
    PPS_TRY(pairs_compact_rows_ex(d_qid, nq, p.pair_off.as<int32_t>(), p.pair_q.as<int32_t>(), p.pair_g.as<int32_t>(),
                                  p.n_pairs, p.offset, p.offset + p.ngl, p.gp_ws.p, p.gp_rows.as<int32_t>(),
                                  p.pair_col.as<int32_t>(), d_tot + 3, ss, dn_pairs, spec ? p.cap_rows : 0));
    if (spec) {
      p.n_rows = p.cap_rows;
    } else {
      PPS_CUDA_TRY(cudaMemcpyAsync(h_tot + 3, d_tot + 3, 4, cudaMemcpyDeviceToHost, ss));
      PPS_CUDA_TRY(cudaEventRecord(p.ev_rows, ss));
      {
        HostWait hw(c, 2);
        PPS_CUDA_TRY(cudaEventSynchronize(p.ev_rows));
      }
      p.n_rows = h_tot[3];
    }
  }
  if (!spec) {     // what the next pass of this shape may assume
    hn.valid = true; hn.nq = nq; hn.ng_global = ng_global; hn.ngl = ng_local; hn.offset = gallery_offset; hn.mbb = max_block_bytes;
    hn.dim = dim; hn.dtype = dtype; hn.topk = topk; hn.world = world; hn.precision = precision;
    hn.n_cand = p.prefilter ? p.n_cand : 0; hn.n_pairs = p.n_pairs; hn.max_pairs = p.max_pairs; hn.n_rows = p.n_rows;
  }
  PPS_CUDA_TRY(cudaEventRecord(c->ev_pairs, ss));

  // ---- distance of the first block (host input, one block: per slab, as the rows arrive) ----
  if (slabbed) {
    unsigned char* dg = static_cast<unsigned char*>(const_cast<void*>(d_g));
    int si = 0;
    for (long long r0 = 0; r0 < ng_local; r0 += slab, ++si) {
      const long long nr = std::min(slab, ng_local - r0);
      const size_t off = (size_t)r0 * dim * esz_in;
      PPS_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_slab[si], 0));
      const void* bp;
      {
        TimedLaunch t(c, cs, 1);
        if (p.g_inplace) {
          bp = dg + off;
          PPS_TRY(pps_row_sqnorm(dg + off, PPS_DTYPE_F16, nr, dim, dim, p.gn.as<float>() + r0, cs));
        } else {
          bp = p.gs.as<unsigned char>() + (size_t)r0 * p.kpad * 2;
          PPS_TRY(pps_split_rows_slab(dg, dtype, r0, nr, ng_local, dim, dim, p.split_planes, p.gs.p, p.gn.as<float>(), cs));
        }
      }
      TimedLaunch t(c, cs, 2);
      PPS_TRY(pps_dist_tc(p.qs.p, p.qn.as<float>(), nq, p.planes, 0, bp, p.gn.as<float>() + r0, nr, p.planes, ng_local, dim,
                          precision, world > 1 ? PPS_DIST_RESERVE_SM_PAIR : 0, p.dist.as<float>() + r0, p.ldd, cs));
    }
  } else if (p.n_blocks > 0) {
    PPS_TRY(distance_block(c, gp0, p.blk_rows[0], p.offset + p.blk_row0[0], false, cs));
  }
  PPS_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_pairs, 0));

  // ---- thresholds ----
  if (p.n_pairs > 0 && p.n_blocks == 1) {
    PPS_TRY(rank_gather_ex(p.dist.as<float>(), p.ldd, nq, p.blk_rows[0], p.offset, p.pair_q.as<int32_t>(),
                           p.pair_g.as<int32_t>(), p.n_pairs, p.pair_d.as<float>(), cs, dn_pairs));
  } else if (p.n_rows > 0) {
    if (p.g_from_host)       // the compacted rows come from all over the shard: the whole upload must have landed
      PPS_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_slab[p.n_blocks - 1], 0));
    const long long tcols = std::min<long long>(p.n_rows, kThreshCols);
    const long long ldt = (tcols + 3) / 4 * 4;
    PPS_TRY(p.tdist.ensure((size_t)nq * ldt * 4));
    PPS_TRY(p.ts.ensure((size_t)pps_split_bytes(tcols, dim, p.planes)));
    PPS_TRY(p.tn.ensure((size_t)tcols * 8));
    for (long long c0 = 0; c0 < p.n_rows; c0 += tcols) {
      const long long rows = std::min(tcols, p.n_rows - c0);
      {
        TimedLaunch t(c, cs, 1);
        PPS_TRY(pps_split_rows_gather(d_g, dtype, p.gp_rows.as<int32_t>() + c0, p.offset, rows, dim, dim, p.split_planes,
                                      p.ts.p, p.tn.as<float>(), cs));
      }
      {
        TimedLaunch t(c, cs, 2);
        PPS_TRY(pps_dist_tc(p.qs.p, p.qn.as<float>(), nq, p.planes, 0, p.ts.p, p.tn.as<float>(), rows, p.planes, 0, dim,
                            precision, 0, p.tdist.as<float>(), ldt, cs));
      }
      PPS_TRY(rank_gather_ex(p.tdist.as<float>(), ldt, nq, rows, c0, p.pair_q.as<int32_t>(), p.pair_col.as<int32_t>(),
                             p.n_pairs, p.pair_d.as<float>(), cs, dn_pairs));
    }
  }
  p.active = true;
  if (d_x1) *d_x1 = p.pair_d.as<int32_t>();
  if (n_x1) *n_x1 = p.n_pairs;
  return PPS_OK;
}

extern "C" int pps_pass_count(pps_ctx* c, const void* d_gathered_x1, void* stream, void** d_x2, long long* x2_bytes) {
  if (!c || !c->pass.active) return PPS_ERR_INVALID_ARG;
  PassState& p = c->pass;
  if (p.world > 1 && p.n_pairs > 0 && !d_gathered_x1) return PPS_ERR_INVALID_ARG;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const bool have_pairs = p.n_pairs > 0;
  if (p.world > 1 && have_pairs) {     // thresholds: every pair lives on one shard, the others hold 0 -> the sum is exact
    pass_reduce_counters_kernel<<<(unsigned)((p.n_pairs + 255) / 256), 256, 0, cs>>>(
        static_cast<const unsigned char*>(d_gathered_x1), (size_t)p.n_pairs * 4, 0, p.world, p.n_pairs,
        p.pair_d.as<uint32_t>());
    PPS_LAUNCH_CHECK("pass_reduce_counters_kernel");
  }
  auto count_block = [&](long long rows, long long col0, bool with_topk) -> int {
    TimedLaunch t(c, cs, 4);
    if (with_topk)
      return pps_rank_sweep(p.dist.as<float>(), p.ldd, p.nq, rows, col0, p.pair_off.as<int32_t>(),
                            have_pairs ? p.pair_g.as<int32_t>() : nullptr, have_pairs ? p.pair_pos.as<uint8_t>() : nullptr,
                            have_pairs ? p.pair_d.as<float>() : nullptr, have_pairs ? p.max_pairs : 0,
                            have_pairs ? p.cnt_le() : nullptr, p.cnt_first(), reinterpret_cast<uint64_t*>(p.keys()), p.topk,
                            1, cs);
    if (!have_pairs) return PPS_OK;
    return pps_rank_count(p.dist.as<float>(), p.ldd, p.nq, rows, col0, p.pair_off.as<int32_t>(), p.pair_g.as<int32_t>(),
                          p.pair_pos.as<uint8_t>(), p.pair_d.as<float>(), p.max_pairs, p.cnt_le(), p.cnt_first(), cs);
  };
  if (p.fused) {                       // tables of the counting epilogue from the (now global) thresholds
    const size_t elems = (size_t)pps_rank_tab_elems(p.nq, p.p_cap);
    PPS_TRY(p.thr_tab.ensure(elems * 4));
    PPS_TRY(p.tpair_tab.ensure(elems * 4));
    PPS_TRY(p.cnt_tab.ensure(elems * 4));
    PPS_TRY(p.dstar.ensure((size_t)p.nq * 4));
    PPS_TRY(p.gstar.ensure((size_t)p.nq * 4));
    TimedLaunch t(c, cs, 4);
    PPS_TRY(pps_rank_tab_prep(p.nq, p.pair_off.as<int32_t>(), p.pair_g.as<int32_t>(), p.pair_pos.as<uint8_t>(),
                              p.pair_d.as<float>(), p.p_cap, p.thr_tab.as<float>(), p.tpair_tab.as<int32_t>(),
                              p.cnt_tab.as<uint32_t>(), p.dstar.as<float>(), p.gstar.as<int32_t>(), p.small.as<int32_t>() + 1, cs));
  }
  for (int b = 0; b < p.n_blocks; ++b) {
    const long long r0 = p.blk_row0[b], rows = p.blk_rows[b], col0 = p.offset + r0;
    if (b == 0) {                      // its distance block was enqueued by pps_pass_begin
      PPS_TRY(count_block(rows, col0, p.topk > 0));
      continue;
    }
    const void* gp = nullptr;
    PPS_TRY(split_block(c, b, r0, rows, cs, &gp));
    if (p.fused) {
      if (p.epi_topk && b == 1)
        PPS_TRY(pps_topk_bound(reinterpret_cast<const uint64_t*>(p.keys()), p.nq, p.topk, p.tk_bound.as<uint32_t>(),
                               p.tk_cnt.as<uint32_t>(), cs));
      {
        TimedLaunch t(c, cs, 2);
        PPS_TRY(pps_dist_rank_topk_tc(p.qs.p, p.qn.as<float>(), p.nq, p.planes, 0, gp, p.gn.as<float>(), rows, p.planes, 0,
                                      p.dim, p.precision, 0, col0, p.p_cap, p.thr_tab.as<float>(), p.cnt_tab.as<uint32_t>(),
                                      p.dstar.as<float>(), p.gstar.as<int32_t>(), p.cnt_first(),
                                      p.epi_topk ? p.tk_bound.as<uint32_t>() : nullptr,
                                      p.epi_topk ? p.tk_cnt.as<uint32_t>() : nullptr,
                                      p.epi_topk ? p.tk_cand.as<uint64_t>() : nullptr, p.tk_cap, cs));
      }
      if (p.epi_topk) {
        TimedLaunch t(c, cs, 4);
        PPS_TRY(pps_topk_merge(reinterpret_cast<uint64_t*>(p.keys()), p.nq, p.topk, p.tk_cand.as<uint64_t>(), p.tk_cap,
                               p.tk_cnt.as<uint32_t>(), p.tk_bound.as<uint32_t>(), p.pair_off.as<int32_t>(),
                               p.pair_g.as<int32_t>(), p.pair_pos.as<uint8_t>(), p.max_pairs, 1, p.small.as<int32_t>(), cs));
      }
    } else if (p.epi_topk) {
      if (b == 1)
        PPS_TRY(pps_topk_bound(reinterpret_cast<const uint64_t*>(p.keys()), p.nq, p.topk, p.tk_bound.as<uint32_t>(),
                               p.tk_cnt.as<uint32_t>(), cs));
      PPS_TRY(distance_block(c, gp, rows, col0, true, cs));
      PPS_TRY(count_block(rows, col0, false));
      TimedLaunch t(c, cs, 4);
      PPS_TRY(pps_topk_merge(reinterpret_cast<uint64_t*>(p.keys()), p.nq, p.topk, p.tk_cand.as<uint64_t>(), p.tk_cap,
                             p.tk_cnt.as<uint32_t>(), p.tk_bound.as<uint32_t>(), p.pair_off.as<int32_t>(),
                             have_pairs ? p.pair_g.as<int32_t>() : nullptr, have_pairs ? p.pair_pos.as<uint8_t>() : nullptr,
                             have_pairs ? p.max_pairs : 0, 1, p.small.as<int32_t>(), cs));
    } else {
      PPS_TRY(distance_block(c, gp, rows, col0, false, cs));
      PPS_TRY(count_block(rows, col0, p.topk > 0));
    }
  }
  if (p.fused) {                       // counters of the tables -> cnt_le of the pairs; a query with more positives than the
    TimedLaunch t(c, cs, 4);           // tables hold (only a speculative pass can meet one) -> flags[1]: the pass is repeated
    PPS_TRY(pps_rank_tab_finish(p.nq, p.p_cap, p.tpair_tab.as<int32_t>(), p.cnt_tab.as<uint32_t>(), p.cnt_le(), cs));
    pass_or_flag_kernel<<<1, 32, 0, cs>>>(p.small.as<int32_t>() + 1, p.flags_dev() + 1);
    PPS_LAUNCH_CHECK("pass_or_flag_kernel");
  }
  if (p.speculative) {                 // bounds of the speculative layout -> flags[1] (summed over ranks by the exchange)
    pass_check_bounds_kernel<<<1, 32, 0, cs>>>(p.totals.as<int32_t>(), (int32_t)p.n_pairs, (int32_t)p.max_pairs,
                                               p.prefilter ? (int32_t)p.cap_cand : -1,
                                               p.n_blocks > 1 ? (int32_t)p.cap_rows : -1, p.flags_dev() + 1);
    PPS_LAUNCH_CHECK("pass_check_bounds_kernel");
  }
  if (p.epi_topk) {                    // candidate-buffer overflow of any block -> flags[0]
    pass_or_flag_kernel<<<1, 32, 0, cs>>>(p.small.as<int32_t>(), p.flags_dev());
    PPS_LAUNCH_CHECK("pass_or_flag_kernel");
  }
  if (d_x2) *d_x2 = p.packed.p;
  if (x2_bytes) *x2_bytes = (long long)p.packed_bytes;
  return PPS_OK;
}

extern "C" int pps_pass_end(pps_ctx* c, const void* d_gathered, int cmc_topk, void* stream, double* out_map, double* out_cmc,
                            double* out_ap, uint8_t* out_valid, int32_t* out_first_rank, int32_t* out_topk_index,
                            float* out_topk_dist) {
  if (!c || !c->pass.active || cmc_topk < 0) return PPS_ERR_INVALID_ARG;
  PassState& p = c->pass;
  if (p.world > 1 && !d_gathered) return PPS_ERR_INVALID_ARG;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const long long nq = p.nq;
  const int topk = p.topk;
  if (p.world > 1) {
    const unsigned char* gth = static_cast<const unsigned char*>(d_gathered);
    const long long words = (long long)p.counter_words();
    pass_reduce_counters_kernel<<<(unsigned)((words + 255) / 256), 256, 0, cs>>>(gth, p.packed_bytes, (size_t)nq * topk * 8,
                                                                                p.world, words, p.cnt_first());
    PPS_LAUNCH_CHECK("pass_reduce_counters_kernel");
    if (topk > 0) {
      int n_pow2 = 2;
      while (n_pow2 < p.world * topk) n_pow2 <<= 1;
      pass_merge_keys_kernel<<<(unsigned)nq, 256, (size_t)n_pow2 * 8, cs>>>(gth, p.packed_bytes, p.world, topk, n_pow2, p.keys());
      PPS_LAUNCH_CHECK("pass_merge_keys_kernel");
    }
  }
  {
    TimedLaunch t(c, cs, 5);
    PPS_TRY(pps_rank_finalize(nq, p.pair_off.as<int32_t>(), p.pair_g.as<int32_t>(), p.pair_pos.as<uint8_t>(),
                              p.pair_d.as<float>(), p.cnt_le(), p.cnt_first(), p.ap.as<double>(),
                              p.ap.as<uint8_t>() + p.res_valid_off,
                              reinterpret_cast<int32_t*>(p.ap.as<unsigned char>() + p.res_first_off), nullptr, cs));
    pass_copy_flags_kernel<<<1, 32, 0, cs>>>(p.flags_dev(), reinterpret_cast<uint32_t*>(p.ap.as<unsigned char>() + p.res_flags_off));
    PPS_LAUNCH_CHECK("pass_copy_flags_kernel");
    if (topk > 0) {
      PPS_TRY(p.tki.ensure((size_t)nq * topk * 4));
      PPS_TRY(p.tkd.ensure((size_t)nq * topk * 4));
      PPS_TRY(pps_topk_unpack(reinterpret_cast<const uint64_t*>(p.keys()), nq, topk, p.tkd.as<float>(), p.tki.as<int32_t>(), cs));
    }
  }
  const Staging& st = c->st;
  PPS_CUDA_TRY(cudaMemcpyAsync(st.ap, p.ap.p, p.res_bytes, cudaMemcpyDeviceToHost, cs));
  if (topk > 0 && out_topk_index)
    PPS_CUDA_TRY(cudaMemcpyAsync(out_topk_index, p.tki.p, (size_t)nq * topk * 4, cudaMemcpyDeviceToHost, cs));
  if (topk > 0 && out_topk_dist)
    PPS_CUDA_TRY(cudaMemcpyAsync(out_topk_dist, p.tkd.p, (size_t)nq * topk * 4, cudaMemcpyDeviceToHost, cs));
  {
    HostWait hw(c, 3);
    PPS_CUDA_TRY(cudaStreamSynchronize(cs));
  }
  p.active = false;
  if (c->timing) {                      // phase sums of this pass: 1 split, 2 distance, 4 counting / merge, 5 finalize
    for (int i = 0; i < PPS_N_PHASES; ++i) c->phase_ms[i] = 0.f;
    for (int i = 0; i < p.n_timed; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.ev_t[i][0], p.ev_t[i][1]) == cudaSuccess) c->phase_ms[p.timed_kind[i]] += ms;
    }
    // host-side stalls of this pass: 0 candidate count + pair totals + compacted-row count (pps_pass_begin), 3 unused,
    // 6 the final synchronise of pps_pass_end
    c->phase_ms[0] = p.host_wait_ms[0] + p.host_wait_ms[1] + p.host_wait_ms[2];
    c->phase_ms[6] = p.host_wait_ms[3];
  }
  const uint32_t* h_flags = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(st.ap) + p.res_flags_off);
  if (h_flags[1] != 0) {               // a speculative bound was too small somewhere: identical on every rank (flags are summed)
    p.hints.valid = false;
    return PPS_ERR_PASS_RESIZE;
  }
  if (h_flags[0] != 0) return PPS_ERR_TOPK_OVERFLOW;       // identical on every rank (the flags were summed)
  double ap_sum = 0.0;
  long long n_valid = 0;
  std::vector<double> hist((size_t)std::max(cmc_topk, 1), 0.0);
  for (long long i = 0; i < nq; ++i) {
    if (!st.valid[i]) continue;
    ++n_valid;
    ap_sum += st.ap[i];
    if (st.first[i] >= 0 && st.first[i] < cmc_topk) hist[(size_t)st.first[i]] += 1.0;
  }
  if (out_ap) std::memcpy(out_ap, st.ap, (size_t)nq * 8);
  if (out_valid) std::memcpy(out_valid, st.valid, (size_t)nq);
  if (out_first_rank) std::memcpy(out_first_rank, st.first, (size_t)nq * 4);
  if (n_valid == 0) return PPS_ERR_NO_VALID_QUERY;
  if (out_map) *out_map = ap_sum / (double)n_valid;
  double run = 0.0;
  for (int k = 0; k < cmc_topk && out_cmc; ++k) {
    run += hist[(size_t)k];
    out_cmc[k] = run / (double)n_valid;
  }
  return PPS_OK;
}

// What the last pps_pass_begin decided (tests, bench records): 0 blocks of the local shard, 1 counting epilogue used for the
// blocks after the first, 2 speculative layout, 3 thresholds per query in the epilogue tables, 4 top-k admission in the epilogue.
extern "C" long long pps_pass_stat(const pps_ctx* c, int which) {
  if (!c) return -1;
  const PassState& p = c->pass;
  switch (which) {
    case 0: return p.n_blocks;
    case 1: return p.fused ? 1 : 0;
    case 2: return p.speculative ? 1 : 0;
    case 3: return p.fused ? p.p_cap : 0;
    case 4: return p.epi_topk ? 1 : 0;
    default: return -1;
  }
}

// Host sources of the NEXT pps_pass_begin: its d_q / d_g arguments are then device STAGING buffers the call fills from
// these (pinned) host rows - queries on `stream`, the gallery on the ctx's copy stream, block by block (or slab by slab
// when the shard is one block), overlapping the distance of what has already arrived.  NULL keeps a buffer as it is.
extern "C" int pps_pass_set_host_input(pps_ctx* c, const void* h_q, const void* h_g) {
  if (!c) return PPS_ERR_INVALID_ARG;
  c->pass.h_q = h_q;
  c->pass.h_g = h_g;
  return PPS_OK;
}

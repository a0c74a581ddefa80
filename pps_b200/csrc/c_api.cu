// Host-side pieces of the C ABI: error text, launch counter, same-id pair lists, and the
// whole-evaluation entry point that works from host buffers (pps_evaluate_host).
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

namespace pps {

static std::atomic<unsigned long long> g_launches{0};
static thread_local std::string g_cuda_err;

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what) {
  g_cuda_err = std::string(what ? what : "?") + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return PPS_ERR_CUDA;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

// gallery indices sorted by (id, index); the same-id run of a query is a contiguous range
struct IdIndex {
  std::vector<int32_t> order;
  std::vector<int64_t> sorted_ids;
  IdIndex(const int64_t* ids, long long n) : order((size_t)n), sorted_ids((size_t)n) {
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [ids](int32_t a, int32_t b) { return ids[a] < ids[b]; });
    for (long long i = 0; i < n; ++i) sorted_ids[(size_t)i] = ids[order[(size_t)i]];
  }
  void range(int64_t id, size_t* lo, size_t* hi) const {
    auto r = std::equal_range(sorted_ids.begin(), sorted_ids.end(), id);
    *lo = (size_t)(r.first - sorted_ids.begin());
    *hi = (size_t)(r.second - sorted_ids.begin());
  }
};

}  // namespace pps

using namespace pps;

extern "C" int pps_abi_version(void) { return PPS_ABI_VERSION; }

extern "C" const char* pps_strerror(int code) {
  switch (code) {
    case PPS_OK: return "ok";
    case PPS_ERR_INVALID_ARG: return "invalid argument (null pointer, negative size or unknown enum)";
    case PPS_ERR_SHAPE: return "shape check failed";
    case PPS_ERR_ALIGN: return "pointer or leading dimension not aligned";
    case PPS_ERR_CUDA: return "CUDA call failed";
    case PPS_ERR_UNSUPPORTED: return "request not supported by this build";
    case PPS_ERR_WORKSPACE: return "workspace too small";
    case PPS_ERR_NO_VALID_QUERY: return "No valid query";
    case PPS_ERR_PASS_RESIZE: return "a speculative size bound of the pass was too small: repeat it with PPS_PASS_SIZING";
    case PPS_ERR_TOPK_OVERFLOW: return "top-k candidate buffer overflow: repeat the pass with PPS_PASS_NO_EPILOGUE_TOPK";
    default: return "unknown error code";
  }
}

extern "C" const char* pps_last_cuda_error(void) { return g_cuda_err.c_str(); }

extern "C" unsigned long long pps_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" long long pps_pairs_count(const int64_t* query_ids, long long nq, const int64_t* gallery_ids,
                                     long long ng) {
  if (nq < 0 || ng < 0 || ng > 0x7fffffffLL) return PPS_ERR_INVALID_ARG;
  if (nq == 0 || ng == 0) return 0;
  if (!query_ids || !gallery_ids) return PPS_ERR_INVALID_ARG;
  std::vector<int64_t> sorted(gallery_ids, gallery_ids + ng);
  std::sort(sorted.begin(), sorted.end());
  long long total = 0;
  for (long long i = 0; i < nq; ++i) {
    auto r = std::equal_range(sorted.begin(), sorted.end(), query_ids[i]);
    total += (long long)(r.second - r.first);
  }
  return total;
}

extern "C" int pps_pairs_fill(const int64_t* query_ids, const int64_t* query_cams, long long nq,
                              const int64_t* gallery_ids, const int64_t* gallery_cams, long long ng,
                              int32_t* pair_off, int32_t* pair_q, int32_t* pair_g, uint8_t* pair_pos) {
  if (nq < 0 || ng < 0 || ng > 0x7fffffffLL || nq > 0x7fffffffLL) return PPS_ERR_INVALID_ARG;
  if (!pair_off) return PPS_ERR_INVALID_ARG;
  pair_off[0] = 0;
  if (nq == 0) return PPS_OK;
  if (!query_ids || !query_cams || (ng > 0 && (!gallery_ids || !gallery_cams))) return PPS_ERR_INVALID_ARG;
  IdIndex idx(gallery_ids, ng);
  long long e = 0;
  for (long long i = 0; i < nq; ++i) {
    size_t lo, hi;
    idx.range(query_ids[i], &lo, &hi);
    if (e + (long long)(hi - lo) > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
    for (size_t s = lo; s < hi; ++s, ++e) {   // stable sort => ascending gallery index inside the run
      const int32_t g = idx.order[s];
      if (!pair_q || !pair_g || !pair_pos) return PPS_ERR_INVALID_ARG;
      pair_q[e] = (int32_t)i;
      pair_g[e] = g;
      pair_pos[e] = gallery_cams[g] != query_cams[i] ? 1 : 0;
    }
    pair_off[i + 1] = (int32_t)e;
  }
  return PPS_OK;
}

#include "ctx.cuh"

namespace {
inline void mark(pps_ctx* c, int i, cudaStream_t s) {
  if (c->timing && c->ev_phase[i]) cudaEventRecord(c->ev_phase[i], s);
}
}  // namespace

extern "C" int pps_ctx_create(int device, pps_ctx** out) {
  if (!out) return PPS_ERR_INVALID_ARG;
  *out = nullptr;
  PPS_CUDA_TRY(cudaSetDevice(device));
  pps_ctx* c = new (std::nothrow) pps_ctx();
  if (!c) return PPS_ERR_INVALID_ARG;
  c->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&c->copy_s, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->comp_s, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->side_s, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pairs, cudaEventDisableTiming);
  for (int i = 0; i < kMaxSlabs && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_slab[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_totals, cudaEventDisableTiming);
  for (int i = 0; i <= PPS_N_PHASES && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev_phase[i]);
  if (e != cudaSuccess) {
    pps_ctx_destroy(c);
    return cuda_fail(e, "pps_ctx_create");
  }
  *out = c;
  return PPS_OK;
}

extern "C" int pps_ctx_destroy(pps_ctx* c) {
  if (!c) return PPS_OK;
  cudaSetDevice(c->device);
  if (c->comp_s) cudaStreamSynchronize(c->comp_s);
  if (c->copy_s) cudaStreamSynchronize(c->copy_s);
  if (c->side_s) cudaStreamSynchronize(c->side_s);
  GrowBuf* bufs[] = {&c->qf, &c->gf, &c->qs, &c->gs, &c->qn, &c->gn, &c->dist, &c->ids, &c->pair_ws, &c->pair_off,
                     &c->totals, &c->pair_q, &c->pair_pos, &c->xbuf, &c->counters, &c->ap,
                     &c->valid, &c->first, &c->topk, &c->tki, &c->tkd};
  for (GrowBuf* b : bufs) b->release();
  c->h_small.release();
  c->pass.release();
  for (int i = 0; i < kMaxSlabs; ++i) if (c->ev_slab[i]) cudaEventDestroy(c->ev_slab[i]);
  if (c->ev_totals) cudaEventDestroy(c->ev_totals);
  if (c->ev_in) cudaEventDestroy(c->ev_in);
  if (c->ev_pairs) cudaEventDestroy(c->ev_pairs);
  if (c->side_s) cudaStreamDestroy(c->side_s);
  for (int i = 0; i <= PPS_N_PHASES; ++i) if (c->ev_phase[i]) cudaEventDestroy(c->ev_phase[i]);
  if (c->copy_s) cudaStreamDestroy(c->copy_s);
  if (c->comp_s) cudaStreamDestroy(c->comp_s);
  delete c;
  return PPS_OK;
}

namespace {

int eval_shape(long long nq, long long ng, long long col0, int dim, int precision, int topk, EvalShape* e) {
  if (nq <= 0 || ng <= 0 || dim <= 0 || topk < 0 || topk > PPS_TOPK_MAX || col0 < 0) return PPS_ERR_INVALID_ARG;
  if (nq > 0x7fffffffLL || col0 + ng > 0x7fffffffLL) return PPS_ERR_UNSUPPORTED;
  switch (precision) {
    case PPS_PREC_BF16X1: e->planes = 1; break;
    case PPS_PREC_BF16X3: e->planes = 2; break;
    case PPS_PREC_BF16X6: e->planes = 3; break;
    case PPS_PREC_F16X3: e->planes = 2; break;
    default: return PPS_ERR_INVALID_ARG;
  }
  e->split_planes = precision == PPS_PREC_F16X3 ? (2 | PPS_SPLIT_F16_SCALED) : e->planes;
  e->nq = nq; e->ng = ng; e->col0 = col0; e->dim = dim; e->precision = precision; e->topk = topk;
  e->ldd = (ng + 3) & ~3LL;
  e->kpad = pps_kpad(dim);
  // the distance block is materialised once: bound it (larger galleries go through the chunked
  // two-sweep path of the Python layer, evaluator.RankEngine)
  if ((double)nq * (double)e->ldd * 4.0 > 64.0 * (double)(1LL << 30)) return PPS_ERR_UNSUPPORTED;
  return PPS_OK;
}

int ensure_common(pps_ctx* c, const EvalShape& e) {
  PPS_TRY(c->qs.ensure((size_t)pps_split_bytes(e.nq, e.dim, e.planes)));
  PPS_TRY(c->gs.ensure((size_t)pps_split_bytes(e.ng, e.dim, e.planes)));
  PPS_TRY(c->qn.ensure((size_t)e.nq * 8));      // |x|^2 [rows] (+ the inverse row scales [rows] of PPS_PREC_F16X3)
  PPS_TRY(c->gn.ensure((size_t)e.ng * 8));
  PPS_TRY(c->dist.ensure((size_t)e.nq * e.ldd * 4));
  PPS_TRY(c->pair_ws.ensure((size_t)pps_pairs_workspace_bytes(e.nq, e.ng)));
  PPS_TRY(c->pair_off.ensure(((size_t)e.nq + 1) * 4));
  PPS_TRY(c->totals.ensure(16));
  PPS_TRY(c->ap.ensure((size_t)e.nq * 8));
  PPS_TRY(c->valid.ensure((size_t)e.nq));
  PPS_TRY(c->first.ensure((size_t)e.nq * 4));
  // pinned staging: [0,16) totals | ap[nq] f64 | first[nq] i32 | valid[nq] u8
  const size_t off_ap = 16, off_first = off_ap + (size_t)e.nq * 8, off_valid = off_first + (size_t)e.nq * 4;
  PPS_TRY(c->h_small.ensure(off_valid + (size_t)e.nq));
  unsigned char* base = c->h_small.as<unsigned char>();
  c->st.totals = reinterpret_cast<int32_t*>(base);
  c->st.ap = reinterpret_cast<double*>(base + off_ap);
  c->st.first = reinterpret_cast<int32_t*>(base + off_first);
  c->st.valid = base + off_valid;
  return PPS_OK;
}

// Pair lists, step 1: hits of every query in the LOCAL gallery block.  Single GPU: on the ctx's side stream (ordered
// after `ready_on`, the stream the ids become valid on) so that the pair kernels hide under the distance GEMM.
// Sharded: on the caller's stream, because the caller all-gathers the counts right after.
int pairs_local_count(pps_ctx* c, cudaStream_t ready_on) {
  const EvalShape& e = c->cur;
  cudaStream_t ps = c->pair_stream(ready_on);
  if (ready_on != ps) {
    PPS_CUDA_TRY(cudaEventRecord(c->ev_in, ready_on));
    PPS_CUDA_TRY(cudaStreamWaitEvent(ps, c->ev_in, 0));
  }
  return pps_pairs_local_count(c->d_qid, e.nq, c->d_gid, e.ng, c->pair_ws.p, &c->local_cnt, ps);
}

// step 2: global offsets from the (all-gathered) per-block counts; {n_pairs, max_pairs} on their way to the host
int pairs_offsets(pps_ctx* c, const int32_t* cnt_all, int rank, cudaStream_t cs) {
  const EvalShape& e = c->cur;
  cudaStream_t ps = c->pair_stream(cs);
  PPS_TRY(pps_pairs_offsets(cnt_all, c->world, rank, e.nq, e.ng, c->pair_ws.p, c->pair_off.as<int32_t>(),
                            c->totals.as<int32_t>(), ps));
  PPS_CUDA_TRY(cudaMemcpyAsync(c->st.totals, c->totals.p, 8, cudaMemcpyDeviceToHost, ps));
  PPS_CUDA_TRY(cudaEventRecord(c->ev_totals, ps));
  return PPS_OK;
}

// step 3, after the distance block is enqueued: size + fill the local pairs into their global slots, gather the
// thresholds that live in this block.  counters = [cnt_first (nq) | cnt_le (n_pairs)], zero-filled by the fill.
int pairs_and_thresholds(pps_ctx* c, cudaStream_t cs) {
  const EvalShape& e = c->cur;
  cudaStream_t ps = c->pair_stream(cs);
  PPS_CUDA_TRY(cudaEventSynchronize(c->ev_totals));     // came back while the GEMM is still running
  c->n_pairs = c->st.totals[0];
  c->max_pairs = c->st.totals[1];
  const size_t np1 = (size_t)std::max<long long>(c->n_pairs, 1);
  PPS_TRY(c->pair_q.ensure(np1 * 4));
  PPS_TRY(c->pair_pos.ensure(np1));
  PPS_TRY(c->xbuf.ensure(np1 * 12));
  PPS_TRY(c->counters.ensure(((size_t)e.nq + np1) * 4));
  uint32_t* cnt_first = c->counters.as<uint32_t>();
  uint32_t* cnt_le = cnt_first + e.nq;
  if (c->world > 1) {
    // slots of pairs that live on other shards: zero (they are summed in by the exchange) and marked unfilled
    PPS_CUDA_TRY(cudaMemsetAsync(c->xbuf.p, 0, np1 * 12, ps));
    PPS_CUDA_TRY(cudaMemsetAsync(c->pair_q.p, 0xFF, np1 * 4, ps));
    PPS_CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, ((size_t)e.nq + np1) * 4, ps));
  }
  PPS_TRY(pps_pairs_fill_local(c->d_qid, c->d_qcam, e.nq, c->d_gid, c->d_gcam, e.ng, e.col0, c->pair_ws.p,
                               c->pair_q.as<int32_t>(), c->pair_g(), c->pair_pos.as<uint8_t>(),
                               c->world > 1 ? c->pair_pos32() : nullptr, c->pair_d(), cnt_le, cnt_first, c->n_pairs, ps));
  if (ps != cs) {
    PPS_CUDA_TRY(cudaEventRecord(c->ev_pairs, ps));
    PPS_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_pairs, 0));   // the rank sweeps need the lists; the GEMM did not
  }
  PPS_TRY(pps_rank_gather(c->dist.as<float>(), e.ldd, e.nq, e.ng, e.col0, c->pair_q.as<int32_t>(), c->pair_g(),
                          c->n_pairs, c->pair_d(), cs));
  return PPS_OK;
}

int count_local(pps_ctx* c, cudaStream_t cs) {
  const EvalShape& e = c->cur;
  uint32_t* cnt_first = c->counters.as<uint32_t>();
  if (c->world > 1)    // after the exchange every rank holds all pairs; the kernels read the flags as bytes
    PPS_TRY(pps_pairs_unpack_pos(c->pair_pos32(), c->n_pairs, c->pair_pos.as<uint8_t>(), cs));
  if (e.topk > 0 && c->world == 1) {
    // counts and the k nearest valid items from ONE read of the distance block
    PPS_TRY(c->topk.ensure((size_t)e.nq * e.topk * 8));
    PPS_TRY(pps_topk_init(c->topk.as<uint64_t>(), e.nq, e.topk, cs));
    return pps_rank_sweep(c->dist.as<float>(), e.ldd, e.nq, e.ng, e.col0, c->pair_off.as<int32_t>(), c->pair_g(),
                          c->pair_pos.as<uint8_t>(), c->pair_d(), c->max_pairs, cnt_first + e.nq, cnt_first,
                          c->topk.as<uint64_t>(), e.topk, 1, cs);
  }
  return pps_rank_count(c->dist.as<float>(), e.ldd, e.nq, e.ng, e.col0, c->pair_off.as<int32_t>(), c->pair_g(),
                        c->pair_pos.as<uint8_t>(), c->pair_d(), c->max_pairs, cnt_first + e.nq, cnt_first, cs);
}

// finalize from the (reduced) counters, optional local top-k, results back, the reference's averaging
int finalize_and_fetch(pps_ctx* c, int cmc_topk, cudaStream_t cs, double* out_map, double* out_cmc, double* out_ap,
                       uint8_t* out_valid, int32_t* out_first_rank, int32_t* out_topk_index, float* out_topk_dist) {
  const EvalShape& e = c->cur;
  const long long nq = e.nq;
  const int topk = e.topk;
  const Staging& st = c->st;
  uint32_t* cnt_first = c->counters.as<uint32_t>();
  PPS_TRY(pps_rank_finalize(nq, c->pair_off.as<int32_t>(), c->pair_g(), c->pair_pos.as<uint8_t>(),
                            c->pair_d(), cnt_first + nq, cnt_first, c->ap.as<double>(),
                            c->valid.as<uint8_t>(), c->first.as<int32_t>(), nullptr, cs));
  if (topk > 0) {
    PPS_TRY(c->topk.ensure((size_t)nq * topk * 8));
    PPS_TRY(c->tki.ensure((size_t)nq * topk * 4));
    PPS_TRY(c->tkd.ensure((size_t)nq * topk * 4));
    if (c->world != 1) {       // (single device: count_local already swept the block for counts and top-k together)
      PPS_TRY(pps_topk_init(c->topk.as<uint64_t>(), nq, topk, cs));
      PPS_TRY(pps_topk_update(c->dist.as<float>(), e.ldd, nq, e.ng, e.col0, c->pair_off.as<int32_t>(), c->pair_g(),
                              c->pair_pos.as<uint8_t>(), c->topk.as<uint64_t>(), topk, cs));
    }
    PPS_TRY(pps_topk_unpack(c->topk.as<uint64_t>(), nq, topk, c->tkd.as<float>(), c->tki.as<int32_t>(), cs));
  }
  mark(c, 6, cs);
  PPS_CUDA_TRY(cudaMemcpyAsync(st.ap, c->ap.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, cs));
  PPS_CUDA_TRY(cudaMemcpyAsync(st.first, c->first.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, cs));
  PPS_CUDA_TRY(cudaMemcpyAsync(st.valid, c->valid.p, (size_t)nq, cudaMemcpyDeviceToHost, cs));
  if (topk > 0 && out_topk_index)
    PPS_CUDA_TRY(cudaMemcpyAsync(out_topk_index, c->tki.p, (size_t)nq * topk * 4, cudaMemcpyDeviceToHost, cs));
  if (topk > 0 && out_topk_dist)
    PPS_CUDA_TRY(cudaMemcpyAsync(out_topk_dist, c->tkd.p, (size_t)nq * topk * 4, cudaMemcpyDeviceToHost, cs));
  mark(c, 7, cs);
  PPS_CUDA_TRY(cudaStreamSynchronize(cs));

  // ---- host averaging, as the reference does it (:360-362, :437-438) ----
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

void collect_phase_times(pps_ctx* c) {
  if (!c->timing) return;
  for (int i = 0; i < PPS_N_PHASES; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_phase[i], c->ev_phase[i + 1]) != cudaSuccess) ms = -1.f;
    c->phase_ms[i] = ms;
  }
}

}  // namespace

extern "C" int pps_evaluate_host_ctx(pps_ctx* c, const float* q_feats, long long nq, const float* g_feats, long long ng,
                                     int dim, const int64_t* query_ids, const int64_t* query_cams,
                                     const int64_t* gallery_ids, const int64_t* gallery_cams, int precision,
                                     int cmc_topk, int topk, double* out_map, double* out_cmc, double* out_ap,
                                     uint8_t* out_valid, int32_t* out_first_rank, int32_t* out_topk_index,
                                     float* out_topk_dist) {
  if (!c || cmc_topk < 0) return PPS_ERR_INVALID_ARG;
  PPS_TRY(eval_shape(nq, ng, 0, dim, precision, topk, &c->cur));
  c->world = 1;
  c->ext_pair_s = nullptr;
  const EvalShape& e = c->cur;
  if (!q_feats || !g_feats || !query_ids || !query_cams || !gallery_ids || !gallery_cams) return PPS_ERR_INVALID_ARG;
  if (!out_map || (cmc_topk > 0 && !out_cmc)) return PPS_ERR_INVALID_ARG;
  PPS_CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t cs = c->comp_s, ps = c->copy_s, ss = c->side_s;
  const int planes = e.planes, kpad = e.kpad;
  const long long ldd = e.ldd;

  PPS_TRY(ensure_common(c, e));
  PPS_TRY(c->qf.ensure((size_t)nq * dim * 4));
  PPS_TRY(c->gf.ensure((size_t)ng * dim * 4));
  PPS_TRY(c->ids.ensure((size_t)(2 * nq + 2 * ng) * 8));
  int64_t* d_qid = c->ids.as<int64_t>();
  int64_t* d_qcam = d_qid + nq;
  int64_t* d_gid = d_qcam + nq;
  int64_t* d_gcam = d_gid + ng;
  c->d_qid = d_qid; c->d_qcam = d_qcam; c->d_gid = d_gid; c->d_gcam = d_gcam;

  // ---- ids up, pair counts + offsets, totals back: all on the side stream, hidden under the feature upload ----
  PPS_CUDA_TRY(cudaMemcpyAsync(d_qid, query_ids, (size_t)nq * 8, cudaMemcpyHostToDevice, ss));
  PPS_CUDA_TRY(cudaMemcpyAsync(d_qcam, query_cams, (size_t)nq * 8, cudaMemcpyHostToDevice, ss));
  PPS_CUDA_TRY(cudaMemcpyAsync(d_gid, gallery_ids, (size_t)ng * 8, cudaMemcpyHostToDevice, ss));
  PPS_CUDA_TRY(cudaMemcpyAsync(d_gcam, gallery_cams, (size_t)ng * 8, cudaMemcpyHostToDevice, ss));
  PPS_TRY(pairs_local_count(c, ss));
  PPS_TRY(pairs_offsets(c, c->local_cnt, 0, ss));

  // ---- queries up + split ----
  PPS_CUDA_TRY(cudaMemcpyAsync(c->qf.p, q_feats, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, cs));
  PPS_TRY(pps_split_rows(c->qf.p, PPS_DTYPE_F32, nq, dim, dim, e.split_planes, c->qs.p, c->qn.as<float>(), cs));

  // ---- gallery in row slabs: H2D on the copy stream; split + distance of slab s overlap the copy of s+1 ----
  long long slab = ((ng + 7) / 8 + 255) & ~255LL;              // ~8 slabs, whole 256-column tiles
  if (slab < 1024) slab = 1024;
  while ((ng + slab - 1) / slab > kMaxSlabs) slab *= 2;
  int si = 0;
  const size_t esz = 2;   // bf16 planes
  for (long long r0 = 0; r0 < ng; r0 += slab, ++si) {
    const long long nr = std::min(slab, ng - r0);
    PPS_CUDA_TRY(cudaMemcpyAsync(c->gf.as<float>() + r0 * dim, g_feats + r0 * dim, (size_t)nr * dim * 4,
                                 cudaMemcpyHostToDevice, ps));
    PPS_CUDA_TRY(cudaEventRecord(c->ev_slab[si], ps));
    PPS_CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_slab[si], 0));
    PPS_TRY(pps_split_rows_slab(c->gf.p, PPS_DTYPE_F32, r0, nr, ng, dim, dim, e.split_planes, c->gs.p, c->gn.as<float>(), cs));
    PPS_TRY(pps_dist_tc(c->qs.p, c->qn.as<float>(), nq, planes, 0,
                        c->gs.as<unsigned char>() + (size_t)r0 * kpad * esz, c->gn.as<float>() + r0, nr, planes, ng, dim,
                        precision, 0, c->dist.as<float>() + r0, ldd, cs));
  }
  int rc = pairs_and_thresholds(c, cs);
  if (rc == PPS_OK) rc = count_local(c, cs);
  if (rc == PPS_OK)
    rc = finalize_and_fetch(c, cmc_topk, cs, out_map, out_cmc, out_ap, out_valid, out_first_rank, out_topk_index,
                            out_topk_dist);
  cudaStreamSynchronize(ps);
  return rc;
}

// ------------------------------------------------------------------------------------
// Resident evaluation in steps, so that a gallery sharded over several GPUs can put its exchanges in between:
//   pps_rank_begin        hits of each query in the LOCAL block; operand split + distance of the local block enqueued
//                                                                      -> all-gather *d_local_cnt over ranks [nq int32 each]
//   pps_rank_thresholds   global pair offsets, local pairs into their global slots, their thresholds
//                                                                      -> all-reduce(SUM, int32) *d_exchange [*n_words]
//   pps_rank_count_local  counting sweep over the local block          -> all-reduce(SUM) *d_counters [nq + n_pairs u32]
//   pps_rank_end          finalize, (local) top-k, results to the host, the reference's averaging
// With world == 1 the exchanges are skipped and the pair kernels run on a side stream under the GEMM.
// ------------------------------------------------------------------------------------
extern "C" int pps_rank_begin(pps_ctx* c, const float* d_q, long long nq, const float* d_g, long long ng_local, int dim,
                              const int64_t* d_qid, const int64_t* d_qcam, const int64_t* d_gid_local,
                              const int64_t* d_gcam_local, long long gallery_offset, int world, int precision, int topk,
                              void* stream, void* pair_stream, int32_t** d_local_cnt) {
  if (!c || world < 1) return PPS_ERR_INVALID_ARG;
  PPS_TRY(eval_shape(nq, ng_local, gallery_offset, dim, precision, topk, &c->cur));
  if (!d_q || !d_g || !d_qid || !d_qcam || !d_gid_local || !d_gcam_local) return PPS_ERR_INVALID_ARG;
  PPS_CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  c->d_qid = d_qid; c->d_qcam = d_qcam; c->d_gid = d_gid_local; c->d_gcam = d_gcam_local;
  c->d_q = d_q; c->d_g = d_g;
  c->world = world;
  c->ext_pair_s = static_cast<cudaStream_t>(pair_stream);
  PPS_TRY(ensure_common(c, c->cur));
  const EvalShape& e = c->cur;
  mark(c, 0, cs);
  PPS_TRY(pairs_local_count(c, cs));
  if (d_local_cnt) *d_local_cnt = c->local_cnt;
  mark(c, 1, cs);
  // the distance block is enqueued NOW, so that the caller's all-gather (host latency included) hides under it
  PPS_TRY(pps_split_rows(d_q, PPS_DTYPE_F32, e.nq, e.dim, e.dim, e.split_planes, c->qs.p, c->qn.as<float>(), cs));
  PPS_TRY(pps_split_rows(d_g, PPS_DTYPE_F32, e.ng, e.dim, e.dim, e.split_planes, c->gs.p, c->gn.as<float>(), cs));
  mark(c, 2, cs);
  PPS_TRY(pps_dist_tc(c->qs.p, c->qn.as<float>(), e.nq, e.planes, 0, c->gs.p, c->gn.as<float>(), e.ng, e.planes, 0,
                      e.dim, e.precision, world > 1 ? PPS_DIST_RESERVE_SM_PAIR : 0, c->dist.as<float>(), e.ldd, cs));
  mark(c, 3, cs);
  return PPS_OK;
}

extern "C" int pps_rank_thresholds(pps_ctx* c, const int32_t* d_cnt_all, int rank, void* stream, long long* n_pairs,
                                   int32_t** d_exchange, long long* n_words) {
  if (!c || c->cur.nq <= 0 || rank < 0 || rank >= c->world) return PPS_ERR_INVALID_ARG;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  PPS_TRY(pairs_offsets(c, d_cnt_all ? d_cnt_all : c->local_cnt, rank, cs));
  PPS_TRY(pairs_and_thresholds(c, cs));
  mark(c, 4, cs);
  if (n_pairs) *n_pairs = c->n_pairs;
  if (d_exchange) *d_exchange = c->xbuf.as<int32_t>();
  if (n_words) *n_words = 3 * c->n_pairs;
  return PPS_OK;
}

extern "C" int pps_rank_count_local(pps_ctx* c, void* stream, uint32_t** d_counters, long long* n_counters) {
  if (!c || c->cur.nq <= 0) return PPS_ERR_INVALID_ARG;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  PPS_TRY(count_local(c, cs));
  mark(c, 5, cs);
  if (d_counters) *d_counters = c->counters.as<uint32_t>();
  if (n_counters) *n_counters = c->cur.nq + c->n_pairs;
  return PPS_OK;
}

extern "C" int pps_rank_end(pps_ctx* c, int cmc_topk, void* stream, double* out_map, double* out_cmc, double* out_ap,
                            uint8_t* out_valid, int32_t* out_first_rank, int32_t* out_topk_index,
                            float* out_topk_dist) {
  if (!c || c->cur.nq <= 0 || cmc_topk < 0) return PPS_ERR_INVALID_ARG;
  const int rc = finalize_and_fetch(c, cmc_topk, static_cast<cudaStream_t>(stream), out_map, out_cmc, out_ap, out_valid,
                                    out_first_rank, out_topk_index, out_topk_dist);
  if (rc == PPS_OK || rc == PPS_ERR_NO_VALID_QUERY) collect_phase_times(c);
  return rc;
}

// Same evaluation with everything already RESIDENT on one device: what bench.py times as `value` at N = 1.
extern "C" int pps_evaluate_device_ctx(pps_ctx* c, const float* d_q, long long nq, const float* d_g, long long ng,
                                       int dim, const int64_t* d_qid, const int64_t* d_qcam, const int64_t* d_gid,
                                       const int64_t* d_gcam, int precision, int cmc_topk, int topk, void* stream,
                                       double* out_map, double* out_cmc, double* out_ap, uint8_t* out_valid,
                                       int32_t* out_first_rank, int32_t* out_topk_index, float* out_topk_dist) {
  if (!out_map || (cmc_topk > 0 && !out_cmc)) return PPS_ERR_INVALID_ARG;
  PPS_TRY(pps_rank_begin(c, d_q, nq, d_g, ng, dim, d_qid, d_qcam, d_gid, d_gcam, 0, 1, precision, topk, stream, nullptr,
                         nullptr));
  PPS_TRY(pps_rank_thresholds(c, nullptr, 0, stream, nullptr, nullptr, nullptr));
  PPS_TRY(pps_rank_count_local(c, stream, nullptr, nullptr));
  return pps_rank_end(c, cmc_topk, stream, out_map, out_cmc, out_ap, out_valid, out_first_rank, out_topk_index,
                      out_topk_dist);
}

extern "C" int pps_ctx_set_timing(pps_ctx* c, int enabled) {
  if (!c) return PPS_ERR_INVALID_ARG;
  c->timing = enabled != 0;
  return PPS_OK;
}

extern "C" int pps_ctx_phase_ms(const pps_ctx* c, float* out_ms) {
  if (!c || !out_ms) return PPS_ERR_INVALID_ARG;
  for (int i = 0; i < PPS_N_PHASES; ++i) out_ms[i] = c->phase_ms[i];
  return PPS_OK;
}

extern "C" int pps_evaluate_host(const float* q_feats, long long nq, const float* g_feats, long long ng, int dim,
                                 const int64_t* query_ids, const int64_t* query_cams, const int64_t* gallery_ids,
                                 const int64_t* gallery_cams, int precision, int cmc_topk, int topk, int device,
                                 double* out_map, double* out_cmc, double* out_ap, uint8_t* out_valid,
                                 int32_t* out_first_rank, int32_t* out_topk_index, float* out_topk_dist) {
  pps_ctx* c = nullptr;
  int rc = pps_ctx_create(device, &c);
  if (rc != PPS_OK) return rc;
  rc = pps_evaluate_host_ctx(c, q_feats, nq, g_feats, ng, dim, query_ids, query_cams, gallery_ids, gallery_cams,
                             precision, cmc_topk, topk, out_map, out_cmc, out_ap, out_valid, out_first_rank,
                             out_topk_index, out_topk_dist);
  pps_ctx_destroy(c);
  return rc;
}

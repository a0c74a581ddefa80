// Host-side pieces of the C ABI: error text, launch counter, same-id pair lists, and the
// whole-evaluation entry point that works from host buffers (pps_evaluate_host).
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
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

// ------------------------------------------------------------------------------------
// pps_evaluate_host
// ------------------------------------------------------------------------------------
namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc"); }
    return PPS_OK;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct StreamGuard {
  cudaStream_t s = nullptr;
  ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};
struct EventGuard {
  cudaEvent_t e = nullptr;
  ~EventGuard() { if (e) cudaEventDestroy(e); }
};

#define PPS_TRY(expr) do { int _rc = (expr); if (_rc != PPS_OK) return _rc; } while (0)

}  // namespace

extern "C" int pps_evaluate_host(const float* q_feats, long long nq, const float* g_feats, long long ng, int dim,
                                 const int64_t* query_ids, const int64_t* query_cams, const int64_t* gallery_ids,
                                 const int64_t* gallery_cams, int precision, int cmc_topk, int topk, int device,
                                 double* out_map, double* out_cmc, double* out_ap, uint8_t* out_valid,
                                 int32_t* out_first_rank, int32_t* out_topk_index, float* out_topk_dist) {
  if (nq <= 0 || ng <= 0 || dim <= 0 || cmc_topk < 0 || topk < 0 || topk > PPS_TOPK_MAX) return PPS_ERR_INVALID_ARG;
  if (!q_feats || !g_feats || !query_ids || !query_cams || !gallery_ids || !gallery_cams) return PPS_ERR_INVALID_ARG;
  if (!out_map || (cmc_topk > 0 && !out_cmc)) return PPS_ERR_INVALID_ARG;
  int planes;
  switch (precision) {
    case PPS_PREC_BF16X1: planes = 1; break;
    case PPS_PREC_BF16X3: planes = 2; break;
    case PPS_PREC_BF16X6: planes = 3; break;
    default: return PPS_ERR_INVALID_ARG;
  }
  PPS_CUDA_TRY(cudaSetDevice(device));

  // ---- host: same-id pair lists (tiny next to the feature upload) ----
  const long long n_pairs = pps_pairs_count(query_ids, nq, gallery_ids, ng);
  if (n_pairs < 0) return (int)n_pairs;
  std::vector<int32_t> pair_off((size_t)nq + 1), pair_q((size_t)n_pairs), pair_g((size_t)n_pairs);
  std::vector<uint8_t> pair_pos((size_t)n_pairs);
  PPS_TRY(pps_pairs_fill(query_ids, query_cams, nq, gallery_ids, gallery_cams, ng, pair_off.data(), pair_q.data(),
                         pair_g.data(), pair_pos.data()));
  int max_pairs = 0;
  std::vector<int32_t> junk_off((size_t)nq + 1, 0), junk_g;
  for (long long i = 0; i < nq; ++i) {
    max_pairs = std::max(max_pairs, pair_off[i + 1] - pair_off[i]);
    for (int e = pair_off[i]; e < pair_off[i + 1]; ++e)
      if (!pair_pos[e]) junk_g.push_back(pair_g[e]);
    junk_off[i + 1] = (int32_t)junk_g.size();
  }

  // the distance block is materialised once: bound it (larger galleries go through the chunked
  // two-sweep path of the Python layer, evaluator.rank_eval)
  const long long ldd = (ng + 3) & ~3LL;
  if ((double)nq * (double)ldd * 4.0 > 64.0 * (double)(1LL << 30)) return PPS_ERR_UNSUPPORTED;

  StreamGuard copy_s, comp_s;
  PPS_CUDA_TRY(cudaStreamCreateWithFlags(&copy_s.s, cudaStreamNonBlocking));
  PPS_CUDA_TRY(cudaStreamCreateWithFlags(&comp_s.s, cudaStreamNonBlocking));

  DevBuf d_qf, d_gf, d_qs, d_gs, d_qn, d_gn, d_dist, d_poff, d_pq, d_pg, d_ppos, d_pd, d_cle, d_cfirst, d_ap, d_valid,
      d_first, d_joff, d_jg, d_topk, d_tki, d_tkd;
  PPS_TRY(d_qf.alloc((size_t)nq * dim * 4));
  PPS_TRY(d_gf.alloc((size_t)ng * dim * 4));
  PPS_TRY(d_qs.alloc((size_t)pps_split_bytes(nq, dim, planes)));
  PPS_TRY(d_gs.alloc((size_t)pps_split_bytes(ng, dim, planes)));
  PPS_TRY(d_qn.alloc((size_t)nq * 4));
  PPS_TRY(d_gn.alloc((size_t)ng * 4));
  PPS_TRY(d_dist.alloc((size_t)nq * ldd * 4));
  PPS_TRY(d_poff.alloc(((size_t)nq + 1) * 4));
  PPS_TRY(d_pq.alloc((size_t)n_pairs * 4));
  PPS_TRY(d_pg.alloc((size_t)n_pairs * 4));
  PPS_TRY(d_ppos.alloc((size_t)n_pairs));
  PPS_TRY(d_pd.alloc((size_t)n_pairs * 4));
  PPS_TRY(d_cle.alloc((size_t)n_pairs * 4));
  PPS_TRY(d_cfirst.alloc((size_t)nq * 4));
  PPS_TRY(d_ap.alloc((size_t)nq * 8));
  PPS_TRY(d_valid.alloc((size_t)nq));
  PPS_TRY(d_first.alloc((size_t)nq * 4));
  if (topk > 0) {
    PPS_TRY(d_joff.alloc(((size_t)nq + 1) * 4));
    PPS_TRY(d_jg.alloc(junk_g.size() * 4));
    PPS_TRY(d_topk.alloc((size_t)nq * topk * 8));
    PPS_TRY(d_tki.alloc((size_t)nq * topk * 4));
    PPS_TRY(d_tkd.alloc((size_t)nq * topk * 4));
  }

  // ---- uploads: queries + pair lists on the compute stream, gallery in row chunks on the copy stream ----
  PPS_CUDA_TRY(cudaMemcpyAsync(d_qf.p, q_feats, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, comp_s.s));
  PPS_CUDA_TRY(cudaMemcpyAsync(d_poff.p, pair_off.data(), ((size_t)nq + 1) * 4, cudaMemcpyHostToDevice, comp_s.s));
  if (n_pairs > 0) {
    PPS_CUDA_TRY(cudaMemcpyAsync(d_pq.p, pair_q.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, comp_s.s));
    PPS_CUDA_TRY(cudaMemcpyAsync(d_pg.p, pair_g.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, comp_s.s));
    PPS_CUDA_TRY(cudaMemcpyAsync(d_ppos.p, pair_pos.data(), (size_t)n_pairs, cudaMemcpyHostToDevice, comp_s.s));
  }
  PPS_CUDA_TRY(cudaMemsetAsync(d_pd.p, 0, std::max<size_t>(16, (size_t)n_pairs * 4), comp_s.s));
  PPS_CUDA_TRY(cudaMemsetAsync(d_cle.p, 0, std::max<size_t>(16, (size_t)n_pairs * 4), comp_s.s));
  PPS_CUDA_TRY(cudaMemsetAsync(d_cfirst.p, 0, (size_t)nq * 4, comp_s.s));
  if (topk > 0) {
    PPS_CUDA_TRY(cudaMemcpyAsync(d_joff.p, junk_off.data(), ((size_t)nq + 1) * 4, cudaMemcpyHostToDevice, comp_s.s));
    if (!junk_g.empty())
      PPS_CUDA_TRY(cudaMemcpyAsync(d_jg.p, junk_g.data(), junk_g.size() * 4, cudaMemcpyHostToDevice, comp_s.s));
    PPS_TRY(pps_topk_init(d_topk.as<uint64_t>(), nq, topk, comp_s.s));
  }
  PPS_TRY(pps_split_rows(d_qf.p, PPS_DTYPE_F32, nq, dim, dim, planes, d_qs.p, d_qn.as<float>(), comp_s.s));

  // gallery upload + split in row slabs so the H2D copy overlaps the split kernels
  const long long slab = std::max<long long>(1024, (64LL << 20) / ((long long)dim * 4));
  std::vector<EventGuard> evs((size_t)((ng + slab - 1) / slab));
  size_t ei = 0;
  for (long long r0 = 0; r0 < ng; r0 += slab, ++ei) {
    const long long nr = std::min(slab, ng - r0);
    PPS_CUDA_TRY(cudaMemcpyAsync(d_gf.as<float>() + r0 * dim, g_feats + r0 * dim, (size_t)nr * dim * 4,
                                 cudaMemcpyHostToDevice, copy_s.s));
    PPS_CUDA_TRY(cudaEventCreateWithFlags(&evs[ei].e, cudaEventDisableTiming));
    PPS_CUDA_TRY(cudaEventRecord(evs[ei].e, copy_s.s));
    PPS_CUDA_TRY(cudaStreamWaitEvent(comp_s.s, evs[ei].e, 0));
    // all slabs write into the one [planes][ng][kpad] buffer
    PPS_TRY(pps_split_rows_slab(d_gf.p, PPS_DTYPE_F32, r0, nr, ng, dim, dim, planes, d_gs.p, d_gn.as<float>(),
                                comp_s.s));
  }

  PPS_TRY(pps_dist_tc(d_qs.p, d_qn.as<float>(), nq, planes, d_gs.p, d_gn.as<float>(), ng, planes, dim, precision, 0,
                      d_dist.as<float>(), ldd, comp_s.s));
  PPS_TRY(pps_rank_gather(d_dist.as<float>(), ldd, nq, ng, 0, d_pq.as<int32_t>(), d_pg.as<int32_t>(), n_pairs,
                          d_pd.as<float>(), comp_s.s));
  PPS_TRY(pps_rank_count(d_dist.as<float>(), ldd, nq, ng, 0, d_poff.as<int32_t>(), d_pg.as<int32_t>(),
                         d_ppos.as<uint8_t>(), d_pd.as<float>(), max_pairs, d_cle.as<uint32_t>(),
                         d_cfirst.as<uint32_t>(), comp_s.s));
  PPS_TRY(pps_rank_finalize(nq, d_poff.as<int32_t>(), d_pg.as<int32_t>(), d_ppos.as<uint8_t>(), d_pd.as<float>(),
                            d_cle.as<uint32_t>(), d_cfirst.as<uint32_t>(), d_ap.as<double>(), d_valid.as<uint8_t>(),
                            d_first.as<int32_t>(), nullptr, comp_s.s));
  if (topk > 0) {
    PPS_TRY(pps_topk_update(d_dist.as<float>(), ldd, nq, ng, 0, d_joff.as<int32_t>(), d_jg.as<int32_t>(),
                            d_topk.as<uint64_t>(), topk, comp_s.s));
    PPS_TRY(pps_topk_unpack(d_topk.as<uint64_t>(), nq, topk, d_tkd.as<float>(), d_tki.as<int32_t>(), comp_s.s));
  }

  // ---- results back ----
  std::vector<double> ap((size_t)nq);
  std::vector<uint8_t> valid((size_t)nq);
  std::vector<int32_t> first((size_t)nq);
  PPS_CUDA_TRY(cudaMemcpyAsync(ap.data(), d_ap.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, comp_s.s));
  PPS_CUDA_TRY(cudaMemcpyAsync(valid.data(), d_valid.p, (size_t)nq, cudaMemcpyDeviceToHost, comp_s.s));
  PPS_CUDA_TRY(cudaMemcpyAsync(first.data(), d_first.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, comp_s.s));
  if (topk > 0 && out_topk_index)
    PPS_CUDA_TRY(cudaMemcpyAsync(out_topk_index, d_tki.p, (size_t)nq * topk * 4, cudaMemcpyDeviceToHost, comp_s.s));
  if (topk > 0 && out_topk_dist)
    PPS_CUDA_TRY(cudaMemcpyAsync(out_topk_dist, d_tkd.p, (size_t)nq * topk * 4, cudaMemcpyDeviceToHost, comp_s.s));
  PPS_CUDA_TRY(cudaStreamSynchronize(comp_s.s));
  PPS_CUDA_TRY(cudaStreamSynchronize(copy_s.s));

  // ---- host averaging, as the reference does it (:360-362, :437-438) ----
  double ap_sum = 0.0;
  long long n_valid = 0;
  std::vector<double> hist((size_t)std::max(cmc_topk, 1), 0.0);
  for (long long i = 0; i < nq; ++i) {
    if (!valid[i]) continue;
    ++n_valid;
    ap_sum += ap[i];
    if (first[i] >= 0 && first[i] < cmc_topk) hist[(size_t)first[i]] += 1.0;
  }
  if (out_ap) std::memcpy(out_ap, ap.data(), (size_t)nq * 8);
  if (out_valid) std::memcpy(out_valid, valid.data(), (size_t)nq);
  if (out_first_rank) std::memcpy(out_first_rank, first.data(), (size_t)nq * 4);
  if (n_valid == 0) return PPS_ERR_NO_VALID_QUERY;
  *out_map = ap_sum / (double)n_valid;
  double run = 0.0;
  for (int k = 0; k < cmc_topk; ++k) {
    run += hist[(size_t)k];
    out_cmc[k] = run / (double)n_valid;
  }
  return PPS_OK;
}

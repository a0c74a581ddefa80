// The evaluation context shared by the host-side entry points (c_api.cu: pps_evaluate_*, pps_rank_*; pass.cu: pps_pass_*):
// streams, events and grow-only device / pinned scratch, so that repeated evaluations allocate nothing.
#pragma once

#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace pps {

struct GrowBuf {                       // device buffer that only ever grows
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return PPS_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return ::pps::cuda_fail(e, "cudaMalloc"); }
    cap = want;
    return PPS_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {                        // pinned host staging, grow-only
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return PPS_OK;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMallocHost(&p, bytes + 256);
    if (e != cudaSuccess) { p = nullptr; return ::pps::cuda_fail(e, "cudaMallocHost"); }
    cap = bytes + 256;
    return PPS_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

constexpr int kMaxSlabs = 64;

#define PPS_TRY(expr) do { int _rc = (expr); if (_rc != PPS_OK) return _rc; } while (0)

struct EvalShape {
  long long nq, ng /*rows of the local gallery block (and of its id arrays)*/, col0 /*its first global row*/, ldd;
  int dim, kpad, planes, precision, topk;
  int split_planes;      // `planes` argument of pps_split_rows: planes, or 2 | PPS_SPLIT_F16_SCALED for PPS_PREC_F16X3
};
struct Staging {   // pinned: totals, per-query results
  int32_t* totals; double* ap; int32_t* first; uint8_t* valid;
};

// Device-count forms of a few entry points (pairs.cu, rank.cu), used by the speculative pass: sizes are host UPPER BOUNDS,
// the actual counts are read on the device, so nothing has to come back to the host mid-pass.
int pairs_count_device_ex(const int64_t* query_ids, long long nq, const int64_t* gallery_ids, long long ng_cap, void* workspace,
                          int32_t* pair_off, int32_t* totals, void* stream, const int32_t* ng_dev);
int pairs_fill_device_ex(const int64_t* query_ids, const int64_t* query_cams, long long nq, const int64_t* gallery_ids,
                         const int64_t* gallery_cams, long long ng_cap, const void* workspace, int32_t* pair_q, int32_t* pair_g,
                         uint8_t* pair_pos, float* zero_f32, uint32_t* zero_u32, uint32_t* zero_per_query, long long capacity,
                         void* stream, const int32_t* ng_dev);
int pairs_remap_ex(int32_t* pair_g, long long n_cap, const int32_t* cand_rows, long long offset, void* stream,
                   const int32_t* n_dev);
int pairs_compact_rows_ex(const int64_t* query_ids, long long nq, const int32_t* pair_off, const int32_t* pair_q,
                          const int32_t* pair_g, long long n_pairs, long long row_lo, long long row_hi, void* workspace,
                          int32_t* gp_rows, int32_t* pair_col, int32_t* n_rows, void* stream, const int32_t* n_pairs_dev,
                          long long rows_cap);
int rank_gather_ex(const float* dist, long long ldd, long long nq, long long ncols, long long col0, const int32_t* pair_q,
                   const int32_t* pair_g, long long n_pairs, float* pair_d, void* stream, const int32_t* n_dev);

// State of a multi-block / sharded pass (pass.cu: pps_pass_begin .. pps_pass_end)
constexpr int kPassMaxBlocks = 4096;
constexpr int kPassTimedLaunches = 256;
struct PassState {
  bool active = false;
  long long nq = 0, ngl = 0, ng_global = 0, offset = 0, chunk = 0, ldd = 0, n_pairs = 0, n_cand = 0, n_rows = 0;
  int dim = 0, kpad = 0, dtype = 0, planes = 0, split_planes = 0, precision = 0, topk = 0, world = 1, rank = 0, flags = 0;
  int max_pairs = 0, n_blocks = 0, tk_cap = 2048;
  bool prefilter = false, g_inplace = false, epi_topk = false;
  bool fused = false;            // blocks after the first: counting (+ admission) in the distance epilogue, nothing written
  int p_cap = 0;                 // thresholds per query in the tables of the counting epilogue
  const void *d_q = nullptr, *d_g = nullptr;
  const void *h_q = nullptr, *h_g = nullptr;      // host sources of the NEXT pass (pps_pass_set_host_input), else null
  bool g_from_host = false;
  const int64_t *d_qid = nullptr, *d_qcam = nullptr, *d_gid = nullptr, *d_gcam = nullptr;
  long long blk_row0[kPassMaxBlocks], blk_rows[kPassMaxBlocks];
  size_t packed_bytes = 0;
  size_t res_bytes = 0, res_first_off = 0, res_valid_off = 0, res_flags_off = 0;   // layout of the per-query result buffer (ap)
  // buffers
  GrowBuf qs, qn, gs, gn, dist, tdist, ts, tn, pair_ws, pair_off, totals, pair_q, pair_g, pair_pos, pair_d, packed, pf_ws,
      cand_rows, cand_gid, cand_gcam, gp_rows, pair_col, gp_ws, tk_bound, tk_cnt, tk_cand, small, ap, valid, first, tki, tkd,
      thr_tab, tpair_tab, cnt_tab, dstar, gstar;
  cudaEvent_t ev_a = nullptr, ev_rows = nullptr;
  cudaEvent_t ev_t[kPassTimedLaunches][2] = {};
  int n_timed = 0, timed_kind[kPassTimedLaunches] = {};
  float host_wait_ms[4] = {};
  // sizes of the last pass that read its counts back (the key they belong to), used as upper bounds by the next one
  struct Hints { bool valid = false; long long nq = 0, ng_global = 0, ngl = 0, offset = 0, mbb = 0; int dim = 0, dtype = 0, topk = 0,
                 world = 0, precision = 0; long long n_cand = 0, n_pairs = 0, n_rows = 0; int max_pairs = 0; } hints;
  bool speculative = false;
  long long cap_cand = 0, cap_rows = 0;
  void release() {
    GrowBuf* bufs[] = {&qs, &qn, &gs, &gn, &dist, &tdist, &ts, &tn, &pair_ws, &pair_off, &totals, &pair_q, &pair_g, &pair_pos,
                       &pair_d, &packed, &pf_ws, &cand_rows, &cand_gid, &cand_gcam, &gp_rows, &pair_col, &gp_ws, &tk_bound,
                       &tk_cnt, &tk_cand, &small, &ap, &valid, &first, &tki, &tkd, &thr_tab, &tpair_tab, &cnt_tab, &dstar, &gstar};
    for (GrowBuf* b : bufs) b->release();
    if (ev_a) cudaEventDestroy(ev_a);
    if (ev_rows) cudaEventDestroy(ev_rows);
    ev_a = ev_rows = nullptr;
    for (auto& e : ev_t) for (auto& x : e) { if (x) cudaEventDestroy(x); x = nullptr; }
  }
  // packed exchange buffer of the pass: [top-k keys nq*k u64 | cnt_first nq u32 | cnt_le n_pairs u32 | flags 2 u32]
  unsigned long long* keys() const { return packed.as<unsigned long long>(); }
  uint32_t* cnt_first() const { return reinterpret_cast<uint32_t*>(packed.as<unsigned char>() + (size_t)nq * topk * 8); }
  uint32_t* cnt_le() const { return cnt_first() + nq; }
  uint32_t* flags_dev() const { return cnt_le() + n_pairs; }
  size_t counter_words() const { return (size_t)nq + (size_t)n_pairs + 2; }
};

}  // namespace pps

struct pps_ctx {
  int device = 0;
  pps::PassState pass;
  cudaStream_t copy_s = nullptr, comp_s = nullptr, side_s = nullptr;
  cudaEvent_t ev_slab[pps::kMaxSlabs] = {};
  cudaEvent_t ev_totals = nullptr, ev_in = nullptr, ev_pairs = nullptr;
  pps::GrowBuf qf, gf, qs, gs, qn, gn, dist, ids, pair_ws, pair_off, totals, pair_q, pair_pos, xbuf, counters,
      ap, valid, first, topk, tki, tkd;
  pps::PinBuf h_small;      // totals + per-query results
  // state of the evaluation in flight (pps_rank_begin .. pps_rank_end)
  pps::EvalShape cur = {};
  pps::Staging st = {};
  const int64_t *d_qid = nullptr, *d_qcam = nullptr, *d_gid = nullptr, *d_gcam = nullptr;
  const float *d_q = nullptr, *d_g = nullptr;
  long long n_pairs = 0;
  int max_pairs = 0;
  int world = 1;
  int32_t* local_cnt = nullptr;                 // inside pair_ws
  // exchange buffer xbuf = [pair_d bits (E) | pair_g (E) | pair_pos as int32 (E)], E = n_pairs
  float* pair_d() const { return xbuf.as<float>(); }
  int32_t* pair_g() const { return xbuf.as<int32_t>() + n_pairs; }
  int32_t* pair_pos32() const { return xbuf.as<int32_t>() + 2 * n_pairs; }
  cudaStream_t ext_pair_s = nullptr;            // caller-provided stream for the pair kernels (sharded runs)
  cudaStream_t pair_stream(cudaStream_t cs) const { return ext_pair_s ? ext_pair_s : (world == 1 ? side_s : cs); }
  // optional phase timing of pps_evaluate_device_ctx (events on the caller's stream)
  bool timing = false;
  cudaEvent_t ev_phase[PPS_N_PHASES + 1] = {};
  float phase_ms[PPS_N_PHASES] = {};
};


// TEST-ONLY library (pps_b200/_C/libpps_b200_testhooks.so, built by pps_b200.build.build_test_hooks()): instrumentation the
// CPU tests use, kept out of the product ABI (include/pps_b200.h declares none of it).
#include "dist_tiles.cuh"

using namespace pps;

// Instrumentation: the tile schedule of the 2-CTA kernel, replayed on the host (the same TileWalk the producer, MMA and
// epilogue warps run).  out[i] = {group, m tile, n tile | run_start << 30 | run_end << 31} for the i-th tile of CTA pair
// `pair` out of `npairs`; returns the number of tiles (also when it exceeds cap).  rank_order != 0: the EPI_RANK walk.
extern "C" long long pps_test_tile_walk(int rank_order, int m_tiles, int n_tiles, int groups, long long npairs,
                                         long long pair, int32_t* out, long long cap) {
  if (m_tiles < 0 || n_tiles < 0 || groups < 1 || npairs < 1 || pair < 0 || pair >= npairs) return PPS_ERR_INVALID_ARG;
  Gemm2Args ga{};
  ga.g.m_tiles = m_tiles; ga.g.n_tiles = n_tiles; ga.groups = groups;
  long long n = 0;
  auto emit = [&](long long grp, int m, int nt, bool rs, bool re) {
    if (out && n < cap) {
      out[3 * n + 0] = (int32_t)grp;
      out[3 * n + 1] = m;
      out[3 * n + 2] = (int32_t)((uint32_t)nt | (rs ? 0x40000000u : 0u) | (re ? 0x80000000u : 0u));
    }
    ++n;
  };
  if (rank_order) {
    TileWalk<EPI_RANK> w(ga, pair, npairs);
    while (w.next()) emit(w.grp, w.m_tile, w.n_tile, w.run_start, w.run_end);
  } else {
    TileWalk<EPI_DIST> w(ga, pair, npairs);
    while (w.next()) emit(w.grp, w.m_tile, w.n_tile, false, false);
  }
  return n;
}

// plan.h -- host-side schedule of a fused top-k search (shared by api.cu and the C test shim).
//
// A search never materialises the [Q, N] score matrix (code/search_image.py:107 + :109 would).
// Instead the gallery's row tiles are visited in a few nested, strided phases:
//
//   phase 0 ("dense")  tiles 0, s0, 2*s0, ...      every score becomes a key; exact top-k of
//                                                 this sample gives a lower bound thr[q] on the
//                                                 global k-th best score (a sample's k-th best
//                                                 can only be <= the gallery's)
//   phase p ("filter") tiles that are multiples   only scores >= thr[q] are appended
//                      of s_p but not of s_(p-1)  (expected k * s_(p-1)/s_p per query); an exact
//                                                 select over {previous top-k} U {appended}
//                                                 tightens thr[q]
//   last phase         s_p = 1                    select writes (values, indices)
//
// Strides are powers of two with s_(p-1)/s_p = 2^ratio_log2, so every tile is visited exactly
// once and the result is exact for any data; a strided (not prefix) sample keeps the bound
// tight when the gallery is grouped by class, as the reference's caches are
// (code/search_image.py:146-149 builds the cache class by class).  If a list still outgrows
// its capacity the select raises kFlagOverflow and the caller re-runs the query exhaustively.
#pragma once
#include <stdint.h>

namespace mmrs {

struct PlanPhase {
  int32_t inc;    // visit tiles t = j * inc
  int32_t exc;    // ... except t % exc == 0 (0 = none)
  int32_t n_sel;  // number of j
};

struct SearchPlan {
  int32_t n_tiles;
  int32_t n_phases;
  PlanPhase phase[12];
  int32_t dense_rows;  // key slots written by phase 0 per query
  int32_t cap;         // key slots per query
};

inline SearchPlan make_search_plan(int64_t n_rows, int32_t k, int32_t tile_rows,
                                   int32_t ratio_log2, int32_t dense_tiles) {
  SearchPlan pl{};
  const int64_t T = (n_rows + tile_rows - 1) / tile_rows;
  pl.n_tiles = static_cast<int32_t>(T);
  if (ratio_log2 < 1) ratio_log2 = 1;
  if (dense_tiles < 16) dense_tiles = 16;
  // the sample must hold k valid rows even when it contains the (partial) last tile
  const int64_t min_dense = (k + tile_rows - 1) / tile_rows + 1;
  if (dense_tiles < min_dense) dense_tiles = static_cast<int32_t>(min_dense);

  if (T <= 2 * static_cast<int64_t>(dense_tiles)) {
    pl.n_phases = 1;
    pl.phase[0] = PlanPhase{1, 0, static_cast<int32_t>(T)};
    pl.dense_rows = static_cast<int32_t>(T * tile_rows);
    pl.cap = pl.dense_rows;
    return pl;
  }
  int a = 0;
  while (((T + (1ll << a) - 1) >> a) > dense_tiles) ++a;
  int exps[12];
  int n = 0;
  exps[n++] = a;
  for (int e = a - ratio_log2; e > 2 && n < 10; e -= ratio_log2) exps[n++] = e;
  exps[n++] = 0;
  pl.n_phases = n;
  int max_ratio_log2 = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t inc = 1ll << exps[i];
    pl.phase[i].inc = static_cast<int32_t>(inc);
    pl.phase[i].exc = i == 0 ? 0 : (1 << exps[i - 1]);
    pl.phase[i].n_sel = static_cast<int32_t>((T + inc - 1) / inc);
    if (i > 0 && exps[i - 1] - exps[i] > max_ratio_log2) max_ratio_log2 = exps[i - 1] - exps[i];
  }
  pl.dense_rows = pl.phase[0].n_sel * tile_rows;
  // expected appends per phase: k * ratio (Beta(k, n) tail: 4x the mean of k+32 is < 1e-9)
  int64_t cap = 4ll * (k + 32) * (1ll << max_ratio_log2);
  if (cap < pl.dense_rows) cap = pl.dense_rows;
  cap = (cap + 127) / 128 * 128;
  pl.cap = static_cast<int32_t>(cap);
  return pl;
}

// ---- which tiles a phase launch visits, enumerated densely (used by the CTA-pair scan) --------
// The tiles of a phase are t = j * inc, j in [0, n_sel), except those with t % exc == 0, i.e.
// j % R == 0 with R = exc / inc (strides are powers of two, so inc divides exc).
#ifdef __CUDACC__
#define MMRS_HD __host__ __device__ __forceinline__
#else
#define MMRS_HD inline
#endif
MMRS_HD int32_t plan_exclusion_ratio(int32_t inc, int32_t exc) { return exc ? exc / inc : 0; }
MMRS_HD int32_t plan_n_visited(int32_t n_sel, int32_t R) {
  return R == 1 ? 0 : (R ? n_sel - (n_sel + R - 1) / R : n_sel);
}
// the v-th visited tile's j, v in [0, plan_n_visited)
MMRS_HD int32_t plan_visited_to_j(int32_t v, int32_t R) { return R ? (v / (R - 1)) * R + v % (R - 1) + 1 : v; }

// Work units of the CTA-pair scan: unit u = (tile pair u / n_chunks, query chunk u % n_chunks);
// CTA `rank` of the pair takes visited tile 2 * (tile pair) + rank.  An odd tile count leaves the
// last pair half empty: that CTA repeats its partner's tile and `valid` tells its epilogue to skip it.
struct PairUnit { int32_t j, chunk; bool valid; };
MMRS_HD int32_t plan_pair_units(int32_t n_visited, int32_t n_chunks) { return (n_visited + 1) / 2 * n_chunks; }
MMRS_HD PairUnit plan_pair_unit(int32_t u, int32_t rank, int32_t n_chunks, int32_t n_visited, int32_t R) {
  const int32_t tp = u / n_chunks;
  const int32_t v = 2 * tp + rank;
  const bool valid = v < n_visited;
  return PairUnit{plan_visited_to_j(valid ? v : 2 * tp, R), u - tp * n_chunks, valid};
}

}  // namespace mmrs

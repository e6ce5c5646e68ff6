// selfjoin_mma.cu -- K5 on tensor cores: bf16 tcgen05 prefilter + exact fp32 recheck.
//
// N(N-1)/2 dot products is a GEMM (5.0e13 pairs at N = 10 M), so the
// candidate pass runs on the 5th-gen tensor cores over a bf16 copy of the unit-norm embeddings:
//   S[256 rows i, 256 rows j] = X16[i-blocks] * X16[j-block]^T     (tcgen05.mma.cta_group::2, M=256, N=256)
// by a PAIR of CTAs (one TPC): each CTA stages its own 128-row i-block and half of the j-block,
// CTA 0's MMA warp issues for both, each CTA's 128 x 256 accumulator lands in its own TMEM (less
// operand traffic per SM, see scan_mma.cu); long joins use single CTAs (M=128) instead, see
// sjm_pair_mode.  The epilogue keeps pairs (i < j) with S >= tau - margin.  Rounding unit rows to bf16 moves a
// dot product by at most ||a|| ||b - b^|| + ||a - a^|| ||b^|| <= 2 * 2^-9 (1 + 2^-9) < 0.004
// (Cauchy-Schwarz; round-to-nearest is within 2^-9 relative per element), so with margin >= 0.004
// every true pair survives; the few survivors are then re-scored in fp32 with exactly the
// arithmetic of the exact kernel (selfjoin.cu: one fmaf chain, k ascending), which makes the
// emitted pair set identical to the exact mode's.
//
// Schedule: the upper triangle is cut into panels of 8 j-blocks (2048 rows: 2 MB of bf16 at
// D = 512, L2 resident); inside a panel tiles are numbered j-fastest, so CTAs that run together
// share their A rows and the B panel through L2.  Panels are dealt round-robin to the ranks of a
// multi-GPU run (work per panel grows linearly with its index, so cyclic dealing balances).
#include <stdlib.h>

#include "tcgen05_utils.cuh"

namespace mmrs {

constexpr int kSjmThreads = 256;
constexpr int kSjmBM = 128;
constexpr int kSjmBN = 256;
constexpr int kSjmBK = 64;
constexpr int kSjmMaxStages = 6;
constexpr int kSjmPanel = 8;   // j-blocks per panel
constexpr int kSjmABytes = kSjmBM * kSjmBK * 2;
// B rows a CTA stages per k-block: the whole j-block, or half of it in pair mode; ring depth to match
template <bool PAIR> struct SjmCfg {
  static constexpr int kBBytes = (PAIR ? kSjmBN / 2 : kSjmBN) * kSjmBK * 2;
  static constexpr int kStageBytes = kSjmABytes + kBBytes;
  static constexpr int kStages = PAIR ? 6 : 4;
};

struct SjmShared {
  uint64_t full[kSjmMaxStages];
  uint64_t empty[kSjmMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  volatile uint32_t abort;
};

struct SjmParams {
  int64_t n_rows;
  int32_t k_blocks;
  float thr_lo;                    // tau - margin
  int64_t nbi, nbj;                // number of 128-row i-blocks / 256-row j-blocks
  int32_t n_my_panels;             // panels dealt to this rank: panel_begin + m * panel_stride
  int32_t panel_begin, panel_stride;
  const int64_t* panel_start;      // [n_my_panels + 1] first tile number of each of my panels
  int64_t total_tiles;
  int32_t ib_per_tile;             // i-blocks per work unit: 2 in pair mode, else 1
  int64_t* cand;                   // [cand_cap, 2]
  int64_t cand_cap;
  unsigned long long* cand_count;
};

// work-unit number -> (first i-block of the unit, j-block); returns false when that i-block lies
// entirely on/below the diagonal.  In pair mode a unit is two i-blocks: if the first is below the
// diagonal so is the second; when only the second is, its CTA computes a tile whose columns all
// fail the epilogue's j > i test.
__device__ __forceinline__ bool sjm_decode(const SjmParams& p, int64_t t, int64_t& ib, int64_t& jb) {
  int lo = 0, hi = p.n_my_panels;           // last m with panel_start[m] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.panel_start[mid] <= t) lo = mid; else hi = mid;
  }
  const int64_t panel = p.panel_begin + static_cast<int64_t>(lo) * p.panel_stride;
  const int64_t u = t - p.panel_start[lo];
  int64_t nj = p.nbj - panel * kSjmPanel;
  if (nj > kSjmPanel) nj = kSjmPanel;
  ib = p.ib_per_tile * (u / nj);
  jb = panel * kSjmPanel + u % nj;
  return jb * kSjmBN + (kSjmBN - 1) > ib * kSjmBM;
}

template <bool PAIR>
__global__ void __launch_bounds__(kSjmThreads, 1)
selfjoin_mma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const SjmParams p, int32_t* flags) {
  constexpr int kSjmStages = SjmCfg<PAIR>::kStages;
  constexpr int kSjmStageBytes = SjmCfg<PAIR>::kStageBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SjmShared* sh = reinterpret_cast<SjmShared*>(ring + kSjmStages * kSjmStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0: pair leader
  // work units of this CTA (pair): pair, pair + n_pairs, ...
  const int64_t pair = PAIR ? blockIdx.x >> 1 : blockIdx.x, n_pairs = PAIR ? gridDim.x >> 1 : gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSjmStages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
    // pair mode: the leader's tmem_empty collects the four epilogue warps of both CTAs
    for (int a = 0; a < 2; ++a) { mbar_init(&sh->tmem_full[a], 1); mbar_init(&sh->tmem_empty[a], PAIR ? 8 : 4); }
    sh->abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if constexpr (PAIR) {
    __syncwarp();
    cluster_sync_all();   // the peer's barriers exist before anything is signalled across
  } else {
    __syncthreads();
  }
  tcgen05_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer (both CTAs; bytes are credited to the leader's barrier) =====
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      for (int64_t t = pair; t < p.total_tiles && ok; t += n_pairs) {
        int64_t ib, jb;
        if (!sjm_decode(p, t, ib, jb)) continue;
        const int32_t a_row = static_cast<int32_t>((ib + rank) * kSjmBM);   // past the last block: zero-filled
        const int32_t b_row = static_cast<int32_t>(jb * kSjmBN + rank * (kSjmBN / 2));
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (!mbar_wait(&sh->empty[stage], phase ^ 1, &sh->abort, flags)) { ok = false; break; }
          uint8_t* a_dst = ring + static_cast<size_t>(stage) * kSjmStageBytes;
          if constexpr (PAIR) {
            const uint32_t full_leader = mapa_u32(smem_u32(&sh->full[stage]), 0);
            if (rank == 0) mbar_expect_tx(&sh->full[stage], 2 * kSjmStageBytes);
            tma_load_2d_pair(a_dst, &map_a, full_leader, kb * kSjmBK, a_row, kEvictLast);
            tma_load_2d_pair(a_dst + kSjmABytes, &map_b, full_leader, kb * kSjmBK, b_row, kEvictLast);
          } else {
            mbar_expect_tx(&sh->full[stage], kSjmStageBytes);
            tma_load_2d(a_dst, &map_a, &sh->full[stage], kb * kSjmBK, a_row, kEvictLast);
            tma_load_2d(a_dst + kSjmABytes, &map_b, &sh->full[stage], kb * kSjmBK, b_row, kEvictLast);
          }
          if (++stage == kSjmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader's whole warp walks the loop, one elected lane issues (descriptors
    // stay in uniform registers; with the loop inside `if (lane == 0)` the issue sequence, not the
    // tensor pipe, set the pace -- see scan_mma.cu) =====
    if (rank == 0) {
      const uint32_t idesc = make_idesc(kSjmBN, PAIR ? 256u : 128u);
      const uint32_t ring_addr = smem_u32(ring);
      uint32_t stage = 0, phase = 0, it = 0;
      bool ok = true;
      for (int64_t t = pair; t < p.total_tiles && ok; t += n_pairs) {
        int64_t ib, jb;
        if (!sjm_decode(p, t, ib, jb)) continue;
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        if (!__all_sync(0xffffffffu, mbar_wait(&sh->tmem_empty[as], aphase ^ 1, &sh->abort, flags))) break;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * kSjmBN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (!__all_sync(0xffffffffu, mbar_wait(&sh->full[stage], phase, &sh->abort, flags))) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t a_addr = ring_addr + stage * kSjmStageBytes;
          if (elect_one()) {
            const uint64_t adesc = make_sw128_desc(a_addr);
            const uint64_t bdesc = make_sw128_desc(a_addr + kSjmABytes);
#pragma unroll
            for (int k = 0; k < kSjmBK / 16; ++k) {
              if constexpr (PAIR)
                umma_bf16_pair(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                               (kb | k) != 0 ? 1u : 0u);
              else
                umma_bf16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
            }
            if constexpr (PAIR) umma_commit_pair(&sh->empty[stage]); else umma_commit(&sh->empty[stage]);
            if (kb == p.k_blocks - 1) {
              if constexpr (PAIR) umma_commit_pair(&sh->tmem_full[as]); else umma_commit(&sh->tmem_full[as]);
            }
          }
          __syncwarp();
          if (++stage == kSjmStages) { stage = 0; phase ^= 1; }
        }
        ++it;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: keep (i < j) with S >= tau - margin =====
    const int ew = warp - 4;
    const uint32_t tmem_empty_leader = PAIR ? mapa_u32(smem_u32(&sh->tmem_empty[0]), 0) : 0u;
    uint32_t it = 0;
    for (int64_t t = pair; t < p.total_tiles; t += n_pairs) {
      int64_t ib, jb;
      if (!sjm_decode(p, t, ib, jb)) continue;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      if (!mbar_wait(&sh->tmem_full[as], aphase, &sh->abort, flags)) break;
      tcgen05_fence_after();
      const int64_t i = (ib + rank) * kSjmBM + ew * 32 + lane;
      const int64_t j0 = jb * kSjmBN;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * kSjmBN;
      // a warp whose 32 rows all lie at or beyond the j-block's last column has nothing to keep
      if ((ib + rank) * kSjmBM + ew * 32 < j0 + kSjmBN - 1) {
        for (int c0 = 0; c0 < kSjmBN; c0 += 16) {
          uint32_t acc[16];
          __syncwarp();
          tmem_ld16(taddr0 + static_cast<uint32_t>(c0), acc);
          tmem_ld_wait();
          uint32_t bits = 0;
#pragma unroll
          for (int c = 0; c < 16; ++c) bits |= (__uint_as_float(acc[c]) >= p.thr_lo) ? (1u << c) : 0u;
          while (bits) {
            const int c = __ffs(bits) - 1;
            bits &= bits - 1;
            const int64_t j = j0 + c0 + c;
            if (j > i && j < p.n_rows) {   // i < j < n_rows
              const unsigned long long pos = atomicAdd(p.cand_count, 1ull);
              if (pos < static_cast<unsigned long long>(p.cand_cap)) {
                p.cand[2 * pos] = i;
                p.cand[2 * pos + 1] = j;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader + as * 8u);
        else mbar_arrive(&sh->tmem_empty[as]);
      }
      ++it;
    }
  }

  tcgen05_fence_before();
  if constexpr (PAIR) {
    __syncwarp();
    cluster_sync_all();   // neither CTA may retire while its peer can still signal or read it
  } else {
    __syncthreads();
  }
  tcgen05_fence_after();
  if (warp == 2) {
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Exact fp32 re-score of the candidates: the same single fmaf chain, k ascending, as
// selfjoin_f32_kernel, so both modes take bit-identical decisions.
__global__ void __launch_bounds__(256) selfjoin_recheck_kernel(const int64_t* __restrict__ cand, int64_t n_cand,
                                                               const float* __restrict__ emb, int64_t ld, int32_t dim,
                                                               float threshold, int64_t* __restrict__ out_pairs,
                                                               int64_t capacity, unsigned long long* out_count) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c >= n_cand) return;
  const int64_t i = cand[2 * c], j = cand[2 * c + 1];
  const float4* a = reinterpret_cast<const float4*>(emb + i * ld);
  const float4* b = reinterpret_cast<const float4*>(emb + j * ld);
  float acc = 0.f;
  for (int k = 0; k < dim / 4; ++k) {
    const float4 x = a[k], y = b[k];
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  if (acc >= threshold) {
    const unsigned long long pos = atomicAdd(out_count, 1ull);
    if (pos < static_cast<unsigned long long>(capacity)) {
      out_pairs[2 * pos] = i;
      out_pairs[2 * pos + 1] = j;
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
// Fills h_panel_start (n_my_panels + 1 entries) and returns the launch parameters.
// CTA pairs or single CTAs?  Pairs keep the tensor pipe busier (+11-14 % on joins that finish within
// ~0.1 s), but a long join runs against the board's power cap, where what counts is energy per
// flop and the single-CTA kernel came out 5 % ahead (same-box A/B on one rank's share of C3:
// 5.91 s vs 6.20-6.31 s, profiles/r01_selfjoin_ab.log).  So: pairs below ~0.12 s of estimated work.
bool sjm_pair_mode(int64_t n_rows, int32_t dim, int32_t world) {
  if (const char* e = getenv("MMRS_SJ_PAIR")) return atoi(e) != 0;
  const double flops = static_cast<double>(dim) * static_cast<double>(n_rows) * static_cast<double>(n_rows) /
                       static_cast<double>(world > 0 ? world : 1);
  return flops < 1.44e14;
}

int64_t sjm_plan(int64_t n_rows, int32_t rank, int32_t world, bool pair, int64_t* h_panel_start, int32_t max_panels,
                 int32_t* n_my_panels) {
  const int64_t nbi = (n_rows + kSjmBM - 1) / kSjmBM, nbj = (n_rows + kSjmBN - 1) / kSjmBN;
  const int64_t n_panels = (nbj + kSjmPanel - 1) / kSjmPanel;
  int32_t m = 0;
  int64_t tiles = 0;
  for (int64_t pnl = rank; pnl < n_panels && m < max_panels; pnl += world, ++m) {
    if (h_panel_start) h_panel_start[m] = tiles;
    int64_t ni = 2 * kSjmPanel * (pnl + 1);
    if (ni > nbi) ni = nbi;
    int64_t nj = nbj - pnl * kSjmPanel;
    if (nj > kSjmPanel) nj = kSjmPanel;
    tiles += (pair ? (ni + 1) / 2 : ni) * nj;   // a CTA pair takes two i-blocks at a time
  }
  if (h_panel_start) h_panel_start[m] = tiles;
  *n_my_panels = m;
  return tiles;
}

int64_t sjm_max_panels(int64_t n_rows) {
  const int64_t nbj = (n_rows + kSjmBN - 1) / kSjmBN;
  return (nbj + kSjmPanel - 1) / kSjmPanel;
}

cudaError_t launch_selfjoin_mma(const __nv_bfloat16* emb16, int64_t n_rows, int32_t dim, int64_t ld16,
                                float thr_lo, int32_t rank, int32_t world, bool pair, const int64_t* d_panel_start,
                                int32_t n_my_panels, int64_t total_tiles, int64_t* cand, int64_t cand_cap,
                                unsigned long long* cand_count, int32_t* flags, int sm_count,
                                cudaStream_t stream) {
  if (total_tiles <= 0) return cudaSuccess;
  if (n_rows > 0x7fffffffll - kSjmBN) return cudaErrorInvalidValue;
  SjmParams p{};
  p.n_rows = n_rows;
  p.k_blocks = (dim + kSjmBK - 1) / kSjmBK;
  p.thr_lo = thr_lo;
  p.nbi = (n_rows + kSjmBM - 1) / kSjmBM;
  p.nbj = (n_rows + kSjmBN - 1) / kSjmBN;
  p.n_my_panels = n_my_panels;
  p.panel_begin = rank;
  p.panel_stride = world;
  p.panel_start = d_panel_start;
  p.total_tiles = total_tiles;
  p.ib_per_tile = pair ? 2 : 1;
  p.cand = cand; p.cand_cap = cand_cap; p.cand_count = cand_count;
  CUtensorMap map_a, map_b;
  if (!make_map(&map_a, emb16, static_cast<uint64_t>(n_rows), static_cast<uint64_t>(dim), static_cast<uint64_t>(ld16), kSjmBM))
    return cudaErrorNotSupported;
  if (!make_map(&map_b, emb16, static_cast<uint64_t>(n_rows), static_cast<uint64_t>(dim), static_cast<uint64_t>(ld16), pair ? kSjmBN / 2 : kSjmBN))
    return cudaErrorNotSupported;
  const size_t smem = 1024 + sizeof(SjmShared) +
                      (pair ? static_cast<size_t>(SjmCfg<true>::kStages) * SjmCfg<true>::kStageBytes
                            : static_cast<size_t>(SjmCfg<false>::kStages) * SjmCfg<false>::kStageBytes);
  auto kernel = pair ? selfjoin_mma_kernel<true> : selfjoin_mma_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int64_t units = pair ? sm_count / 2 : sm_count;          // one CTA (pair) per SM (TPC)
  if (units > total_tiles) units = total_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(pair ? 2 * units : units)); cfg.blockDim = dim3(kSjmThreads);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = pair ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, map_a, map_b, p, flags);
}

cudaError_t launch_selfjoin_recheck(const int64_t* cand, int64_t n_cand, const float* emb, int64_t ld, int32_t dim,
                                    float threshold, int64_t* out_pairs, int64_t capacity,
                                    unsigned long long* out_count, cudaStream_t stream) {
  if (n_cand <= 0) return cudaSuccess;
  selfjoin_recheck_kernel<<<static_cast<int>((n_cand + 255) / 256), 256, 0, stream>>>(
      cand, n_cand, emb, ld, dim, threshold, out_pairs, capacity, out_count);
  return cudaGetLastError();
}

}  // namespace mmrs

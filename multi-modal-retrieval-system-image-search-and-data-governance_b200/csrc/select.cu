// select.cu -- K3/K4: exact per-query top-k over a list of unique u64 keys.
//
// One 1024-thread CTA per query.  The list (survivors of the previous phase + the keys the scan
// appended, or a dense seed sample) is read once, coalesced, into registers.  The k-th largest
// key is then found by a bit-wise descent, two bits per step: the block counts how many keys are
// >= each of three pivots (register compares + one REDUX per warp + one barrier) and keeps the
// largest pivot that still has >= k keys above it; leading bits shared by all keys are skipped
// and the walk stops as soon as a pivot has exactly k keys above it.  No shared-memory atomics:
// a histogram radix select costs one ATOMS per key, which at ~50 clk per warp-wide ATOMS made the
// first version of this kernel 11-20 us per call (profiles/r01_launches_b16.csv).
// The k winners are ranked by counting (rank = #keys greater), which sorts them without any
// barrier-separated network stages, and are written either back to the front of the list with
// the new score bound (between scan phases) or as (values, indices) in torch.topk order
// (code/utils.py:17: largest, sorted).  Keys are (score, ~row), so the result is exactly "score
// descending, row ascending" whatever order the scan kernels appended the candidates in.
//
// The same kernel merges the per-GPU lists after the all-gather (mmrs_topk_merge): pack_keys
// turns (value, global index) pairs back into keys.
#include "common.cuh"

namespace mmrs {

constexpr int kSelThreads = 1024;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelMaxK = 1024;
constexpr int kSelMaxKpt = 16;                     // register-resident lists: up to 16 Ki keys
constexpr int kSelWinners = kSelMaxK + kSelMaxK / 4 + 8;

struct SelShared {
  uint32_t warp_cnt[2][kSelWarps];   // packed per-warp pivot counts, double buffered
  uint32_t wide_cnt[3][kSelWarps];   // generic path: unpacked
  uint64_t red_min[kSelWarps];
  uint64_t red_max[kSelWarps];
  uint32_t n_out;
  uint64_t winners[kSelWinners];
};

// Between phases only a BOUND is needed: any pivot with at least k and at most this many keys
// above it ends the descent; the (unsorted) keys above it become the next list.
__device__ __forceinline__ uint32_t relaxed_limit(uint32_t k) { return k + k / 4 + 8; }

__global__ void __launch_bounds__(kSelThreads, 1) select_topk_kernel(const SelectParams p) {
  __shared__ SelShared sh;
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* list = p.cand + static_cast<int64_t>(q) * p.cap;

  uint32_t n;
  if (p.fixed_n >= 0) {
    n = static_cast<uint32_t>(p.fixed_n);
  } else {
    n = p.cnt[q];
    if (n > static_cast<uint32_t>(p.cap)) {
      if (tid == 0) atomicOr(p.flags, kFlagOverflow);
      n = p.cap;
    }
  }
  const uint32_t k = p.k;
  if (n < k) {  // cannot happen on a correct schedule; never read past the list
    if (tid == 0) atomicOr(p.flags, kFlagShort);
    return;
  }
  const uint32_t limit = p.final_pass ? k : relaxed_limit(k);

  // Lists up to 16 Ki keys live in registers, `kpt` (block-uniform) slots per thread; longer ones
  // (the exhaustive fallback) are re-read from global memory at every step.
  const int kpt = static_cast<int>((n + kSelThreads - 1) / kSelThreads);
  const bool in_regs = kpt <= kSelMaxKpt;
  uint64_t key[kSelMaxKpt];
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kSelMaxKpt; ++i) {
      const uint32_t idx = tid + i * kSelThreads;
      key[i] = (i < kpt && idx < n) ? list[idx] : 0ull;   // 0 is below every real key
    }
  }
  auto for_each_key = [&](auto&& f) {
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < kSelMaxKpt; ++i)
        if (i < kpt) f(key[i]);
    } else {
      // block-uniform trip count (the callbacks use warp collectives); 0 pads the tail
      for (uint32_t base = 0; base < n; base += kSelThreads) {
        const uint32_t idx = base + tid;
        f(idx < n ? list[idx] : 0ull);
      }
    }
  };

  // ---- leading bits shared by every key (zeros = padding / rows past the gallery end are ignored) --
  uint64_t mn = ~0ull, mx = 0ull;
  for_each_key([&](uint64_t x) {
    if (x != 0ull) { mn = x < mn ? x : mn; mx = x > mx ? x : mx; }
  });
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o);
    const uint64_t b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }
  if (lane == 0) { sh.red_min[warp] = mn; sh.red_max[warp] = mx; }
  if (tid == 0) sh.n_out = 0;
  __syncthreads();
  mn = sh.red_min[lane];
  mx = sh.red_max[lane];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o);
    const uint64_t b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }

  // ---- bit-wise descent towards the k-th largest key, 2 bits per step ---------------------------------
  uint64_t thr_key;
  if (n <= limit) {
    thr_key = mn;                        // every real key is selected
  } else {
    const uint64_t diff = mn ^ mx;
    int pos = diff == 0 ? 0 : 64 - __clzll(static_cast<long long>(diff));   // undecided low bits
    uint64_t prefix = pos >= 64 ? 0ull : (mx >> pos) << pos;
    int buf = 0;
    while (pos > 0) {
      const int w = pos >= 2 ? 2 : 1;
      const int shift = pos - w;
      const uint64_t p1 = prefix | (1ull << shift);
      const uint64_t p2 = prefix | (2ull << shift);     // only meaningful when w == 2
      const uint64_t p3 = prefix | (3ull << shift);
      uint32_t c1 = 0, c2 = 0, c3 = 0;
      for_each_key([&](uint64_t x) {
        c1 += x >= p1;
        c2 += x >= p2;
        c3 += x >= p3;
      });
      uint32_t t1, t2, t3;
      if (in_regs) {
        // per-warp sums are <= 32 * 16 = 512: three 10-bit fields in one REDUX
        const uint32_t packed = __reduce_add_sync(0xffffffffu, c1 | (c2 << 10) | (c3 << 20));
        if (lane == 0) sh.warp_cnt[buf][warp] = packed;
        __syncthreads();
        const uint32_t v = sh.warp_cnt[buf][lane];
        t1 = __reduce_add_sync(0xffffffffu, v & 1023u);
        t2 = __reduce_add_sync(0xffffffffu, (v >> 10) & 1023u);
        t3 = __reduce_add_sync(0xffffffffu, v >> 20);
        buf ^= 1;
      } else {
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        __syncthreads();                       // previous step's readers are done
        if (lane == 0) { sh.wide_cnt[0][warp] = c1; sh.wide_cnt[1][warp] = c2; sh.wide_cnt[2][warp] = c3; }
        __syncthreads();
        t1 = __reduce_add_sync(0xffffffffu, sh.wide_cnt[0][lane]);
        t2 = __reduce_add_sync(0xffffffffu, sh.wide_cnt[1][lane]);
        t3 = __reduce_add_sync(0xffffffffu, sh.wide_cnt[2][lane]);
      }
      // counts are non-increasing in the pivot; take the largest pivot with >= k keys above it
      uint32_t cnt_sel;
      if (w == 2 && t3 >= k) { prefix = p3; cnt_sel = t3; }
      else if (w == 2 && t2 >= k) { prefix = p2; cnt_sel = t2; }
      else if (t1 >= k) { prefix = p1; cnt_sel = t1; }
      else { cnt_sel = 0xffffffffu; }            // digit 0: prefix unchanged, count unknown (> limit)
      pos = shift;
      if (cnt_sel <= limit) break;               // final: exactly k keys are >= prefix
    }
    thr_key = prefix;
  }

  // ---- gather the winners -------------------------------------------------------------------------
  __syncthreads();
  for_each_key([&](uint64_t x) {
    const bool win = x >= thr_key && x != 0ull;
    const uint32_t m = __ballot_sync(0xffffffffu, win);
    if (m) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&sh.n_out, static_cast<uint32_t>(__popc(m)));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (win) {
        const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < kSelWinners) sh.winners[slot] = x;
      }
    }
  });
  __syncthreads();
  const uint32_t n_win = sh.n_out;
  if (n_win < k || n_win > limit) {  // uniqueness violated (bug guard)
    if (tid == 0) atomicOr(p.flags, kFlagShort);
    if (n_win > limit) return;
  }

  if (p.final_pass) {
    // rank by counting: position = number of winners with a larger key
    for (uint32_t i = tid; i < n_win; i += kSelThreads) {
      const uint64_t mine = sh.winners[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n_win; ++j) rank += sh.winners[j] > mine;
      p.out_values[static_cast<int64_t>(q) * k + rank] = key_score(mine);
      p.out_indices[static_cast<int64_t>(q) * k + rank] = static_cast<int64_t>(key_row(mine)) + p.index_offset;
    }
  } else {
    // next list = the winners (unsorted); bound = score part of the pivot (<= every winner's score)
    for (uint32_t i = tid; i < n_win; i += kSelThreads) list[i] = sh.winners[i];
    if (tid == 0) {
      p.cnt[q] = n_win;
      p.thr[q] = key_score(thr_key);
    }
  }
}

cudaError_t launch_select(const SelectParams& p, int32_t n_queries, cudaStream_t stream) {
  if (n_queries <= 0) return cudaSuccess;
  if (p.k < 1 || p.k > kSelMaxK) return cudaErrorInvalidValue;
  select_topk_kernel<<<n_queries, kSelThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// (value, global index) -> key lists laid out [n_queries, cap] with cap = n_lists * k_in.
__global__ void pack_keys_kernel(const float* __restrict__ values,
                                 const int64_t* __restrict__ indices, int32_t n_lists,
                                 int32_t n_queries, int32_t k_in, uint64_t* __restrict__ cand,
                                 int32_t cap) {
  const int64_t total = static_cast<int64_t>(n_lists) * n_queries * k_in;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int j = static_cast<int>(i % k_in);
  const int q = static_cast<int>((i / k_in) % n_queries);
  const int l = static_cast<int>(i / (static_cast<int64_t>(k_in) * n_queries));
  cand[static_cast<int64_t>(q) * cap + l * k_in + j] =
      make_key(values[i], static_cast<uint32_t>(indices[i]));
}

cudaError_t launch_pack_keys(const float* values, const int64_t* indices, int32_t n_lists,
                             int32_t n_queries, int32_t k_in, uint64_t* cand, int32_t cap,
                             cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(n_lists) * n_queries * k_in;
  if (total <= 0) return cudaSuccess;
  pack_keys_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
      values, indices, n_lists, n_queries, k_in, cand, cap);
  return cudaGetLastError();
}

}  // namespace mmrs

// select.cu -- K3/K4: exact per-query top-k over a list of unique u64 keys.
//
// One CTA per query.  The k-th largest key is found by an MSB-first radix select (8-bit
// digits, shared-memory histogram, leading bits common to all keys skipped), the k winners are
// gathered, sorted by a bitonic network in shared memory and written either back to the front
// of the list together with the new score bound (between scan phases) or as
// (values, indices) in torch.topk order (code/utils.py:17: largest, sorted).  Because keys are
// (score, ~row) the result is exactly "score descending, row ascending" whatever order the
// scan kernels appended the candidates in.
//
// The same kernel merges the per-GPU lists after the all-gather (mmrs_topk_merge): pack_keys
// turns (value, global index) pairs back into keys.
#include "common.cuh"

namespace mmrs {

constexpr int kSelThreads = 256;
constexpr int kSelSmemKeys = 8192;  // lists up to this long are staged in shared memory
constexpr int kSelMaxK = 1024;

struct SelShared {
  uint32_t hist[256];
  uint64_t red_min[kSelThreads / 32];
  uint64_t red_max[kSelThreads / 32];
  uint64_t prefix;
  uint32_t need;
  uint32_t done;
  uint32_t bucket;
  uint32_t n_out;
};

__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(const SelectParams p) {
  extern __shared__ __align__(16) unsigned char sel_smem_raw[];
  // layout: [sort buffer: P u64][staged keys: up to kSelSmemKeys u64]
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

  int P = 1;
  while (P < static_cast<int>(k)) P <<= 1;
  uint64_t* sort_buf = reinterpret_cast<uint64_t*>(sel_smem_raw);
  uint64_t* staged = sort_buf + P;

  const uint64_t* src = list;
  if (n <= kSelSmemKeys) {
    for (uint32_t i = tid; i < n; i += kSelThreads) staged[i] = list[i];
    src = staged;
  }
  // min / max over the list -> number of leading bits every key shares
  uint64_t mn = ~0ull, mx = 0ull;
  __syncthreads();
  for (uint32_t i = tid; i < n; i += kSelThreads) {
    const uint64_t x = src[i];
    mn = x < mn ? x : mn;
    mx = x > mx ? x : mx;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o);
    const uint64_t b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }
  if (lane == 0) { sh.red_min[warp] = mn; sh.red_max[warp] = mx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kSelThreads / 32; ++w) {
      mn = sh.red_min[w] < mn ? sh.red_min[w] : mn;
      mx = sh.red_max[w] > mx ? sh.red_max[w] : mx;
    }
    const uint64_t diff = mn ^ mx;
    const int common = diff == 0 ? 64 : __clzll(static_cast<long long>(diff));
    // bits [64-common, 64) are fixed; the radix walk starts below them
    sh.prefix = common == 0 ? 0ull : (common == 64 ? mx : (mx >> (64 - common)) << (64 - common));
    sh.bucket = 64 - common;  // "pos": number of undecided low bits
    sh.need = k;
    sh.done = (n == k || common == 64) ? 1u : 0u;
    if (n == k) { sh.prefix = mn; }   // everything is selected: threshold = smallest key
  }
  __syncthreads();

  // ---- radix walk ---------------------------------------------------------------------------
  while (!sh.done) {
    const int pos = sh.bucket;
    const int w = pos < 8 ? pos : 8;
    const int shift = pos - w;
    const uint64_t prefix = sh.prefix;
    sh.hist[tid] = 0;  // kSelThreads == 256 bins
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kSelThreads) {
      const uint64_t x = src[i];
      const bool match = pos == 64 ? true : ((x >> pos) == (prefix >> pos));
      if (match) atomicAdd(&sh.hist[(x >> shift) & ((1u << w) - 1u)], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns digits 255-8l .. 248-8l (descending); find where the running count from
      // the top reaches `need`.
      uint32_t c[8], local = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = sh.hist[255 - (lane * 8 + i)]; local += c[i]; }
      uint32_t incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const uint32_t need = sh.need;
      const uint32_t before = incl - local;
      const bool here = before < need && incl >= need;
      if (here) {  // exactly one lane
        uint32_t run = before;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (run < need && run + c[i] >= need) {
            const uint32_t digit = 255 - (lane * 8 + i);
            sh.prefix = prefix | (static_cast<uint64_t>(digit) << shift);
            sh.need = need - run;
            sh.bucket = shift;
            // the whole bucket is taken (or no bits are left): every key >= prefix wins
            sh.done = (c[i] == need - run || shift == 0) ? 1u : 0u;
          }
          run += c[i];
        }
      }
    }
    __syncthreads();
  }

  // ---- gather the winners (key >= threshold), pad, sort descending -------------------------------
  const uint64_t thr_key = sh.prefix;
  if (tid == 0) sh.n_out = 0;
  for (int i = tid; i < P; i += kSelThreads) sort_buf[i] = 0ull;
  __syncthreads();
  for (uint32_t i = tid; i < n; i += kSelThreads) {
    const uint64_t x = src[i];
    if (x >= thr_key) {
      const uint32_t slot = atomicAdd(&sh.n_out, 1u);
      if (slot < static_cast<uint32_t>(P)) sort_buf[slot] = x;
    }
  }
  __syncthreads();
  if (sh.n_out != k && tid == 0) atomicOr(p.flags, kFlagShort);  // uniqueness violated (bug guard)

  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += kSelThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = sort_buf[lo], b = sort_buf[hi];
        if ((a < b) == desc) { sort_buf[lo] = b; sort_buf[hi] = a; }
      }
      __syncthreads();
    }
  }

  if (p.final_pass) {
    for (uint32_t i = tid; i < k; i += kSelThreads) {
      const uint64_t x = sort_buf[i];
      p.out_values[static_cast<int64_t>(q) * k + i] = key_score(x);
      p.out_indices[static_cast<int64_t>(q) * k + i] =
          static_cast<int64_t>(key_row(x)) + p.index_offset;
    }
  } else {
    for (uint32_t i = tid; i < k; i += kSelThreads) list[i] = sort_buf[i];
    if (tid == 0) {
      p.cnt[q] = k;
      p.thr[q] = key_score(sort_buf[k - 1]);
    }
  }
}

cudaError_t launch_select(const SelectParams& p, int32_t n_queries, cudaStream_t stream) {
  if (n_queries <= 0) return cudaSuccess;
  if (p.k < 1 || p.k > kSelMaxK) return cudaErrorInvalidValue;
  int P = 1;
  while (P < p.k) P <<= 1;
  const size_t smem = static_cast<size_t>(P + kSelSmemKeys) * sizeof(uint64_t);
  static bool attr_set = false;  // idempotent attribute; a benign race at worst sets it twice
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(select_topk_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>((kSelMaxK + kSelSmemKeys) * sizeof(uint64_t)));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  select_topk_kernel<<<n_queries, kSelThreads, smem, stream>>>(p);
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

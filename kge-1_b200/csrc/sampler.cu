// On-device negative sampling: the step before the negative-sampling hot path, which the reference
// runs on the CPU inside DataLoader workers (kge/util/sampler.py).
//   * uniform negatives                -- KgeUniformSampler._sample            (sampler.py:195-198, torch.randint)
//   * filter known positives + redraw  -- _filter_and_resample[_fast/_numba]   (sampler.py:148-176, 257-315)
//   * shared negatives (WOR / WR)      -- KgeUniformSampler._sample_shared     (sampler.py:200-255)
// Random numbers are Philox4x32-10 (counter-based: element e of a call uses subsequence e, retries advance the low
// counter word), so results depend only on (seed, offset) -- never on the launch shape or thread timing -- and the
// numpy restatement in oracle/sampler_oracle.py reproduces them bit for bit.  The reference's own stream
// (torch / numpy / random generators) is not reproducible on a device; the distributional contract is what the tests
// pin: range, no known positive after filtering, distinctness of shared WOR samples, the row's own positive dropped.
// (seed, offset) live in device memory so that a CUDA-graph replay draws fresh numbers: kgeb_philox_advance bumps the
// offset on the stream after each sampling call.
#include <cstdint>

#include "common.cuh"

namespace kgeb {

struct Philox {
  uint32_t c[4];
};
__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox p;
  p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
  return p;
}

// draw number `t` of element `elem` of stream `stream_id` of the call at (seed, offset): a value in [0, range)
// (multiply-shift of one 32-bit word; bias <= range / 2^32, the same order as torch.randint's modulo)
__device__ __forceinline__ int64_t draw(const uint64_t seed, const uint64_t offset, uint32_t stream_id, uint64_t elem,
                                        uint32_t t, uint32_t range) {
  const uint64_t ctr = offset + t;
  const Philox p = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)elem,
                                 (uint32_t)(elem >> 32) ^ (stream_id << 28), (uint32_t)seed, (uint32_t)(seed >> 32));
  return (int64_t)(((uint64_t)p.c[0] * range) >> 32);
}

__global__ void philox_words_kernel(uint64_t seed, uint64_t offset, uint64_t elem0, int64_t n, uint32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t elem = elem0 + i;
  const Philox p = philox4x32_10((uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)elem, (uint32_t)(elem >> 32),
                                 (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int k = 0; k < 4; ++k) out[4 * i + k] = p.c[k];
}

__global__ void philox_advance_kernel(uint64_t* state, uint64_t inc) { state[1] += inc; }

__global__ void sample_uniform_kernel(const uint64_t* __restrict__ state, uint32_t vocab, int64_t n,
                                      int64_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = draw(state[0], state[1], 0, (uint64_t)i, 0, vocab);
}

constexpr uint32_t kMaxTries = 1u << 16;   // redraws per element before the call reports an unsatisfiable row

__device__ __forceinline__ bool in_sorted(const int64_t* __restrict__ a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == v;
}

// one warp per positive triple: lane 0 finds the key row of (key_a, key_b) in the filter index, the lanes stride over
// the row's N samples; a sample found among the known positives is redrawn until it is a true negative
__global__ void __launch_bounds__(256)
sample_filter_kernel(const uint64_t* __restrict__ state, uint32_t vocab, kgeb_index_t ix,
                     const int64_t* __restrict__ key_a, const int64_t* __restrict__ key_b, int64_t B, int64_t N,
                     int64_t* __restrict__ neg, int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  int64_t r = -1;
  if (lane == 0) {
    const int64_t a = key_a[row], b = key_b[row];
    int64_t lo = 0, hi = ix.num_keys;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int64_t ka = ix.keys[2 * mid], kb = ix.keys[2 * mid + 1];
      if (ka < a || (ka == a && kb < b)) lo = mid + 1; else hi = mid;
    }
    if (lo < ix.num_keys && ix.keys[2 * lo] == a && ix.keys[2 * lo + 1] == b) r = lo;
  }
  r = __shfl_sync(0xffffffffu, r, 0);
  if (r < 0) return;                                   // no known positives for this pair
  const int64_t p0 = ix.offsets[r], np = ix.offsets[r + 1] - p0;
  const int64_t* pos = ix.values + p0;
  const uint64_t seed = state[0], offset = state[1];
  for (int64_t j = lane; j < N; j += 32) {
    int64_t v = neg[row * N + j];
    if (!in_sorted(pos, np, v)) continue;
    uint32_t t = 1;
    for (; t < kMaxTries; ++t) {
      v = draw(seed, offset, 1, (uint64_t)(row * N + j), t, vocab);
      if (!in_sorted(pos, np, v)) break;
    }
    if (t == kMaxTries) atomicExch(status, 1); else neg[row * N + j] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// shared negatives
// ---------------------------------------------------------------------------------------------
constexpr int32_t kEmpty = -1;
__device__ __forceinline__ uint32_t hash_slot(int32_t v, uint32_t mask) { return ((uint32_t)v * 2654435761u) & mask; }

// inserts v with owner tag `tag` (smaller tag wins); returns true when v was not in the table before this call
// raced... (the return value is only used by the distinct count, where every winner of an empty slot counts once)
__device__ __forceinline__ bool table_insert(int32_t* keys, int32_t* owner, uint32_t mask, int32_t v, int32_t tag) {
  uint32_t h = hash_slot(v, mask);
  while (true) {
    const int32_t k = atomicCAS(&keys[h], kEmpty, v);
    if (k == kEmpty || k == v) {
      atomicMin(&owner[h], tag);
      return k == kEmpty;
    }
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ int32_t table_owner(const int32_t* keys, const int32_t* owner, uint32_t mask, int32_t v) {
  uint32_t h = hash_slot(v, mask);
  while (true) {
    const int32_t k = keys[h];
    if (k == v) return owner[h];
    if (k == kEmpty) return -1;
    h = (h + 1) & mask;
  }
}

// One block.  meta[0] = num_distinct, meta[1] = 1 when the draw did not converge.
//   WR:  num_distinct = number of distinct values among N draws from vocab - 1 (sampler.py:209-219)
//   WOR: num_distinct = N
//   shared[0 .. num_distinct] = num_distinct + 1 distinct values (sampler.py:230-232): draws whose value is already
//   taken are redrawn; among equal values of one round the smallest position keeps it (deterministic)
//   up_idx[k], k in [num_distinct, N) = column to copy for the WR upsample (sampler.py:249-253)
__global__ void __launch_bounds__(1024)
shared_draw_kernel(const uint64_t* __restrict__ state, uint32_t vocab, int N, int with_replacement, int32_t* keys,
                   int32_t* owner, uint32_t cap, int64_t* __restrict__ shared, int32_t* __restrict__ up_idx,
                   int32_t* __restrict__ meta) {
  __shared__ int s_count, s_pending;
  const uint32_t mask = cap - 1;
  const uint64_t seed = state[0], offset = state[1];
  const int tid = threadIdx.x, nt = blockDim.x;
  auto clear = [&]() {
    for (uint32_t h = tid; h < cap; h += nt) { keys[h] = kEmpty; owner[h] = 0x7fffffff; }
    __syncthreads();
  };
  int nd = N;
  if (with_replacement) {
    if (tid == 0) s_count = 0;
    clear();
    int mine = 0;
    for (int j = tid; j < N; j += nt)
      mine += table_insert(keys, owner, mask, (int32_t)draw(seed, offset, 2, (uint64_t)j, 0, vocab - 1), 0) ? 1 : 0;
    atomicAdd(&s_count, mine);
    __syncthreads();
    nd = s_count;
    __syncthreads();
  }
  clear();
  const int m = nd + 1;
  // this thread's positions j = tid, tid + nt, ...; tries[] lives in registers for up to 32 positions per thread
  uint32_t tries[32];
  int64_t val[32];
  bool done[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) { tries[q] = 0; done[q] = false; val[q] = 0; }
  int round = 0;
  for (; round < 4096; ++round) {
    if (tid == 0) s_pending = 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int j = tid + q * nt;
      if (j < m && !done[q]) {
        val[q] = draw(seed, offset, 3, (uint64_t)j, tries[q], vocab);
        table_insert(keys, owner, mask, (int32_t)val[q], (round << 16) | j);
      }
    }
    __syncthreads();
    int pend = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int j = tid + q * nt;
      if (j < m && !done[q]) {
        if (table_owner(keys, owner, mask, (int32_t)val[q]) == ((round << 16) | j)) {
          done[q] = true;
          shared[j] = val[q];
        } else {
          ++tries[q];
          ++pend;
        }
      }
    }
    if (pend) atomicAdd(&s_pending, pend);
    __syncthreads();
    const int left = s_pending;
    __syncthreads();
    if (left == 0) break;
  }
  for (int k = nd + tid; k < N; k += nt) up_idx[k] = (int32_t)draw(seed, offset, 4, (uint64_t)k, 0, (uint32_t)nd);
  if (tid == 0) {
    meta[0] = nd;
    meta[1] = round >= 4096 ? 1 : 0;
  }
}

// one warp per positive triple (sampler.py:234-253): the row's own positive, if it is among the shared samples, is the
// one dropped (overwritten by the spare sample shared[num_distinct]); otherwise a random position (or none) is
__global__ void __launch_bounds__(256)
shared_rows_kernel(const uint64_t* __restrict__ state, const int64_t* __restrict__ positives, int64_t B, int N,
                   const int32_t* __restrict__ keys, const int32_t* __restrict__ owner, uint32_t cap,
                   const int64_t* __restrict__ shared, const int32_t* __restrict__ up_idx,
                   const int32_t* __restrict__ meta, int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int nd = meta[0];
  int drop = 0;
  if (lane == 0) {
    const int64_t pv = positives[row];
    int32_t o = pv < 0x7fffffff ? table_owner(keys, owner, cap - 1, (int32_t)pv) : -1;
    drop = o >= 0 ? (o & 0xffff) : (int)draw(state[0], state[1], 5, (uint64_t)row, 0, (uint32_t)(nd + 1));
  }
  drop = __shfl_sync(0xffffffffu, drop, 0);
  const int64_t spare = shared[nd];
  for (int c = lane; c < N; c += 32) {
    const int src = c < nd ? c : up_idx[c];
    out[row * N + c] = (src == drop) ? spare : shared[src];   // drop == nd: the spare itself is the one left out
  }
}

static inline uint32_t table_capacity(int64_t N) {
  uint32_t cap = 64;
  while ((int64_t)cap < 2 * (N + 2)) cap <<= 1;
  return cap;
}

}  // namespace kgeb

using namespace kgeb;

extern "C" {

int kgeb_philox_words(uint64_t seed, uint64_t offset, uint64_t elem0, int64_t n, uint32_t* out, void* stream) {
  KGEB_REQUIRE(out && n >= 0, "philox_words: bad arguments");
  if (n == 0) return KGEB_OK;
  philox_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(seed, offset, elem0, n, out);
  KGEB_LAUNCH_CHECK("philox_words");
  return KGEB_OK;
}

int kgeb_philox_advance(uint64_t* state, uint64_t inc, void* stream) {
  KGEB_REQUIRE(state, "philox_advance: bad arguments");
  philox_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(state, inc);
  KGEB_LAUNCH_CHECK("philox_advance");
  return KGEB_OK;
}

int kgeb_sample_uniform(const uint64_t* state, int64_t vocab, int64_t n, int64_t* out, void* stream) {
  KGEB_REQUIRE(state && out && n >= 0, "sample_uniform: bad arguments");
  KGEB_REQUIRE(vocab >= 1 && vocab < ((int64_t)1 << 31), "sample_uniform: vocabulary size %lld out of range", (long long)vocab);
  if (n == 0) return KGEB_OK;
  sample_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(state, (uint32_t)vocab, n, out);
  KGEB_LAUNCH_CHECK("sample_uniform");
  return KGEB_OK;
}

int kgeb_sample_filter(const uint64_t* state, int64_t vocab, const kgeb_index_t* index, const int64_t* key_a,
                       const int64_t* key_b, int64_t B, int64_t N, int64_t* negatives, int32_t* status, void* stream) {
  KGEB_REQUIRE(state && index && key_a && key_b && negatives && status && B >= 0 && N >= 0, "sample_filter: bad arguments");
  KGEB_REQUIRE(vocab >= 1 && vocab < ((int64_t)1 << 31), "sample_filter: vocabulary size %lld out of range", (long long)vocab);
  if (B == 0 || N == 0) return KGEB_OK;
  sample_filter_kernel<<<(unsigned)((B + 7) / 8), 256, 0, as_stream(stream)>>>(state, (uint32_t)vocab, *index, key_a, key_b, B,
                                                                             N, negatives, status);
  KGEB_LAUNCH_CHECK("sample_filter");
  return KGEB_OK;
}

int64_t kgeb_sample_shared_workspace_bytes(int64_t N) {
  if (N < 0) return -1;
  const uint32_t cap = table_capacity(N);
  return (int64_t)cap * 8 + (N + 2) * 8 + (N + 2) * 4 + 64;
}

int kgeb_sample_shared(const uint64_t* state, int64_t vocab, const int64_t* positives, int64_t B, int64_t N,
                       int with_replacement, int64_t* out, int32_t* meta, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  KGEB_REQUIRE(state && positives && out && meta && workspace && B >= 0, "sample_shared: bad arguments");
  KGEB_REQUIRE(N >= 1 && N <= 32766, "sample_shared: N = %lld not in [1, 32766]", (long long)N);
  KGEB_REQUIRE(vocab >= 2 && vocab < ((int64_t)1 << 31), "sample_shared: vocabulary size %lld out of range", (long long)vocab);
  KGEB_REQUIRE(N + 1 <= vocab, "sample_shared: %lld distinct samples from a vocabulary of %lld", (long long)(N + 1),
               (long long)vocab);
  KGEB_REQUIRE(workspace_bytes >= kgeb_sample_shared_workspace_bytes(N), "sample_shared: workspace too small");
  const uint32_t cap = table_capacity(N);
  int32_t* keys = reinterpret_cast<int32_t*>(workspace);
  int32_t* owner = keys + cap;
  int64_t* shared = reinterpret_cast<int64_t*>(owner + cap);
  int32_t* up_idx = reinterpret_cast<int32_t*>(shared + (N + 2));
  shared_draw_kernel<<<1, 1024, 0, as_stream(stream)>>>(state, (uint32_t)vocab, (int)N, with_replacement, keys, owner, cap,
                                                        shared, up_idx, meta);
  KGEB_LAUNCH_CHECK("shared_draw");
  if (B > 0) {
    shared_rows_kernel<<<(unsigned)((B + 7) / 8), 256, 0, as_stream(stream)>>>(state, positives, B, (int)N, keys, owner, cap,
                                                                           shared, up_idx, meta, out);
    KGEB_LAUNCH_CHECK("shared_rows");
  }
  return KGEB_OK;
}

}  // extern "C"

// Rank histograms and ranking metrics on the device: the step after the filtered-ranking path.
//   reference: hist_all / hist_per_relation_type / hist_per_frequency_percentile (kge/job/eval.py:138-224) -- Python
//   loops `hist[r] += 1` over every rank of a batch, with Python list membership tests for the drill-down masks --
//   and EntityRankingJob._compute_metrics (kge/job/entity_ranking.py:553-577).
// Histograms are float32 [num_entities] as in the reference; bins hold integer counts, so atomicAdd(+1.0f) is exact
// (and therefore order-independent) up to 2^24 per bin, the same point at which the reference's own float bins stop
// counting.  Metrics are reduced in a fixed order in double precision.
#include <cstdint>

#include "common.cuh"

namespace kgeb {

__global__ void rank_hist_kernel(const int64_t* __restrict__ ranks, const uint8_t* __restrict__ mask, int64_t n,
                                 int64_t num_entities, float* __restrict__ hist, int32_t* __restrict__ status) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  const int64_t r = ranks[i];
  if (r < 0 || r >= num_entities) {
    if (status) atomicExch(status, 1);
    return;
  }
  atomicAdd(hist + r, 1.0f);
}

// mask[i] = 1 iff values[i] occurs in the ascending list sorted_set[m] (the `id in set` tests of eval.py:187,205-221)
__global__ void isin_sorted_kernel(const void* __restrict__ values, int idx64, int64_t n,
                                   const int64_t* __restrict__ sorted_set, int64_t m, uint8_t* __restrict__ mask) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = load_index(values, idx64, i);
  int64_t lo = 0, hi = m;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sorted_set[mid] < v) lo = mid + 1; else hi = mid;
  }
  mask[i] = (lo < m && sorted_set[lo] == v) ? 1 : 0;
}

constexpr int kMaxHits = 16;
struct HitsK {
  int32_t k[kMaxHits];
  int n;
};
constexpr int kMetricBlocks = 148;

// partial[b] = (count, sum hist*(r+1), sum hist/(r+1), hits@k_0 ...) over the bins of block b's contiguous slice
__global__ void __launch_bounds__(256)
metrics_partial_kernel(const float* __restrict__ hist, int64_t num_entities, HitsK hk, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int64_t per_block = (num_entities + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = min(num_entities, blockIdx.x * per_block), b1 = min(num_entities, b0 + per_block);
  double acc[3 + kMaxHits];
#pragma unroll
  for (int j = 0; j < 3 + kMaxHits; ++j) acc[j] = 0.0;
  for (int64_t r = b0 + threadIdx.x; r < b1; r += blockDim.x) {
    const double h = (double)hist[r];
    if (h != 0.0) {
      const double rank = (double)(r + 1);
      acc[0] += h;
      acc[1] += h * rank;
      acc[2] += h / rank;
#pragma unroll
      for (int j = 0; j < kMaxHits; ++j)
        if (j < hk.n && r < hk.k[j]) acc[3 + j] += h;
    }
  }
  for (int j = 0; j < 3 + hk.n; ++j) {
    sh[threadIdx.x] = acc[j];
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < (int)blockDim.x; ++k) t += sh[k];
      partial[blockIdx.x * (3 + kMaxHits) + j] = t;
    }
    __syncthreads();
  }
}

// out = (count, mean_rank, mean_reciprocal_rank, hits@k_0 ...); all zero for an empty histogram
__global__ void metrics_finish_kernel(const double* __restrict__ partial, int blocks, int nk, double* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= 3 + nk) return;
  double t = 0.0;
  for (int b = 0; b < blocks; ++b) t += partial[b * (3 + kMaxHits) + j];
  double n = 0.0;
  for (int b = 0; b < blocks; ++b) n += partial[b * (3 + kMaxHits)];
  out[j] = (j == 0) ? n : (n > 0.0 ? t / n : 0.0);
}

}  // namespace kgeb

using namespace kgeb;

extern "C" {

int kgeb_rank_hist(const int64_t* ranks, const uint8_t* mask, int64_t n, int64_t num_entities, float* hist,
                   int32_t* status, void* stream) {
  KGEB_REQUIRE(ranks && hist && n >= 0 && num_entities > 0, "rank_hist: bad arguments");
  if (n == 0) return KGEB_OK;
  rank_hist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(ranks, mask, n, num_entities, hist, status);
  KGEB_LAUNCH_CHECK("rank_hist");
  return KGEB_OK;
}

int kgeb_isin_sorted(const void* values, int idx64, int64_t n, const int64_t* sorted_set, int64_t m, uint8_t* mask,
                     void* stream) {
  KGEB_REQUIRE(values && mask && n >= 0 && m >= 0 && (sorted_set || m == 0), "isin_sorted: bad arguments");
  if (n == 0) return KGEB_OK;
  isin_sorted_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(values, idx64, n, sorted_set, m, mask);
  KGEB_LAUNCH_CHECK("isin_sorted");
  return KGEB_OK;
}

int64_t kgeb_rank_metrics_workspace_bytes(void) { return (int64_t)kMetricBlocks * (3 + kMaxHits) * 8; }

int kgeb_rank_metrics(const float* hist, int64_t num_entities, const int32_t* hits_at_k, int num_k, double* out,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  KGEB_REQUIRE(hist && out && workspace && num_entities > 0, "rank_metrics: bad arguments");
  KGEB_REQUIRE(num_k >= 0 && num_k <= kMaxHits && (hits_at_k || num_k == 0), "rank_metrics: at most %d hits@k values", kMaxHits);
  KGEB_REQUIRE(workspace_bytes >= kgeb_rank_metrics_workspace_bytes(), "rank_metrics: workspace too small");
  HitsK hk;
  hk.n = num_k;
  for (int j = 0; j < kMaxHits; ++j) hk.k[j] = j < num_k ? hits_at_k[j] : 0;
  double* partial = reinterpret_cast<double*>(workspace);
  metrics_partial_kernel<<<kMetricBlocks, 256, 0, as_stream(stream)>>>(hist, num_entities, hk, partial);
  KGEB_LAUNCH_CHECK("rank_metrics partial");
  metrics_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(partial, kMetricBlocks, num_k, out);
  KGEB_LAUNCH_CHECK("rank_metrics finish");
  return KGEB_OK;
}

}  // extern "C"

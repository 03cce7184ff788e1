// Data-parallel gradient exchange fused with the Adagrad update over NVLink peer memory ("replicas" row).
//
// For graphs too small to shard (FB15k-237: a 7.4 MB table) every GPU of the box trains on its own batch against a
// full replica and the replicas exchange the dense gradient of both tables every step.  NCCL's all-reduce followed by
// a separate Adagrad kernel costs a collective launch, an all-reduce (reduce-scatter + all-gather of GRADIENTS) and a
// full-table update on every rank.  Here the exchange IS the update, one kernel per table over peer-mapped buffers
// (torch symmetric memory: cuMem VMM allocations mapped into every process of the node; NVSwitch gives every pair of
// GPUs full bandwidth):
//
//   barrier A            every rank's gradient buffer is complete
//   reduce+update        rank r owns the slice [lo_r, hi_r) of the table: it reads that slice of the gradient from all
//                        ranks (P2P loads, summed in rank order -> bit-identical on every replica), applies the Adagrad
//                        step to its slice of parameters and state (the state is therefore SHARDED: 1/world of it is
//                        touched per rank) and pushes the new parameter values into every peer's staging buffer (P2P
//                        stores): the all-gather moves updated WEIGHTS, the gradient never makes a second trip
//   barrier B            all pushes have landed
//   apply                each rank copies the other owners' slices from its staging buffer into its table (+ bf16
//                        mirror); the summed loss is read from the peers directly
//
// All four are plain kernels with peer pointers, so the whole step (compute + exchange) is ONE CUDA graph.  Per rank and
// step 2 * (world-1)/world * table bytes cross NVLink -- what a ring all-reduce moves -- with no second kernel pass over
// the gradient and a world-times smaller optimizer pass.
//
// Barriers: signal pad of rank k = uint32 slots [world]; at epoch e rank r stores e into slot r of every peer's pad
// (release, system scope) and spins until all its own slots reached e (acquire, system scope).  Epochs only grow, so a
// fast peer that already entered the next barrier (slot = e + 1) still satisfies ">= e"; no reset, no ABA.
#include <cstdint>

#include <cuda_bf16.h>

#include "common.cuh"

namespace kgeb {

constexpr int kMaxWorld = 16;

struct PeerPtrs {
  void* p[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

// <<<1, world>>>: thread t talks to peer t.  *epoch (local) is the last completed barrier number.
__global__ void p2p_barrier_kernel(PeerPtrs pads, int rank, int world, uint32_t* epoch, uint32_t* timeout_flag) {
  const int t = threadIdx.x;
  const uint32_t e = *epoch + 1;
  __threadfence_system();   // everything this GPU wrote before the barrier (incl. peer stores) is visible system-wide
  if (t < world) {
    st_release_sys(reinterpret_cast<uint32_t*>(pads.p[t]) + rank, e);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(pads.p[rank]) + t;
    // bounded spin (~10 s): a peer that died must not hang this GPU until the watchdog of the job runner fires
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      __nanosleep(64);
      if (++spins > (1ll << 23)) {
        *timeout_flag = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (t == 0) *epoch = e;
}

__host__ __device__ inline void slice_of(int64_t numel, int world, int r, int64_t& lo, int64_t& hi) {
  const int64_t per = ((numel + 4 * (int64_t)world - 1) / (4 * (int64_t)world)) * 4;   // multiple of 4 elements
  lo = per * r < numel ? per * r : numel;
  hi = lo + per < numel ? lo + per : numel;
}

// reduce + Adagrad on the owner's slice, push of the new values to every peer's staging buffer.
// grads.p[k] / stage.p[k]: rank k's buffers of this table (float, numel each, 16-byte aligned).
__global__ void __launch_bounds__(256)
p2p_adagrad_kernel(PeerPtrs grads, PeerPtrs stage, int rank, int world, float* __restrict__ W, float* __restrict__ state,
                   __nv_bfloat16* __restrict__ mirror, int64_t numel, float clr, float eps) {
  int64_t lo, hi;
  slice_of(numel, world, rank, lo, hi);
  const int64_t n4 = (hi - lo) / 4;           // full float4 groups; the slice start is a multiple of 4
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = lo + q * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < world; ++k) {         // fixed order: the same sum on whichever rank owns the slice
      const float4 h = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grads.p[k]) + i);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
    }
    float4 w = *reinterpret_cast<float4*>(W + i);
    float4 s = *reinterpret_cast<float4*>(state + i);
    s.x = fmaf(g.x, g.x, s.x); s.y = fmaf(g.y, g.y, s.y); s.z = fmaf(g.z, g.z, s.z); s.w = fmaf(g.w, g.w, s.w);
    w.x -= clr * g.x / (sqrtf(s.x) + eps);
    w.y -= clr * g.y / (sqrtf(s.y) + eps);
    w.z -= clr * g.z / (sqrtf(s.z) + eps);
    w.w -= clr * g.w / (sqrtf(s.w) + eps);
    *reinterpret_cast<float4*>(W + i) = w;
    *reinterpret_cast<float4*>(state + i) = s;
    if (mirror) {
      __nv_bfloat162 a = __floats2bfloat162_rn(w.x, w.y), b = __floats2bfloat162_rn(w.z, w.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(mirror + i) = o;
    }
    for (int k = 0; k < world; ++k)
      if (k != rank) *reinterpret_cast<float4*>(reinterpret_cast<float*>(stage.p[k]) + i) = w;
  }
  // tail of the LAST slice only (numel % 4 elements), by one thread
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = lo + n4 * 4; i < hi; ++i) {
      float g = 0.f;
      for (int k = 0; k < world; ++k) g += reinterpret_cast<const float*>(grads.p[k])[i];
      const float s = fmaf(g, g, state[i]);
      state[i] = s;
      const float w = W[i] - clr * g / (sqrtf(s) + eps);
      W[i] = w;
      if (mirror) mirror[i] = __float2bfloat16_rn(w);
      for (int k = 0; k < world; ++k)
        if (k != rank) reinterpret_cast<float*>(stage.p[k])[i] = w;
    }
  }
}

// after barrier B: the slices owned by the other ranks, from the local staging buffer into the table (+ mirror)
__global__ void __launch_bounds__(256)
p2p_apply_kernel(const float* __restrict__ stage, int rank, int world, float* __restrict__ W,
                 __nv_bfloat16* __restrict__ mirror, int64_t numel) {
  int64_t lo, hi;
  slice_of(numel, world, rank, lo, hi);
  const int64_t n4 = numel / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q * 4;
    if (i >= lo && i < hi) continue;            // own slice: already updated in place (slice bounds are multiples of 4)
    const float4 w = *reinterpret_cast<const float4*>(stage + i);
    *reinterpret_cast<float4*>(W + i) = w;
    if (mirror) {
      __nv_bfloat162 a = __floats2bfloat162_rn(w.x, w.y), b = __floats2bfloat162_rn(w.z, w.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(mirror + i) = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 * 4; i < numel; ++i) {
      if (i >= lo && i < hi) continue;
      W[i] = stage[i];
      if (mirror) mirror[i] = __float2bfloat16_rn(stage[i]);
    }
  }
}

// out[0] = sum over ranks of values.p[k][0], in rank order (the loss of the global batch)
__global__ void p2p_sum_scalar_kernel(PeerPtrs values, int world, float* out) {
  float a = 0.f;
  for (int k = 0; k < world; ++k) a += *reinterpret_cast<const float*>(values.p[k]);
  *out = a;
}

static int fill_ptrs(PeerPtrs& pp, const void* const* host_ptrs, int world) {
  for (int k = 0; k < kMaxWorld; ++k) pp.p[k] = k < world ? const_cast<void*>(host_ptrs[k]) : nullptr;
  return 0;
}

}  // namespace kgeb

using namespace kgeb;

extern "C" {

int kgeb_p2p_barrier(const void* const* peer_signal_pads, int rank, int world, uint32_t* epoch, uint32_t* timeout_flag,
                     void* stream) {
  KGEB_REQUIRE(peer_signal_pads && epoch && timeout_flag && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
               "p2p_barrier: bad arguments");
  PeerPtrs pads;
  fill_ptrs(pads, peer_signal_pads, world);
  p2p_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(pads, rank, world, epoch, timeout_flag);
  KGEB_LAUNCH_CHECK("p2p_barrier");
  return KGEB_OK;
}

int kgeb_p2p_adagrad(const void* const* peer_grads, const void* const* peer_stage, int rank, int world, float* W,
                     float* state, void* bf16_mirror, int64_t numel, float clr, float eps, void* stream) {
  KGEB_REQUIRE(peer_grads && peer_stage && W && state && numel >= 0 && world >= 1 && world <= kMaxWorld && rank >= 0 &&
                   rank < world,
               "p2p_adagrad: bad arguments");
  PeerPtrs g, s;
  fill_ptrs(g, peer_grads, world);
  fill_ptrs(s, peer_stage, world);
  uintptr_t bits = reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(state) | reinterpret_cast<uintptr_t>(bf16_mirror);
  for (int k = 0; k < world; ++k) bits |= reinterpret_cast<uintptr_t>(g.p[k]) | reinterpret_cast<uintptr_t>(s.p[k]);
  KGEB_REQUIRE((bits & 15) == 0, "p2p_adagrad: pointers must be 16-byte aligned");
  if (numel == 0) return KGEB_OK;
  int64_t lo, hi;
  slice_of(numel, world, rank, lo, hi);
  int64_t blocks = ((hi - lo) / 4 + 255) / 256 + 1;
  const int grid = (int)(blocks > (int64_t)kNumSMs * 4 ? (int64_t)kNumSMs * 4 : blocks);
  p2p_adagrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, s, rank, world, W, state,
                                                         reinterpret_cast<__nv_bfloat16*>(bf16_mirror), numel, clr, eps);
  KGEB_LAUNCH_CHECK("p2p_adagrad");
  return KGEB_OK;
}

int kgeb_p2p_apply(const float* stage, int rank, int world, float* W, void* bf16_mirror, int64_t numel, void* stream) {
  KGEB_REQUIRE(stage && W && numel >= 0 && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
               "p2p_apply: bad arguments");
  if (numel == 0 || world == 1) return KGEB_OK;
  int64_t blocks = (numel / 4 + 255) / 256 + 1;
  const int grid = (int)(blocks > (int64_t)kNumSMs * 4 ? (int64_t)kNumSMs * 4 : blocks);
  p2p_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(stage, rank, world, W, reinterpret_cast<__nv_bfloat16*>(bf16_mirror),
                                                       numel);
  KGEB_LAUNCH_CHECK("p2p_apply");
  return KGEB_OK;
}

int kgeb_p2p_sum_scalar(const void* const* peer_values, int world, float* out, void* stream) {
  KGEB_REQUIRE(peer_values && out && world >= 1 && world <= kMaxWorld, "p2p_sum_scalar: bad arguments");
  PeerPtrs v;
  fill_ptrs(v, peer_values, world);
  p2p_sum_scalar_kernel<<<1, 1, 0, as_stream(stream)>>>(v, world, out);
  KGEB_LAUNCH_CHECK("p2p_sum_scalar");
  return KGEB_OK;
}

}  // extern "C"

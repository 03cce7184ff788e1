// Issue-rate microbenchmark for the instructions the tile epilogues are made of (sm_100a):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mufu_bench tools/mufu_bench.cu && gpurun_out/mufu_bench
// Prints results per SM per clock (lanes/clk/SM; a packed x2 instruction counts two results per lane).
// Every thread runs ILP independent dependency chains so latency is hidden; 1024 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { EX2_F32, RCP_F32, LG2_F32, TANH_F32, EX2_F16X2, EX2_BF16X2, TANH_F16X2, TANH_BF16X2, FMA_F32, FMA_F32X2, MUL_F32 };

template <int OP>
__device__ __forceinline__ void step(uint32_t& a, uint64_t& w) {
  if (OP == EX2_F32) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a));
  if (OP == RCP_F32) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(a));
  if (OP == LG2_F32) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+r"(a));
  if (OP == TANH_F32) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a));
  if (OP == EX2_F16X2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a));
  if (OP == EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a));
  if (OP == TANH_F16X2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a));
  if (OP == TANH_BF16X2) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a));
  if (OP == FMA_F32) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(a));
  if (OP == MUL_F32) asm volatile("mul.rn.f32 %0, %0, %0;" : "+r"(a));
  if (OP == FMA_F32X2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(w));
}

template <int OP, int ILP>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* clk, int iters, uint32_t seed) {
  uint32_t a[ILP];
  uint64_t w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    a[i] = seed + threadIdx.x + i;
    w[i] = ((uint64_t)a[i] << 32) | a[i];
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) step<OP>(a[i], w[i]);
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int results_per_instr, int sms) {
  constexpr int ILP = 8;
  const int iters = 4096;
  uint32_t* out;
  long long* clk;
  cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
  cudaMalloc(&clk, sizeof(long long) * sms);
  bench<OP, ILP><<<sms, 1024>>>(out, clk, 64, 0x3c003c00u);
  bench<OP, ILP><<<sms, 1024>>>(out, clk, iters, 0x3c003c00u);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-14s failed: %s\n", name, cudaGetErrorString(e)); return; }
  long long c0;
  cudaMemcpy(&c0, clk, sizeof(c0), cudaMemcpyDeviceToHost);
  const double per_clk = (double)iters * ILP * 1024 * results_per_instr / (double)c0;
  printf("%-14s %8.2f results/clk/SM  (%6.2f warp-instr/clk/SM)\n", name, per_clk, per_clk / 32 / results_per_instr);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, sms);
  run<EX2_F32>("ex2.f32", 1, sms);
  run<RCP_F32>("rcp.f32", 1, sms);
  run<LG2_F32>("lg2.f32", 1, sms);
  run<TANH_F32>("tanh.f32", 1, sms);
  run<EX2_F16X2>("ex2.f16x2", 2, sms);
  run<EX2_BF16X2>("ex2.bf16x2", 2, sms);
  run<TANH_F16X2>("tanh.f16x2", 2, sms);
  run<TANH_BF16X2>("tanh.bf16x2", 2, sms);
  run<FMA_F32>("fma.f32", 1, sms);
  run<MUL_F32>("mul.f32", 1, sms);
  run<FMA_F32X2>("fma.f32x2", 2, sms);
  return 0;
}

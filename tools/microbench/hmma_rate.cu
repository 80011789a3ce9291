// Throughput and latency of the warp-level mma.sync.m16n8k16 (f16 x f16 -> f32, SASS HMMA.16816.F32) on sm_100a.
// The expected-OKS decoder's prefilter (pp_decode_mma.cuh) issues ~120 of these per heatmap from a single warp, so
// what matters is the issue rate per SM sub-partition at 1..4 warps per scheduler, not the tcgen05 peak.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kIters = 2048;

template <int ACC>   // independent accumulators per warp (ACC = 1: a dependent chain -> latency)
__global__ void hmma_kernel(float* out, unsigned seed) {
  unsigned a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed + 4, seed + 5};
  float c[ACC][4];
  for (int i = 0; i < ACC; ++i)
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) mma16816(c[i], a, b);
  }
  float s = 0.f;
  for (int i = 0; i < ACC; ++i)
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
double run(float* out, int blocks, int threads) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  hmma_kernel<ACC><<<blocks, threads>>>(out, 0u);
  cudaEventRecord(a);
  hmma_kernel<ACC><<<blocks, threads>>>(out, 0u);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 64 * 32);
  for (int warps_per_sm : {4, 8, 12, 16, 32}) {
    const int threads = 128, blocks = sms * warps_per_sm / 4;
    const double ms1 = run<1>(out, blocks, threads), ms8 = run<8>(out, blocks, threads);
    const double clk = pr.clockRate * 1e3;
    const double n1 = static_cast<double>(blocks) * 4 * kIters, n8 = n1 * 8;
    printf("warps/SM %2d  chain: %.1f clk / HMMA / warp   8 independent: %.3f HMMA/clk/SM (%.1f clk per HMMA per scheduler, %.0f dense TFLOP/s)\n",
           warps_per_sm, ms1 * 1e-3 * clk / kIters, n8 / (ms8 * 1e-3) / clk / sms,
           (ms8 * 1e-3) * clk * sms * 4 / n8, n8 * 4096.0 / (ms8 * 1e-3) / 1e12);
  }
  return 0;
}

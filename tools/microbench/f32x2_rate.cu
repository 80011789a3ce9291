// Throughput of scalar vs two-wide float32 arithmetic on sm_100a (FFMA vs FFMA2, FADD vs FADD2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int kAcc = 8, kIters = 4096;

template <int MODE>
__global__ void rate_kernel(float* out, float x, float y) {
  float a[kAcc];
  f2 p[kAcc];
  for (int i = 0; i < kAcc; ++i) { a[i] = threadIdx.x + i; p[i] = (static_cast<f2>(__float_as_uint(a[i])) << 32) | __float_as_uint(a[i] + 1.f); }
  const f2 xx = (static_cast<f2>(__float_as_uint(x)) << 32) | __float_as_uint(x);
  const f2 yy = (static_cast<f2>(__float_as_uint(y)) << 32) | __float_as_uint(y);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
      if (MODE == 0) a[i] = fma1(a[i], x, y);
      if (MODE == 1) p[i] = fma2(p[i], xx, yy);
      if (MODE == 2) a[i] = add1(a[i], x);
      if (MODE == 3) p[i] = add2(p[i], xx);
    }
  }
  float s = 0.f;
  for (int i = 0; i < kAcc; ++i) s += a[i] + __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(float* out, int blocks, int threads) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  rate_kernel<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(a);
  rate_kernel<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
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
  cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const char* names[4] = {"FFMA ", "FFMA2", "FADD ", "FADD2"};
  for (int warps_per_sm : {4, 8, 16, 32}) {
    const int threads = 128, blocks = sms * warps_per_sm / 4;
    double ms[4] = {run<0>(out, blocks, threads), run<1>(out, blocks, threads), run<2>(out, blocks, threads), run<3>(out, blocks, threads)};
    for (int m = 0; m < 4; ++m) {
      const double warp_instr = static_cast<double>(blocks) * (threads / 32) * kAcc * kIters;
      const double per_clk_sm = warp_instr / (ms[m] * 1e-3) / (pr.clockRate * 1e3) / sms;   // warp instructions / clk / SM
      printf("warps/SM %2d  %s  %.3f ms  %.2f warp-instr/clk/SM  %.1f flop-lanes/clk/SM\n", warps_per_sm, names[m], ms[m], per_clk_sm,
             per_clk_sm * 32 * ((m & 1) ? 2 : 1));
    }
  }
  return 0;
}

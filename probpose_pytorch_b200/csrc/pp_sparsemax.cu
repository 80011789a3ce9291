// Sparsemax-normalised head tail:
//     y = clamp(sparsemax(x / temperature over H*W) * normalize, 0, 1)        head.py:526-532, 237-245
// and its backward.  `sparsemax` is the Euclidean projection onto the simplex (Martins & Astudillo 2016;
// the reference imports it from the PyPI package sparsemax==0.1.9, which is not part of the reference tree):
//     z <- z - max(z);  tau = (sum_{i in S} z_i - 1) / |S|  with  S = {i : z_i > tau};  p = max(0, z - tau)
//     backward:  g_z = [p != 0] * (g_p - mean_{p != 0}(g_p))
//
// The package finds S by sorting; a sort of H*W values per heatmap is the wrong shape for a GPU.  S is found
// here by Michelot's fixed point instead: start from S0 = {z > -1} (tau >= -1 always, because the maximum is
// 0 after the shift and p_max <= 1), recompute tau from the current set, drop what falls below, stop when the
// set no longer shrinks.  The fixed point is the same set the sort yields; sums are carried in double, so tau
// is the correctly rounded value where the package's float32 cumulative sum carries ~1e-7 of noise.
//
// Layout: one CTA per heatmap (grid-stride), the shifted logits live in shared memory (or are re-read from
// global memory / L2 when H*W floats do not fit), every pass is a strided sweep + a block reduction.
// Algorithmic bytes: forward read 1 + write 1, backward read 2 + write 1 planes.  `aux` = (max, tau) per
// heatmap, written by the forward, lets the backward rebuild p exactly without storing it.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/probpose_b200.h"
#include "pp_common.cuh"

namespace {

using namespace pp;

constexpr int kSpThreads = 256;

struct SpReduce {
  double d[2][kSpThreads / 32];
  float f[kSpThreads / 32];
  double out_d[2];
  float out_f;
};

__device__ __forceinline__ float block_max(float v, SpReduce& r) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) r.f[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < kSpThreads / 32 ? r.f[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) r.out_f = t;
  }
  __syncthreads();
  return r.out_f;
}

__device__ __forceinline__ void block_sum2(double& a, double& b, SpReduce& r) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { r.d[0][warp] = a; r.d[1][warp] = b; }
  __syncthreads();
  if (warp == 0) {
    double x = lane < kSpThreads / 32 ? r.d[0][lane] : 0.0;
    double y = lane < kSpThreads / 32 ? r.d[1][lane] : 0.0;
    x = warp_sum(x);
    y = warp_sum(y);
    if (lane == 0) { r.out_d[0] = x; r.out_d[1] = y; }
  }
  __syncthreads();
  a = r.out_d[0];
  b = r.out_d[1];
}

// x / temperature rounded to the tensor dtype, as torch does before the normalisation layer sees it
template <typename T>
__device__ __forceinline__ float scaled_logit(const T* __restrict__ x, int i, float temperature) {
  return Elem<T>::to_f32(Elem<T>::from_f32(__fdiv_rn(Elem<T>::to_f32(x[i]), temperature)));
}

template <typename T, bool kSmem>
__global__ void __launch_bounds__(kSpThreads)
sparsemax_tail_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ aux, int64_t N, int HW,
                      float temperature, float normalize) {
  extern __shared__ __align__(16) float zs[];
  __shared__ SpReduce red;
  const int tid = threadIdx.x;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const T* xp = x + hm * HW;
    T* yp = y + hm * HW;
    float vmax = -INFINITY;
    for (int i = tid; i < HW; i += kSpThreads) {
      const float z = scaled_logit(xp, i, temperature);
      if (kSmem) zs[i] = z;
      vmax = fmaxf(vmax, z);
    }
    vmax = block_max(vmax, red);
    auto shifted = [&](int i) -> float {
      return __fsub_rn(kSmem ? zs[i] : scaled_logit(xp, i, temperature), vmax);
    };
    // Michelot: tau only grows, the set only shrinks; at most HW rounds, a handful in practice
    float tau = -1.0f;
    double prev_cnt = -1.0;
    for (int round = 0; round <= HW; ++round) {
      double s = 0.0, c = 0.0;
      for (int i = tid; i < HW; i += kSpThreads) {
        const float z = shifted(i);
        if (z > tau) { s += static_cast<double>(z); c += 1.0; }
      }
      block_sum2(s, c, red);
      tau = __fdiv_rn(static_cast<float>(s - 1.0), static_cast<float>(c));
      if (c == prev_cnt) break;
      prev_cnt = c;
    }
    for (int i = tid; i < HW; i += kSpThreads) {
      float p = fmaxf(0.0f, __fsub_rn(shifted(i), tau));
      p = Elem<T>::to_f32(Elem<T>::from_f32(p));
      p = Elem<T>::to_f32(Elem<T>::from_f32(__fmul_rn(p, normalize)));
      yp[i] = Elem<T>::from_f32(fminf(fmaxf(p, 0.0f), 1.0f));
    }
    if (tid == 0 && aux) { aux[hm * 2] = vmax; aux[hm * 2 + 1] = tau; }
    __syncthreads();   // zs and the reduction slots are rewritten by the next heatmap
  }
}

template <typename T>
__global__ void __launch_bounds__(kSpThreads)
sparsemax_tail_backward_kernel(const T* __restrict__ x, const T* __restrict__ gy, const float* __restrict__ aux,
                               T* __restrict__ gx, int64_t N, int HW, float temperature, float normalize) {
  __shared__ SpReduce red;
  const int tid = threadIdx.x;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const T* xp = x + hm * HW;
    const T* gp = gy + hm * HW;
    T* op = gx + hm * HW;
    const float vmax = aux[hm * 2], tau = aux[hm * 2 + 1];
    // gradient that reaches the projection: through the clamp (inclusive on [0, 1]) and the scale
    auto upstream = [&](int i, bool* nz) -> float {
      float p = fmaxf(0.0f, __fsub_rn(__fsub_rn(scaled_logit(xp, i, temperature), vmax), tau));
      p = Elem<T>::to_f32(Elem<T>::from_f32(p));
      *nz = p != 0.0f;
      const float q = Elem<T>::to_f32(Elem<T>::from_f32(__fmul_rn(p, normalize)));
      const float g = (q >= 0.0f && q <= 1.0f) ? Elem<T>::to_f32(gp[i]) : 0.0f;
      return Elem<T>::to_f32(Elem<T>::from_f32(__fmul_rn(g, normalize)));
    };
    double s = 0.0, c = 0.0;
    for (int i = tid; i < HW; i += kSpThreads) {
      bool nz;
      const float g = upstream(i, &nz);
      if (nz) { s += static_cast<double>(g); c += 1.0; }
    }
    block_sum2(s, c, red);
    const float mean = static_cast<float>(s / c);
    for (int i = tid; i < HW; i += kSpThreads) {
      bool nz;
      const float g = upstream(i, &nz);
      const float gz = nz ? __fsub_rn(g, mean) : 0.0f;
      op[i] = Elem<T>::from_f32(__fdiv_rn(Elem<T>::to_f32(Elem<T>::from_f32(gz)), temperature));
    }
    __syncthreads();
  }
}


// ---- fast path: H*W <= 32 * 256, 128-bit global accesses, shrinking candidate set ------------------------------
// Michelot's tau only grows, so a pixel that has left the candidate set never returns: every thread keeps
// a bit mask of its still-alive pixels (pixel tid + 256 j  <->  bit j) and later rounds only touch those.
// Rounds cost one barrier each (per-warp partials in parity-alternating slots, summed by every thread in
// a fixed order: deterministic).  When 1 / temperature is a power of two the division is an exact multiply.
constexpr int kSpMaxFast = 32 * kSpThreads;

struct SpSlots {
  double s[2][kSpThreads / 32];
  int c[2][kSpThreads / 32];
  float m[kSpThreads / 32];
};

template <typename T>
__device__ __forceinline__ float scale_logit(float v, float temperature, float inv_t, bool mul_ok) {
  return Elem<T>::to_f32(Elem<T>::from_f32(mul_ok ? __fmul_rn(v, inv_t) : __fdiv_rn(v, temperature)));
}

__device__ __forceinline__ void block_sum_sc(double& s, int& c, SpSlots& r, int& parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s = warp_sum(s);
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) { r.s[parity][warp] = s; r.c[parity][warp] = c; }
  __syncthreads();
  // every warp sums the per-warp partials with the same 3-step butterfly: same order, same bits
  constexpr int kW = kSpThreads / 32;
  s = r.s[parity][lane & (kW - 1)];
  c = r.c[parity][lane & (kW - 1)];
#pragma unroll
  for (int o = kW / 2; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  parity ^= 1;
}

// tau = (sum - 1) / count, rounded once to float32 (the sum is exact to double precision)
__device__ __forceinline__ float tau_of(double s, int c) {
  return __fdiv_rn(static_cast<float>(s - 1.0), static_cast<float>(c));
}

template <typename T>
__global__ void __launch_bounds__(kSpThreads)
sparsemax_tail_fast_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ aux, int64_t N, int HW,
                           float temperature, float inv_t, bool mul_ok, float normalize) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ __align__(16) float zs[];
  __shared__ SpSlots red;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = HW / V, E = (HW + kSpThreads - 1) / kSpThreads;
  int parity = 0;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const T* xp = x + hm * HW;
    T* yp = y + hm * HW;
    float vmax = -INFINITY;
    for (int v = tid; v < nvec; v += kSpThreads) {
      float f[V];
      unpack(ldg_stream_128(xp + v * V), f, T());
#pragma unroll
      for (int j = 0; j < V; ++j) {
        f[j] = scale_logit<T>(f[j], temperature, inv_t, mul_ok);
        vmax = fmaxf(vmax, f[j]);
      }
#pragma unroll
      for (int j = 0; j < V; j += 4) *reinterpret_cast<float4*>(zs + v * V + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
    vmax = warp_max(vmax);
    if (lane == 0) red.m[warp] = vmax;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kSpThreads / 32; ++w) vmax = fmaxf(vmax, red.m[w]);

    // round 0: candidates {z - max > -1}
    unsigned mask = 0;
    double s = 0.0;
    int c = 0;
    for (int j = 0; j < E; ++j) {
      const int i = tid + j * kSpThreads;
      if (i < HW) {
        const float z = __fsub_rn(zs[i], vmax);
        if (z > -1.0f) { mask |= 1u << j; s += static_cast<double>(z); ++c; }
      }
    }
    float tau;
    int prev = -1;
    while (true) {
      block_sum_sc(s, c, red, parity);
      tau = tau_of(s, c);
      if (c == prev) break;
      prev = c;
      unsigned m = mask;
      mask = 0;
      s = 0.0;
      c = 0;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const float z = __fsub_rn(zs[tid + j * kSpThreads], vmax);
        if (z > tau) { mask |= 1u << j; s += static_cast<double>(z); ++c; }
      }
    }
    for (int v = tid; v < nvec; v += kSpThreads) {
      float f[V];
#pragma unroll
      for (int j = 0; j < V; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(zs + v * V + j);
        f[j] = q.x; f[j + 1] = q.y; f[j + 2] = q.z; f[j + 3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float p = fmaxf(0.0f, __fsub_rn(__fsub_rn(f[j], vmax), tau));
        p = Elem<T>::to_f32(Elem<T>::from_f32(p));
        p = Elem<T>::to_f32(Elem<T>::from_f32(__fmul_rn(p, normalize)));
        f[j] = fminf(fmaxf(p, 0.0f), 1.0f);
      }
      stg_stream_128(yp + v * V, pack(f, T()));
    }
    if (tid == 0 && aux) { aux[hm * 2] = vmax; aux[hm * 2 + 1] = tau; }
    __syncthreads();   // zs and red.m are rewritten by the next heatmap
  }
}

// backward, fast path: one sweep reads x and grad_y (128-bit), parks the gradient that reaches the projection in
// shared memory and the support in a bit mask (vector v = tid + 256 jv, component q  <->  bit jv * V + q);
// the second sweep subtracts the support's mean and writes.
template <typename T>
__global__ void __launch_bounds__(kSpThreads)
sparsemax_tail_backward_fast_kernel(const T* __restrict__ x, const T* __restrict__ gy, const float* __restrict__ aux,
                                    T* __restrict__ gx, int64_t N, int HW, float temperature, float inv_t, bool mul_ok,
                                    float normalize) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ __align__(16) float gs[];
  __shared__ SpSlots red;
  const int tid = threadIdx.x;
  const int nvec = HW / V;
  int parity = 0;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const T* xp = x + hm * HW;
    const T* gp = gy + hm * HW;
    T* op = gx + hm * HW;
    const float vmax = aux[hm * 2], tau = aux[hm * 2 + 1];
    unsigned mask = 0;
    double s = 0.0;
    int c = 0, bit = 0;
    for (int v = tid; v < nvec; v += kSpThreads, bit += V) {
      float f[V], g[V];
      unpack(ldg_stream_128(xp + v * V), f, T());
      unpack(ldg_stream_128(gp + v * V), g, T());
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float p = fmaxf(0.0f, __fsub_rn(__fsub_rn(scale_logit<T>(f[j], temperature, inv_t, mul_ok), vmax), tau));
        p = Elem<T>::to_f32(Elem<T>::from_f32(p));
        const float q = Elem<T>::to_f32(Elem<T>::from_f32(__fmul_rn(p, normalize)));
        const float u = (q >= 0.0f && q <= 1.0f) ? g[j] : 0.0f;
        g[j] = Elem<T>::to_f32(Elem<T>::from_f32(__fmul_rn(u, normalize)));
        if (p != 0.0f) { mask |= 1u << (bit + j); s += static_cast<double>(g[j]); ++c; }
      }
#pragma unroll
      for (int j = 0; j < V; j += 4) *reinterpret_cast<float4*>(gs + v * V + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
    }
    block_sum_sc(s, c, red, parity);
    const float mean = static_cast<float>(s / static_cast<double>(c));
    bit = 0;
    for (int v = tid; v < nvec; v += kSpThreads, bit += V) {
      float g[V];
#pragma unroll
      for (int j = 0; j < V; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(gs + v * V + j);
        g[j] = q.x; g[j + 1] = q.y; g[j + 2] = q.z; g[j + 3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float gz = ((mask >> (bit + j)) & 1u) ? Elem<T>::to_f32(Elem<T>::from_f32(__fsub_rn(g[j], mean))) : 0.0f;
        g[j] = mul_ok ? __fmul_rn(gz, inv_t) : __fdiv_rn(gz, temperature);
      }
      stg_stream_128(op + v * V, pack(g, T()));
    }
    // each thread re-reads only what it wrote to gs; the reduction slots alternate: no barrier needed here
  }
}

template <typename T>
bool fast_ok(const void* a, const void* b, const void* c, int HW) {
  return HW <= kSpMaxFast && HW % Elem<T>::kVec == 0 && pp_aligned16(a) && pp_aligned16(b) && (!c || pp_aligned16(c)) &&
         (static_cast<int64_t>(HW) * sizeof(T)) % 16 == 0;
}

// 1 / t exactly representable and x / t == x * (1 / t) for every x: t is a power of two
bool reciprocal_exact(float t, float* inv) {
  int e = 0;
  const float m = std::frexp(t, &e);
  *inv = 1.0f / t;
  return m == 0.5f && e > -100 && e < 100;
}

int check_args(const char* fn, int dtype, int64_t N, int64_t HW, float temperature) {
  PP_REQUIRE(dtype == PP_F32 || dtype == PP_BF16, PP_ERR_INVALID_ARG, "%s: unsupported dtype %d", fn, dtype);
  PP_REQUIRE(N >= 0 && HW > 0 && HW < (1ll << 30), PP_ERR_INVALID_ARG, "%s: bad shape N=%lld HW=%lld", fn,
             static_cast<long long>(N), static_cast<long long>(HW));
  PP_REQUIRE(temperature > 0.0f, PP_ERR_INVALID_ARG, "%s: temperature must be positive", fn);
  return PP_OK;
}

template <typename T>
int launch_forward(const void* x, void* y, float* aux, int64_t N, int HW, float temperature, float normalize,
                   cudaStream_t st) {
  const size_t smem = sizeof(float) * static_cast<size_t>(HW);
  const bool in_smem = static_cast<int64_t>(smem) + 2048 <= pp_smem_optin();
  int per_sm = 1;
  if (fast_ok<T>(x, y, nullptr, HW) && in_smem) {
    auto kern = sparsemax_tail_fast_kernel<T>;
    if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(kern), kSpThreads, smem, &per_sm)) return rc;
    float inv_t = 1.0f;
    const bool mul_ok = reciprocal_exact(temperature, &inv_t);
    const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * per_sm));
    kern<<<grid, kSpThreads, smem, st>>>(static_cast<const T*>(x), static_cast<T*>(y), aux, N, HW, temperature, inv_t, mul_ok,
                                         normalize);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  }
  if (in_smem) {
    if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(sparsemax_tail_kernel<T, true>), kSpThreads, smem, &per_sm))
      return rc;
  } else {
    per_sm = 8;
  }
  const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * per_sm));
  if (in_smem)
    sparsemax_tail_kernel<T, true><<<grid, kSpThreads, smem, st>>>(static_cast<const T*>(x), static_cast<T*>(y), aux, N, HW,
                                                                  temperature, normalize);
  else
    sparsemax_tail_kernel<T, false><<<grid, kSpThreads, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(y), aux, N, HW,
                                                                temperature, normalize);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

template <typename T>
int launch_backward(const void* x, const void* gy, const float* aux, void* gx, int64_t N, int HW, float temperature,
                    float normalize, cudaStream_t st) {
  const size_t smem = sizeof(float) * static_cast<size_t>(HW);
  if (fast_ok<T>(x, gy, gx, HW) && static_cast<int64_t>(smem) + 2048 <= pp_smem_optin()) {
    auto kern = sparsemax_tail_backward_fast_kernel<T>;
    int per_sm = 1;
    if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(kern), kSpThreads, smem, &per_sm)) return rc;
    float inv_t = 1.0f;
    const bool mul_ok = reciprocal_exact(temperature, &inv_t);
    const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * per_sm));
    kern<<<grid, kSpThreads, smem, st>>>(static_cast<const T*>(x), static_cast<const T*>(gy), aux, static_cast<T*>(gx), N, HW,
                                         temperature, inv_t, mul_ok, normalize);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  }
  const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * 8));
  sparsemax_tail_backward_kernel<T><<<grid, kSpThreads, 0, st>>>(static_cast<const T*>(x), static_cast<const T*>(gy), aux,
                                                                 static_cast<T*>(gx), N, HW, temperature, normalize);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // namespace

extern "C" {

PP_API int pp_sparsemax_tail(const void* x, void* y, float* aux, int dtype, int64_t n_heatmaps, int64_t hw,
                             float temperature, float normalize, void* stream) {
  if (int rc = check_args("pp_sparsemax_tail", dtype, n_heatmaps, hw, temperature)) return rc;
  if (n_heatmaps == 0) return PP_OK;
  PP_REQUIRE(x && y, PP_ERR_INVALID_ARG, "pp_sparsemax_tail: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  return dtype == PP_F32
             ? launch_forward<float>(x, y, aux, n_heatmaps, static_cast<int>(hw), temperature, normalize, st)
             : launch_forward<__nv_bfloat16>(x, y, aux, n_heatmaps, static_cast<int>(hw), temperature, normalize, st);
}

PP_API int pp_sparsemax_tail_backward(const void* x, const void* grad_y, const float* aux, void* grad_x, int dtype,
                                      int64_t n_heatmaps, int64_t hw, float temperature, float normalize, void* stream) {
  if (int rc = check_args("pp_sparsemax_tail_backward", dtype, n_heatmaps, hw, temperature)) return rc;
  if (n_heatmaps == 0) return PP_OK;
  PP_REQUIRE(x && grad_y && aux && grad_x, PP_ERR_INVALID_ARG, "pp_sparsemax_tail_backward: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  const int HW = static_cast<int>(hw);
  return dtype == PP_F32
             ? launch_backward<float>(x, grad_y, aux, grad_x, n_heatmaps, HW, temperature, normalize, st)
             : launch_backward<__nv_bfloat16>(x, grad_y, aux, grad_x, n_heatmaps, HW, temperature, normalize, st);
}

}  // extern "C"

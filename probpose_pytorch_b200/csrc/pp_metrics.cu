// Validation metrics on the device: PCK from decoded coordinates
// (_calc_distances / _distance_acc, heatmap.py:55-111; keypoint_pck_accuracy, loss.py:825-866) and the
// mask-select metrics of ProbPoseLoss (get_binary_accuracy / get_mae, loss.py:653-712).
// The heavy half of pose_pck_accuracy (two plain argmax passes over (N, K, H, W), loss.py:817-818) is
// pp_heatmap_maximum; what is here works on (N, K)-sized arrays and runs in a single CTA.

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/probpose_b200.h"
#include "pp_common.cuh"

namespace {

using namespace pp;

constexpr int kMetThreads = 256;
constexpr int kMaxThresholds = 32;

// normalised distance of one keypoint, rounded to float32 like the reference's store into a float32 array;
// arithmetic in the promoted dtype of (float32 coordinates, normalisation factor)
template <typename NormT>
__device__ __forceinline__ bool pck_distance(const float* pred, const float* gt, const uint8_t* mask, const NormT* norm,
                                             int n, int k, int K, float* out) {
  NormT nx = norm[n * 2], ny = norm[n * 2 + 1];
  if (!mask[n * K + k] || nx == NormT(0) || ny == NormT(0)) return false;     // heatmap.py:79-81
  if (nx <= NormT(0)) nx = NormT(1e6);                                          // heatmap.py:85
  if (ny <= NormT(0)) ny = NormT(1e6);
  const float dx = __fsub_rn(pred[(n * K + k) * 2], gt[(n * K + k) * 2]);
  const float dy = __fsub_rn(pred[(n * K + k) * 2 + 1], gt[(n * K + k) * 2 + 1]);
  if constexpr (sizeof(NormT) == 8) {
    const double qx = __ddiv_rn(static_cast<double>(dx), nx), qy = __ddiv_rn(static_cast<double>(dy), ny);
    *out = static_cast<float>(__dsqrt_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy))));
  } else {
    const float qx = __fdiv_rn(dx, nx), qy = __fdiv_rn(dy, ny);
    *out = __fsqrt_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)));
  }
  return true;
}

template <typename NormT>
__global__ void __launch_bounds__(kMetThreads)
pck_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const uint8_t* __restrict__ mask,
           const NormT* __restrict__ norm, int N, int K, float thr, double* __restrict__ acc, double* __restrict__ avg,
           int* __restrict__ cnt, float* __restrict__ distances) {
  __shared__ double s_sum[kMetThreads / 32];
  __shared__ int s_cnt[kMetThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0;     // this warp's keypoints: sum of acc >= 0 and their number (lane 0)
  int valid_k = 0;
  for (int k = warp; k < K; k += kMetThreads / 32) {
    int valid = 0, below = 0;
    for (int n = lane; n < N; n += 32) {
      float d = -1.0f;
      if (pck_distance(pred, gt, mask, norm, n, k, K, &d)) {
        ++valid;
        below += d < thr;
      }
      if (distances) distances[static_cast<int64_t>(k) * N + n] = d;     // (K, N), heatmap.py:90
    }
    valid = __reduce_add_sync(0xffffffffu, valid);
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 0) {
      const double a = valid > 0 ? static_cast<double>(below) / static_cast<double>(valid) : -1.0;
      acc[k] = a;
      if (a >= 0.0) { sum += a; ++valid_k; }
    }
  }
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = valid_k; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int c = 0;
    for (int w = 0; w < kMetThreads / 32; ++w) { t += s_sum[w]; c += s_cnt[w]; }
    *avg = c > 0 ? t / c : 0.0;
    *cnt = c;
  }
}

__global__ void __launch_bounds__(kMetThreads)
binary_accuracy_kernel(const float* __restrict__ dt, const float* __restrict__ gt, const uint8_t* __restrict__ mask, int64_t n,
                       const double* __restrict__ thresholds, int n_thr, float* __restrict__ out, int64_t* __restrict__ counts_out) {
  __shared__ unsigned long long s_counts[kMaxThresholds + 1];
  __shared__ double s_thr[kMaxThresholds];
  if (threadIdx.x <= kMaxThresholds) s_counts[threadIdx.x] = 0;
  if (threadIdx.x < n_thr) s_thr[threadIdx.x] = thresholds[threadIdx.x];
  __syncthreads();
  int local[kMaxThresholds];
#pragma unroll
  for (int t = 0; t < kMaxThresholds; ++t) local[t] = 0;
  int samples = 0;
  for (int64_t i = threadIdx.x; i < n; i += kMetThreads) {
    if (!mask[i]) continue;
    ++samples;
    const double v = static_cast<double>(dt[i]);       // float32 > float64 threshold array: compared in double
    const bool g = gt[i] != 0.0f;
#pragma unroll
    for (int t = 0; t < kMaxThresholds; ++t)
      if (t < n_thr) local[t] += ((v > s_thr[t]) == g);
  }
#pragma unroll
  for (int t = 0; t < kMaxThresholds; ++t) {
    const int w = __reduce_add_sync(0xffffffffu, local[t]);
    if ((threadIdx.x & 31) == 0 && t < n_thr && w) atomicAdd(&s_counts[t], static_cast<unsigned long long>(w));
  }
  samples = __reduce_add_sync(0xffffffffu, samples);
  if ((threadIdx.x & 31) == 0 && samples) atomicAdd(&s_counts[kMaxThresholds], static_cast<unsigned long long>(samples));
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int t = 1; t < n_thr; ++t)
      if (s_counts[t] > s_counts[best]) best = t;      // np.argmax: first maximum
    const double total = static_cast<double>(s_counts[kMaxThresholds]);
    out[0] = static_cast<float>(static_cast<double>(s_counts[best]) / total);   // 0 / 0 -> nan, as numpy
    out[1] = static_cast<float>(s_thr[best]);
    if (counts_out) {
      for (int t = 0; t < n_thr; ++t) counts_out[t] = static_cast<int64_t>(s_counts[t]);
      counts_out[n_thr] = static_cast<int64_t>(s_counts[kMaxThresholds]);
    }
  }
}

__global__ void __launch_bounds__(kMetThreads)
masked_mae_kernel(const float* __restrict__ dt, const float* __restrict__ gt, const uint8_t* __restrict__ mask, int64_t n,
                  float* __restrict__ out) {
  __shared__ double s_sum[kMetThreads / 32];
  __shared__ long long s_cnt[kMetThreads / 32];
  double sum = 0.0;
  long long c = 0;
  for (int64_t i = threadIdx.x; i < n; i += kMetThreads)
    if (mask[i]) { sum += static_cast<double>(fabsf(__fsub_rn(dt[i], gt[i]))); ++c; }
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    long long k = 0;
    for (int w = 0; w < kMetThreads / 32; ++w) { t += s_sum[w]; k += s_cnt[w]; }
    out[0] = static_cast<float>(t / static_cast<double>(k));                  // empty selection -> nan, as numpy
  }
}

}  // namespace

extern "C" {

PP_API int pp_pck_accuracy(const float* pred, const float* gt, const uint8_t* mask, const void* norm_factor,
                           int32_t norm_dtype, int32_t N, int32_t K, double thr, double* acc, double* avg_acc,
                           int32_t* cnt, float* distances, pp_stream_t stream) {
  PP_REQUIRE(N >= 0 && K > 0, PP_ERR_INVALID_ARG, "pp_pck_accuracy: bad shape N=%d K=%d", N, K);
  PP_REQUIRE(norm_dtype == PP_F32 || norm_dtype == PP_F64, PP_ERR_INVALID_ARG, "pp_pck_accuracy: norm dtype %d", norm_dtype);
  PP_REQUIRE(acc && avg_acc && cnt, PP_ERR_INVALID_ARG, "pp_pck_accuracy: null output");
  PP_REQUIRE(N == 0 || (pred && gt && mask && norm_factor), PP_ERR_INVALID_ARG, "pp_pck_accuracy: null input");
  auto st = static_cast<cudaStream_t>(stream);
  const float thr32 = static_cast<float>(thr);     // a Python float meets a float32 array: compared in float32
  if (norm_dtype == PP_F64)
    pck_kernel<double><<<1, kMetThreads, 0, st>>>(pred, gt, mask, static_cast<const double*>(norm_factor), N, K, thr32, acc,
                                                  avg_acc, cnt, distances);
  else
    pck_kernel<float><<<1, kMetThreads, 0, st>>>(pred, gt, mask, static_cast<const float*>(norm_factor), N, K, thr32, acc,
                                                 avg_acc, cnt, distances);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

PP_API int pp_binary_accuracy(const float* dt, const float* gt, const uint8_t* mask, int64_t n, const double* thresholds,
                              int32_t n_thresholds, float* out, int64_t* counts, pp_stream_t stream) {
  PP_REQUIRE(n >= 0 && n_thresholds > 0 && n_thresholds <= kMaxThresholds, PP_ERR_INVALID_ARG,
             "pp_binary_accuracy: n=%lld, %d thresholds (max %d)", static_cast<long long>(n), n_thresholds, kMaxThresholds);
  PP_REQUIRE(out && thresholds && (n == 0 || (dt && gt && mask)), PP_ERR_INVALID_ARG, "pp_binary_accuracy: null pointer");
  binary_accuracy_kernel<<<1, kMetThreads, 0, static_cast<cudaStream_t>(stream)>>>(dt, gt, mask, n, thresholds, n_thresholds,
                                                                                    out, counts);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

PP_API int pp_masked_mae(const float* dt, const float* gt, const uint8_t* mask, int64_t n, float* out, pp_stream_t stream) {
  PP_REQUIRE(n >= 0 && out && (n == 0 || (dt && gt && mask)), PP_ERR_INVALID_ARG, "pp_masked_mae: bad arguments");
  masked_mae_kernel<<<1, kMetThreads, 0, static_cast<cudaStream_t>(stream)>>>(dt, gt, mask, n, out);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // extern "C"

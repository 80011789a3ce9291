// Per-keypoint training targets derived from decoded keypoints (reference: ProbPoseLoss._oks_from_heatmaps,
// loss.py:550-640, with compute_oks(use_area=False, per_kpt=True), loss.py:715-764, and
// ProbPoseLoss._error_from_heatmaps, loss.py:512-548).  The reference decodes both heatmap stacks on the
// host (a device->host copy of 2 B K H W floats per step) and then does this (B, K) arithmetic in NumPy;
// here the decode is pp_decode_argmax_dark and this kernel finishes the job on the device in float64.
#include "pp_common.cuh"

namespace {

__global__ void __launch_bounds__(128)
pose_targets_kernel(const double* __restrict__ gt_kp, const double* __restrict__ dt_kp, const float* __restrict__ weight,
                    const double* __restrict__ sigmas, int B, int K, double area_term, float* __restrict__ oks,
                    float* __restrict__ oks_weight, double* __restrict__ error) {
  __shared__ int any_valid;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (threadIdx.x == 0) any_valid = 0;
    __syncthreads();
    // a keypoint is valid when its visibility 2 * weight is positive (loss.py:597-606)
    if (weight != nullptr) {
      for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (weight[b * K + k] * 2.0f > 0.0f) any_valid = 1;
    }
    __syncthreads();
    const bool valid_sample = any_valid != 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const int i = b * K + k;
      double gx = gt_kp[2 * i], gy = gt_kp[2 * i + 1];
      const double dx0 = dt_kp[2 * i], dy0 = dt_kp[2 * i + 1];
      if (error) {  // NaN ground truth -> -1 (loss.py:541), then the Euclidean distance (loss.py:544)
        const double ex = (isnan(gx) ? -1.0 : gx) - dx0, ey = (isnan(gy) ? -1.0 : gy) - dy0;
        error[i] = sqrt(ex * ex + ey * ey);
      }
      if (oks) {
        const double w = static_cast<double>(weight[i]);
        if (isnan(gx)) gx = 0.0;  // loss.py:588
        if (isnan(gy)) gy = 0.0;
        const double xg = gx * w, yg = gy * w, xd = dx0 * w, yd = dy0 * w;  // loss.py:591-592
        float v = 0.0f;
        if (valid_sample && w * 2.0 > 0.0) {
          const double s2 = sigmas[k] * 2.0;
          const double vars = s2 * s2;
          const double ddx = xd - xg, ddy = yd - yg;
          const double e = (ddx * ddx + ddy * ddy) / vars / area_term / 2.0;  // loss.py:751-752
          v = static_cast<float>(exp(-e));
        }
        oks[i] = v;
      }
    }
    if (oks_weight && threadIdx.x == 0) oks_weight[b] = valid_sample ? 1.0f : 0.0f;
    __syncthreads();
  }
}

}  // namespace

extern "C" int pp_pose_targets(const double* gt_keypoints, const double* dt_keypoints, const float* weight,
                               const double* sigmas, int32_t B, int32_t K, double heatmap_w, double heatmap_h,
                               float* oks, float* oks_weight, double* error, pp_stream_t stream) {
  PP_REQUIRE(B >= 0 && K > 0, PP_ERR_INVALID_ARG, "pp_pose_targets: bad shape B=%d K=%d", B, K);
  if (B == 0) return PP_OK;
  PP_REQUIRE(gt_keypoints && dt_keypoints, PP_ERR_INVALID_ARG, "pp_pose_targets: null keypoints");
  PP_REQUIRE(!oks || (weight && sigmas), PP_ERR_INVALID_ARG, "pp_pose_targets: oks needs weight and sigmas");
  // tmparea = bbox[3] * bbox[2] * 0.53, plus np.spacing(1) (loss.py:751)
  const double area_term = heatmap_w * heatmap_h * 0.53 + 2.220446049250313e-16;
  const int grid = B < pp_sm_count() * 4 ? B : pp_sm_count() * 4;
  pose_targets_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(gt_keypoints, dt_keypoints, weight, sigmas, B, K,
                                                                          area_term, oks, oks_weight, error);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

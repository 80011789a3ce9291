/* Torch-free consumer of libprobpose_b200.so: plain C + the CUDA runtime only.  Allocates with cudaMalloc, builds the
 * per-codec constant tables on the host as include/probpose_b200.h documents them (heatmap.py:170-194), then runs
 * pp_encode -> pp_decode_expected (tensor-core kernel, via pp_oks_mma_table_build) -> pp_oks_loss_forward_encoded on
 * the GPU and checks the round trip.  Built and run by tests/test_gpu_parity.py::test_c_consumer_runs_kernels_without_torch
 * (gcc, -lcudart); proves INTEGRATION.md's claim that the library needs neither Python nor torch. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "probpose_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 10; } } while (0)
#define PP(x) do { int r_ = (x); if (r_ != PP_OK) { printf("%s -> %d: %s\n", #x, r_, pp_last_error_string()); return 11; } } while (0)

enum { B = 24, K = 17, H = 64, W = 48, IW = 192, IH = 256 };
static const double kSigmas[K] = {.026, .025, .025, .035, .035, .079, .079, .072, .072, .062, .062, .107, .107, .087, .087, .089, .089};

static void* to_device(const void* host, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
  if (host && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return d;
}

int main(void) {
  int32_t sms = 0, maj = 0, min = 0;
  int64_t smem = 0;
  PP(pp_device_info(&sms, &maj, &min, &smem));
  printf("device: %d SMs, sm_%d%d, %lld bytes of shared memory; library ABI %d, sources %s\n", sms, maj, min, (long long)smem,
         pp_version(), pp_source_hash());

  /* ---- per-codec tables: s_k = clip((2 sigma)^2 sqrt(H/1.25 W/1.25) 2, 0.55, 3), radius ceil(3 s), normalised taps */
  double two_s[K], *k2d = calloc((size_t)K * PP_OKS_TAPS * PP_OKS_TAPS, sizeof(double));
  float taps[K][PP_OKS_TAPS];
  int32_t radius[K], order[K], mma_index[K];
  memset(taps, 0, sizeof taps);
  for (int k = 0; k < K; ++k) {
    double s = pow(2 * kSigmas[k], 2) * sqrt(H / 1.25 * W / 1.25) * 2;
    s = s < 0.55 ? 0.55 : s > 3.0 ? 3.0 : s;
    two_s[k] = 2 * s;
    const int r = (int)ceil(3 * s), d = 2 * r + 1;
    radius[k] = r; order[k] = k; mma_index[k] = k;
    double sum1 = 0, sum2 = 0, one[PP_OKS_TAPS];
    for (int i = 0; i < d; ++i) { one[i] = exp(-(double)((i - r) * (i - r)) / (2 * s)); sum1 += one[i]; }
    for (int i = 0; i < d; ++i) taps[k][i] = (float)(one[i] / sum1);
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) {
      const double dist = sqrt((double)((i - r) * (i - r) + (j - r) * (j - r)));
      sum2 += (k2d[(size_t)k * PP_OKS_TAPS * PP_OKS_TAPS + i * d + j] = exp(-(dist * dist) / (2 * s)));
    }
    for (int i = 0; i < d * d; ++i) k2d[(size_t)k * PP_OKS_TAPS * PP_OKS_TAPS + i] /= sum2;
  }

  /* ---- inputs: keypoints well inside the image, all visible */
  float kps[B][K][2], vis[B][K];
  srand(7);
  for (int b = 0; b < B; ++b) for (int k = 0; k < K; ++k) {
    kps[b][k][0] = 24.f + (float)(rand() % 14400) / 100.f;   /* 24 .. 168 */
    kps[b][k][1] = 24.f + (float)(rand() % 20800) / 100.f;   /* 24 .. 232 */
    vis[b][k] = 1.f;
  }
  const size_t n = (size_t)B * K, plane = (size_t)H * W;
  void *d_kps = to_device(kps, sizeof kps), *d_vis = to_device(vis, sizeof vis), *d_two_s = to_device(two_s, sizeof two_s);
  void *d_hm = to_device(NULL, n * plane * 4), *d_grad = to_device(NULL, n * plane * 4), *d_w = to_device(NULL, n * 4);
  void *d_radius = to_device(radius, sizeof radius), *d_taps = to_device(taps, sizeof taps), *d_order = to_device(order, sizeof order);
  void *d_k2d = to_device(k2d, (size_t)K * PP_OKS_TAPS * PP_OKS_TAPS * 8), *d_idx = to_device(mma_index, sizeof mma_index);
  void *d_locs = to_device(NULL, n * 8), *d_vals = to_device(NULL, n * 4), *d_arg = to_device(NULL, n * 4), *d_kp64 = to_device(NULL, n * 16);
  if (!d_kps || !d_hm || !d_grad || !d_k2d || !d_kp64) { printf("cudaMalloc failed\n"); return 12; }

  pp_encode_params ep = {B, K, H, W, PP_F32, PP_F32, 2, (IW - 1) / (float)(W - 1), (IH - 1) / (float)(H - 1), (float)IW, (float)IH};
  PP(pp_encode(&ep, d_kps, d_vis, d_two_s, d_hm, d_w, NULL, NULL, NULL));

  /* ---- decode with the tensor-core kernel: operand tables + the larger scratch */
  const int64_t tbytes = pp_oks_mma_table_bytes(K, H, W);
  if (tbytes <= 0) { printf("no tensor-core tables for %dx%d\n", H, W); return 13; }
  void* d_tables = to_device(NULL, (size_t)tbytes);
  PP(pp_oks_mma_table_build(d_taps, d_radius, K, H, W, d_tables, NULL));
  pp_oks_table tab = {d_radius, d_taps, d_k2d, d_order, d_tables, d_idx, H, W};
  pp_decode_params dp = {B, K, H, W, PP_F32, 0, 1.0f, (double)IW, (double)IH};
  const int64_t sbytes = pp_decode_expected_scratch_bytes_for(&dp);
  void* d_scratch = to_device(NULL, (size_t)sbytes);
  PP(pp_decode_expected(&dp, &tab, d_hm, d_locs, d_vals, d_arg, d_kp64, NULL, d_scratch, sbytes, NULL));
  if (pp_decode_expected_last_kernel() != PP_DECODE_KERNEL_MMA) { printf("unexpected decode kernel %d\n", pp_decode_expected_last_kernel()); return 14; }

  /* ---- loss of the encoded maps against targets encoded inside the loss kernel: prediction == target */
  pp_loss_params lp;
  memset(&lp, 0, sizeof lp);
  lp.B = B; lp.K = K; lp.H = H; lp.W = W; lp.dtype = PP_F32; lp.mode = PP_LOSS_PIXEL_MEAN; lp.oks_type = 0;
  lp.smoothing_weight = 0.05; lp.gaussian_weight = 0.0; lp.loss_weight = 1.0;
  void *d_loss = to_device(NULL, 4), *d_lscr = to_device(NULL, (size_t)pp_oks_loss_scratch_bytes(&lp));
  PP(pp_oks_loss_forward_encoded(&lp, &ep, d_hm, d_kps, d_vis, d_two_s, NULL, d_loss, d_grad, 1.0f, NULL, NULL, NULL, d_lscr,
                                 pp_oks_loss_scratch_bytes(&lp), NULL, NULL));
  CK(cudaDeviceSynchronize());

  double kp64[B][K][2];
  float vals[B][K], w[B][K], loss = -1.f, *grad = malloc(n * plane * 4);
  int32_t arg[B][K];
  CK(cudaMemcpy(kp64, d_kp64, sizeof kp64, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vals, d_vals, sizeof vals, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(arg, d_arg, sizeof arg, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(w, d_w, sizeof w, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&loss, d_loss, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(grad, d_grad, n * plane * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int b = 0; b < B; ++b) for (int k = 0; k < K; ++k) {
    /* decode scales back by in / (hm - 1) while encode used (in - 1) / (hm - 1) (codec.py:131-133 vs 237): compare in heatmap px */
    const double hx = kp64[b][k][0] / IW * (W - 1), hy = kp64[b][k][1] / IH * (H - 1);
    const double gx = kps[b][k][0] / ep.scale_x, gy = kps[b][k][1] / ep.scale_y;
    const double err = fmax(fabs(hx - gx), fabs(hy - gy));
    if (err > worst) worst = err;
    const int ax = arg[b][k] % W, ay = arg[b][k] / W;
    if (abs(ax - (int)lrint(gx)) > 1 || abs(ay - (int)lrint(gy)) > 1 || vals[b][k] < 0.5f || vals[b][k] > 1.0f || w[b][k] != 1.0f) {
      printf("heatmap (%d, %d): argmax (%d, %d) for keypoint (%.2f, %.2f), score %g, weight %g\n", b, k, ax, ay, gx, gy, vals[b][k], w[b][k]);
      return 15;
    }
  }
  double gsum = 0;
  for (size_t i = 0; i < n * plane; ++i) { if (!isfinite(grad[i])) { printf("non-finite gradient\n"); return 16; } gsum += fabs(grad[i]); }
  printf("round trip: worst |decoded - encoded keypoint| = %.3f heatmap px over %zu heatmaps; loss %.6g; sum |grad| %.6g\n", worst, n, loss, gsum);
  if (!(worst < 0.6) || !(loss > 0.f) || !(gsum > 0)) return 17;
  printf("ok\n");
  return 0;
}

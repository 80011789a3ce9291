// Target encode: OKS-shaped probability maps (reference: generate_probmaps, codec.py:11-70, and the
// flag computation of ProbMap.encode / ArgMaxProbMap.encode, codec.py:187-200 / 489-502).
//
// Write-only and HBM-bound: H*W*sizeof(T) bytes per heatmap.  The map is separable,
//   exp(-((x-kx)^2 + (y-ky)^2) / (2s)) = ex[x] * ey[y],
// so one CTA computes the W + H factors of a heatmap in float64 (the reference evaluates the map in
// float64 and stores float32; the float64 product rounded to float32 reproduces it) and
// then streams the plane out with 128-bit stores, one fp64 multiply + one conversion per pixel.
#include <algorithm>

#include "pp_common.cuh"

namespace {

using namespace pp;

constexpr int kEncThreads = 128;

template <typename KpT>
__device__ __forceinline__ double to_heatmap_space(KpT v, float scale);
// NumPy: float32 keypoints / float32 scale_factor stay float32; float64 / float32 promote to float64.
template <>
__device__ __forceinline__ double to_heatmap_space<float>(float v, float scale) {
  return static_cast<double>(__fdiv_rn(v, scale));
}
template <>
__device__ __forceinline__ double to_heatmap_space<double>(double v, float scale) {
  return __ddiv_rn(v, static_cast<double>(scale));
}

constexpr int kEncSlots = 4;   // heatmaps a CTA encodes per barrier phase (at most)

struct EncodeSlot {
  int labelled;
};

template <typename T, typename KpT, bool kVector>
__global__ void __launch_bounds__(kEncThreads)
encode_kernel(pp_encode_params p, const KpT* __restrict__ keypoints, const float* __restrict__ visible,
              const double* __restrict__ two_s, T* __restrict__ heatmaps, float* __restrict__ weights,
              uint8_t* __restrict__ in_image, uint8_t* __restrict__ annotated, int slots) {
  extern __shared__ double factors[];  // per slot: ex[W] then ey[H]
  __shared__ EncodeSlot slot_info[kEncSlots];
  const int W = p.W, H = p.H, HW = H * W, F = W + H;
  const int N = p.B * p.K;
  constexpr int V = Elem<T>::kVec;
  const int tid = threadIdx.x;
  // phase-1 role: factor `fi` of slot `fs` (threads beyond slots * F have none)
  const int fs = tid / F, fi = tid - fs * F;

  for (int base = blockIdx.x * slots; base < N; base += gridDim.x * slots) {
    // ---- phase 1: the W + H separable factors of up to `slots` heatmaps, in float64
    // F <= 256: thread t owns factor t % F of slot t / F; larger maps: one slot, threads stride the factors
    const bool small = F <= kEncThreads;
    const int s = small ? fs : 0, i = small ? fi : tid;
    if ((small ? fs < slots : true) && base + s < N) {
      const int hm = base + s;
      const int k = hm % p.K;
      const float vis = visible ? visible[hm] : 1.0f;
      const bool labelled = !(vis < 0.5f);  // codec.py:53
      const KpT kx_in = keypoints[static_cast<size_t>(hm) * p.keypoint_dim + 0];
      const KpT ky_in = keypoints[static_cast<size_t>(hm) * p.keypoint_dim + 1];
      const double kx = to_heatmap_space<KpT>(kx_in, p.scale_x);
      const double ky = to_heatmap_space<KpT>(ky_in, p.scale_y);
      const double div = two_s[k];
      if (labelled) {
        for (int j = i; j < F; j += kEncThreads) {   // F > kEncThreads only for very large maps
          const double d = (j < W) ? (static_cast<double>(j) - kx) : (static_cast<double>(j - W) - ky);
          factors[s * F + j] = exp(-(d * d / div));
        }
      }
      if (i == 0) {
        slot_info[s].labelled = labelled;
        if (annotated) annotated[hm] = vis > 0.0f;
        if (in_image) {
          // comparisons in the keypoint dtype against the (integer) input size, codec.py:189-200
          const KpT w = static_cast<KpT>(p.input_w), h = static_cast<KpT>(p.input_h);
          in_image[hm] = (kx_in >= KpT(0)) && (kx_in < w) && (ky_in >= KpT(0)) && (ky_in < h);
        }
        if (weights) {
          float wgt = vis;  // unlabelled keypoints keep their visibility value (codec.py:46,53-54)
          if (labelled) {
            // weight = (max over the grid of the float64 map) > 0 (codec.py:68): the maximum sits at
            // the grid point nearest to the keypoint.
            const double xn = fmin(fmax(rint(kx), 0.0), static_cast<double>(W - 1));
            const double yn = fmin(fmax(rint(ky), 0.0), static_cast<double>(H - 1));
            const double dx = xn - kx, dy = yn - ky;
            const double dist = sqrt(dx * dx + dy * dy);
            wgt = exp(-(dist * dist / div)) > 0.0 ? 1.0f : 0.0f;
          }
          weights[hm] = wgt;
        }
      }
    }
    __syncthreads();

    // ---- phase 2: stream the planes out
    for (int s = 0; s < slots && base + s < N; ++s) {
      T* plane = heatmaps + static_cast<size_t>(base + s) * HW;
      const double* ex = factors + s * F;
      const double* ey = ex + W;
      const bool labelled = slot_info[s].labelled != 0;
      if (kVector) {
        // a thread keeps one column vector: its V x-factors stay in registers, per row it needs one
        // shared-memory read (ey), V fp64 multiplies + conversions and one 128-bit store
        const int WV = W / V;
        const int rows_per_pass = kEncThreads / WV;   // threads beyond WV * rows_per_pass idle here
        const int ty = tid / WV, xv = tid - ty * WV;
        if (ty < rows_per_pass) {
          if (labelled) {
            double fx[V];
#pragma unroll
            for (int j = 0; j < V; ++j) fx[j] = ex[xv * V + j];
            for (int y = ty; y < H; y += rows_per_pass) {
              const double fy = ey[y];
              float px[V];
#pragma unroll
              for (int j = 0; j < V; ++j) px[j] = static_cast<float>(fx[j] * fy);
              stg_stream_128(plane + (y * WV + xv) * V, pack(px, T()));
            }
          } else {  // unlabelled channel stays zero (codec.py:45)
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int y = ty; y < H; y += rows_per_pass) stg_stream_128(plane + (y * WV + xv) * V, z);
          }
        }
      } else {
        for (int i = tid; i < HW; i += kEncThreads) {
          const int y = i / W, x = i - y * W;
          plane[i] = Elem<T>::from_f32(labelled ? static_cast<float>(ex[x] * ey[y]) : 0.0f);
        }
      }
    }
    __syncthreads();  // factors are rewritten by the next group
  }
}

template <typename T, typename KpT>
int launch_encode(const pp_encode_params& p, const void* keypoints, const float* visible, const double* two_s,
                  void* heatmaps, float* weights, uint8_t* in_image, uint8_t* annotated, cudaStream_t st) {
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  PP_REQUIRE(N < (1ll << 31) && N * p.keypoint_dim < (1ll << 40), PP_ERR_UNSUPPORTED_SHAPE, "pp_encode: too many heatmaps");
  const bool vec = (p.W % Elem<T>::kVec == 0) && pp_aligned16(heatmaps) && p.W / Elem<T>::kVec <= kEncThreads;
  const int F = p.W + p.H;
  const int slots = std::max(1, std::min(kEncSlots, kEncThreads / F));
  const size_t smem = sizeof(double) * F * slots;
  PP_REQUIRE(smem <= 48 * 1024, PP_ERR_UNSUPPORTED_SHAPE, "pp_encode: W + H = %d too large", F);
  const int64_t groups = (N + slots - 1) / slots;
  const int grid = static_cast<int>(std::min<int64_t>(groups, static_cast<int64_t>(pp_sm_count()) * 16));
  auto kp = static_cast<const KpT*>(keypoints);
  auto hm = static_cast<T*>(heatmaps);
  if (vec)
    encode_kernel<T, KpT, true><<<grid, kEncThreads, smem, st>>>(p, kp, visible, two_s, hm, weights, in_image, annotated, slots);
  else
    encode_kernel<T, KpT, false><<<grid, kEncThreads, smem, st>>>(p, kp, visible, two_s, hm, weights, in_image, annotated, slots);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // namespace

extern "C" int pp_encode(const pp_encode_params* p, const void* keypoints, const float* visible, const double* two_s,
                         void* heatmaps, float* keypoint_weights, uint8_t* in_image, uint8_t* annotated,
                         pp_stream_t stream) {
  PP_REQUIRE(p != nullptr, PP_ERR_INVALID_ARG, "pp_encode: null params");
  PP_REQUIRE(p->B >= 0 && p->K > 0 && p->H > 0 && p->W > 0 && p->keypoint_dim >= 2, PP_ERR_INVALID_ARG,
             "pp_encode: bad shape B=%d K=%d H=%d W=%d D=%d", p->B, p->K, p->H, p->W, p->keypoint_dim);
  PP_REQUIRE(p->scale_x != 0.0f && p->scale_y != 0.0f, PP_ERR_INVALID_ARG, "pp_encode: zero scale factor");
  if (p->B == 0) return PP_OK;
  PP_REQUIRE(keypoints && two_s && heatmaps, PP_ERR_INVALID_ARG, "pp_encode: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int hd = p->heatmap_dtype, kd = p->keypoint_dtype;
  if (hd == PP_F32 && kd == PP_F32)
    return launch_encode<float, float>(*p, keypoints, visible, two_s, heatmaps, keypoint_weights, in_image, annotated, st);
  if (hd == PP_F32 && kd == PP_F64)
    return launch_encode<float, double>(*p, keypoints, visible, two_s, heatmaps, keypoint_weights, in_image, annotated, st);
  if (hd == PP_BF16 && kd == PP_F32)
    return launch_encode<__nv_bfloat16, float>(*p, keypoints, visible, two_s, heatmaps, keypoint_weights, in_image, annotated, st);
  if (hd == PP_BF16 && kd == PP_F64)
    return launch_encode<__nv_bfloat16, double>(*p, keypoints, visible, two_s, heatmaps, keypoint_weights, in_image, annotated, st);
  pp_set_error("pp_encode: unsupported dtypes heatmap=%d keypoint=%d", hd, kd);
  return PP_ERR_INVALID_ARG;
}

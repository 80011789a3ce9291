// OKS heatmap loss (reference: OKSHeatmapLoss.forward/_get_mask, loss.py:55-191) forward and backward.
//
//   per pixel:  e   = gx^2 + gy^2            gx, gy = 3x3 Sobel cross-correlations of `output`, zero 'same' pad
//               oks = out (1 - tgt) | (1 - out) tgt | their mean        (loss.py:92-99)
//               mse = (out - tgt)^2                                     (loss.py:103)
//               l   = (w_s e m + w_o oks m + w_g mse m) * loss_weight   (loss.py:112-127, 143)
//   d l / d out = U m (w_o d oks + 2 w_g (out - tgt)) lw  -  Sx * (2 c gx)  -  Sy * (2 c gy),   c = lw w_s U m
//   (the adjoint of a zero-padded cross-correlation with Sx is a cross-correlation with flip(Sx) = -Sx).
//
// General kernel (this file): one CTA owns one heatmap; `output` is staged once into a zero-bordered
// shared plane, P = 2 c gx and Q = 2 c gy go to two more planes, and the gradient is the 3x3 adjoint
// stencil over them.  Supports every mode / weight / mask combination of the reference.
#include <algorithm>
#include <cstdlib>

#include "pp_common.cuh"
#include "pp_loss_fast.cuh"
#include "pp_mailbox.cuh"
#ifdef PP_EXPERIMENTS   // measured-and-rejected variant, outside the default build
#include "../../tools/experiments/pp_loss_pair.cuh"
#endif

namespace {

using namespace pp;

constexpr int kLossThreads = 256;

// upstream gradient: a scalar (host value and/or one device float) broadcast over the forward's output, or a full tensor
enum Upstream : int { kUpScalar = 0, kUpPerPixel = 2, kUpPerKeypoint = 3 };

struct LossArgs {
  pp_loss_params p;
  const void* output;
  const void* target;
  const float* kp_weights;
  const void* pix_weights;
  const void* mask;
  void* loss_map;
  float* loss_kpt;
  int32_t* peak_out;
  double* partials;       // (N) per-heatmap sum of the per-pixel loss
  void* grad;
  int upstream_kind;
  float host_scale;
  const void* upstream;
  const int32_t* peak_in;
  int32_t* range_flag;
};

__device__ __forceinline__ float oks_term(int type, float o, float t) {
  const float minus = __fmul_rn(o, __fsub_rn(1.0f, t));
  if (type == 0) return minus;
  const float plus = __fmul_rn(__fsub_rn(1.0f, o), t);
  if (type == 1) return plus;
  return __fdiv_rn(__fadd_rn(minus, plus), 2.0f);
}
__device__ __forceinline__ float oks_term_grad(int type, float t) {
  return type == 0 ? (1.0f - t) : type == 1 ? -t : 0.5f * (1.0f - 2.0f * t);
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads)
oks_loss_kernel(LossArgs a, bool want_fwd, bool want_grad, int band_h) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_f[8];
  __shared__ double red_d[3][8];
  __shared__ int red_i[8];

  const pp_loss_params& p = a.p;
  const int H = p.H, W = p.W, HW = H * W, S = W + 2;
  // The heatmap is processed in horizontal bands of band_h rows (one band when it fits shared memory):
  //   A : rows y0-2 .. y1+1 of `output` (zero outside the map), (band_h + 4) x S
  //   P, Q : 2 c gx, 2 c gy on rows y0-1 .. y1 (zero outside the map),  (band_h + 2) x S each
  // Column 0 and column W+1 of every plane stay zero (the 'same' zero padding of the Sobel stencils).
  float* A = smem;
  float* P = A + (band_h + 4) * S;
  float* Q = P + (band_h + 2) * S;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const T* out = static_cast<const T*>(a.output);
  const T* tgt = static_cast<const T*>(a.target);
  const T* pixw = static_cast<const T*>(a.pix_weights);
  const T* mask = static_cast<const T*>(a.mask);
  // loss.py:118-120: Python-float arithmetic, rounded to float32 when it meets the tensor
  const float w_s = static_cast<float>(p.smoothing_weight), w_g = static_cast<float>(p.gaussian_weight);
  const float w_o = static_cast<float>(1.0 - p.smoothing_weight - p.gaussian_weight);
  const float lw = static_cast<float>(p.loss_weight);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const int total_floats = (band_h + 4) * S + (want_grad ? 2 * (band_h + 2) * S : 0);
  for (int i = threadIdx.x; i < total_floats; i += kLossThreads) smem[i] = 0.0f;
  __syncthreads();

  int bad_target = 0;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const int b = static_cast<int>(hm / p.K), k = static_cast<int>(hm % p.K);
    const T* o_ptr = out + hm * HW;
    const T* t_ptr = tgt + hm * HW;
    const T* m_ptr = mask ? mask + b * p.mask_stride_b + k * p.mask_stride_k : nullptr;
    const T* pw_ptr = pixw ? pixw + hm * HW : nullptr;
    const T* up_ptr = (want_grad && a.upstream_kind == kUpPerPixel) ? static_cast<const T*>(a.upstream) + hm * HW : nullptr;
    T* g_ptr = want_grad ? static_cast<T*>(a.grad) + hm * HW : nullptr;

    float m_k = a.kp_weights ? a.kp_weights[hm] : 1.0f;
    if (p.skip_empty_channel) {  // (target != 0).any() per channel, loss.py:180-183
      int nz = 0;
      for (int i = threadIdx.x; i < HW; i += kLossThreads) nz |= (Elem<T>::to_f32(t_ptr[i]) != 0.0f);
      nz = __syncthreads_or(nz);
      if (!nz) m_k = 0.0f;
    }

    // upstream coefficient shared by the whole heatmap
    float u_k = 1.0f;
    int peak = -1;
    if (want_grad) {
      if (a.upstream_kind == kUpScalar) {
        u_k = a.host_scale;
        if (a.upstream) u_k *= static_cast<const float*>(a.upstream)[0];
        if (p.mode == PP_LOSS_PIXEL_MEAN) u_k /= static_cast<float>(N * HW);
      } else if (a.upstream_kind == kUpPerKeypoint) {
        u_k = static_cast<const float*>(a.upstream)[hm];
      }
      if (p.mode == PP_LOSS_PER_KEYPOINT) peak = a.peak_in[hm];
    }
    const float wg_grad = (p.mode == PP_LOSS_PER_KEYPOINT) ? w_g / static_cast<float>(HW) : w_g;  // mean over pixels

    double sum_l = 0.0, sum_oks = 0.0, sum_mse = 0.0;
    float max_e = -INFINITY;
    int max_i = 0x7fffffff;

    for (int y0 = 0; y0 < H; y0 += band_h) {
      const int y1 = min(y0 + band_h, H), bh = y1 - y0;
      // stage rows y0-2 .. y1+1
      for (int i = threadIdx.x; i < (bh + 4) * W; i += kLossThreads) {
        const int lr = i / W, x = i - lr * W, y = y0 - 2 + lr;
        A[lr * S + x + 1] = (y >= 0 && y < H) ? Elem<T>::to_f32(o_ptr[y * W + x]) : 0.0f;
      }
      __syncthreads();

      // phase 1: Sobel pair on rows y0-1 .. y1; forward terms on the band's own rows
      for (int i = threadIdx.x; i < (bh + 2) * W; i += kLossThreads) {
        const int pr = i / W, x = i - pr * W, y = y0 - 1 + pr;
        const bool inside = y >= 0 && y < H;
        const float* c = A + (pr + 1) * S + x + 1;
        // Sobel cross-correlations (loss.py:106-109)
        const float gx = (c[-S - 1] - c[-S + 1]) + 2.0f * (c[-1] - c[1]) + (c[S - 1] - c[S + 1]);
        const float gy = (c[-S - 1] + 2.0f * c[-S] + c[-S + 1]) - (c[S - 1] + 2.0f * c[S] + c[S + 1]);
        const int pix = y * W + x;
        float m = m_k;
        if (inside) {
          if (pw_ptr) m *= Elem<T>::to_f32(pw_ptr[pix]);
          if (m_ptr) m *= Elem<T>::to_f32(m_ptr[pix]);
        }
        if (want_fwd && y >= y0 && y < y1) {
          const float o = c[0];
          const float t = Elem<T>::to_f32(t_ptr[pix]);
          bad_target |= !(t >= 0.0f && t <= 1.0f);
          const float e = __fmul_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), m);
          const float ok = __fmul_rn(oks_term(p.oks_type, o, t), m);
          const float d = __fsub_rn(o, t);
          const float ms = __fmul_rn(__fmul_rn(d, d), m);
          if (p.mode == PP_LOSS_PER_KEYPOINT) {
            sum_oks += ok;
            sum_mse += ms;
            if (e > max_e) { max_e = e; max_i = pix; }
          } else {
            const float l = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(w_s, e), __fmul_rn(w_o, ok)), __fmul_rn(w_g, ms)), lw);
            if (p.mode == PP_LOSS_PER_PIXEL) static_cast<T*>(a.loss_map)[hm * HW + pix] = Elem<T>::from_f32(l);
            else sum_l += l;
          }
        }
        if (want_grad) {
          float cs = 0.0f;
          if (inside) {
            const float u = up_ptr ? Elem<T>::to_f32(up_ptr[pix]) : u_k;
            cs = lw * w_s * u * m;
            if (p.mode == PP_LOSS_PER_KEYPOINT && pix != peak) cs = 0.0f;  // max() routes to one pixel
          }
          P[pr * S + x + 1] = 2.0f * cs * gx;
          Q[pr * S + x + 1] = 2.0f * cs * gy;
        }
      }

      if (want_grad) {
        __syncthreads();
        // phase 2: adjoint stencil on the band's own rows
        for (int i = threadIdx.x; i < bh * W; i += kLossThreads) {
          const int lr = i / W, x = i - lr * W, y = y0 + lr;
          const int at = (lr + 1) * S + x + 1;   // row y in P / Q coordinates
          const float* pc = P + at;
          const float* qc = Q + at;
          const float sx = (pc[-S - 1] - pc[-S + 1]) + 2.0f * (pc[-1] - pc[1]) + (pc[S - 1] - pc[S + 1]);
          const float sy = (qc[-S - 1] + 2.0f * qc[-S] + qc[-S + 1]) - (qc[S - 1] + 2.0f * qc[S] + qc[S + 1]);
          const int pix = y * W + x;
          const float o = A[(lr + 2) * S + x + 1];
          const float t = Elem<T>::to_f32(t_ptr[pix]);
          float m = m_k;
          if (pw_ptr) m *= Elem<T>::to_f32(pw_ptr[pix]);
          if (m_ptr) m *= Elem<T>::to_f32(m_ptr[pix]);
          const float u = up_ptr ? Elem<T>::to_f32(up_ptr[pix]) : u_k;
          const float direct = lw * u * m * (w_o * oks_term_grad(p.oks_type, t) + wg_grad * 2.0f * (o - t));
          g_ptr[pix] = Elem<T>::from_f32(direct - sx - sy);
        }
      }
      __syncthreads();   // the planes are rewritten by the next band / heatmap
    }

    if (want_fwd && p.mode != PP_LOSS_PER_PIXEL) {
      // block reductions (fixed order -> deterministic)
      sum_l = warp_sum(sum_l); sum_oks = warp_sum(sum_oks); sum_mse = warp_sum(sum_mse);
      warp_argmax(max_e, max_i);
      if (lane == 0) { red_d[0][warp] = sum_l; red_d[1][warp] = sum_oks; red_d[2][warp] = sum_mse; red_f[warp] = max_e; red_i[warp] = max_i; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double sl = 0, so = 0, sm = 0;
        float me = red_f[0]; int mi = red_i[0];
        for (int w = 0; w < kLossThreads / 32; ++w) { sl += red_d[0][w]; so += red_d[1][w]; sm += red_d[2][w]; }
        for (int w = 1; w < kLossThreads / 32; ++w) argmax_combine(me, mi, red_f[w], red_i[w]);
        if (p.mode == PP_LOSS_PER_KEYPOINT) {
          // w_o * sum(oks) + w_s * max(e) + w_g * mean(mse), then * loss_weight (loss.py:128-134, 143)
          const float v = (w_o * static_cast<float>(so) + w_s * me + w_g * static_cast<float>(sm / HW)) * lw;
          a.loss_kpt[hm] = v;
          if (a.peak_out) a.peak_out[hm] = mi;
          a.partials[hm] = static_cast<double>(v);
        } else {
          a.partials[hm] = sl;
        }
      }
      __syncthreads();
    }
  }
  if (a.range_flag && bad_target) atomicOr(a.range_flag, 1);
}

// sum of N doubles * scale -> one float; single CTA, fixed order.  With a mailbox (has_mb) the first warp also stores
// the loss into the step's slot on every rank and arrives at the slot's publication (pp_records.cu): the multi-GPU
// exchange needs no launch after the loss.
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ partials, int64_t n, double scale,
                                                       float* __restrict__ out, pp_mailbox mb, int64_t mb_records, int has_mb) {
  __shared__ double red[8];
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched programmatically dependent on the kernel that writes `partials`
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += partials[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {                        // every lane forms the same sum, in the same fixed order
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float loss = static_cast<float>(t * scale);
    if (threadIdx.x == 0) out[0] = loss;
    if (has_mb) pp_mailbox_dev::mailbox_store_loss_and_arrive(mb, mb_records, static_cast<double>(loss));   // warp-wide
  }
}

// finalize_kernel right behind the kernel that produces the partial sums, as a programmatic dependent launch: the block
// is resident before the producer has finished (oks_loss_fast_kernel releases its dependents at its start) and waits in
// griddepcontrol.wait; ~2-3 us less at the tail of every step.  PP_LOSS_PDL=0: an ordinary launch.
cudaError_t launch_finalize(const double* partials, int64_t n, double scale, float* out, const pp_mailbox& mb, int64_t mb_records,
                            int has_mb, bool dependent, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const bool pdl = dependent && pp_env_int("PP_LOSS_PDL", 1) != 0;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, finalize_kernel, partials, n, scale, out, mb, mb_records, has_mb);
}

int check_publish(const char* fn, const pp_mailbox* mb, int64_t n_records) {
  if (!mb) return PP_OK;
  PP_REQUIRE(mb->peer_bufs && mb->state && mb->world >= 1 && mb->rank >= 0 && mb->rank < mb->world && mb->slots >= 1 &&
                 mb->slot >= 0 && mb->slot < mb->slots && mb->block_bytes == pp_mailbox_block_bytes(n_records),
             PP_ERR_INVALID_ARG, "%s: inconsistent mailbox", fn);
  return PP_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* __restrict__ data, int64_t numel, const float* __restrict__ scale) {
  const float s = scale[0];
  if (s == 1.0f) return;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < numel; i += stride)
    data[i] = Elem<T>::from_f32(Elem<T>::to_f32(data[i]) * s);
}

int check_loss_params(const char* fn, const pp_loss_params* p) {
  PP_REQUIRE(p != nullptr, PP_ERR_INVALID_ARG, "%s: null params", fn);
  PP_REQUIRE(p->B >= 0 && p->K > 0 && p->H > 0 && p->W > 0, PP_ERR_INVALID_ARG, "%s: bad shape B=%d K=%d H=%d W=%d", fn,
             p->B, p->K, p->H, p->W);
  PP_REQUIRE(p->dtype == PP_F32 || p->dtype == PP_BF16, PP_ERR_INVALID_ARG, "%s: unsupported dtype %d", fn, p->dtype);
  PP_REQUIRE(p->mode >= 0 && p->mode <= 2, PP_ERR_INVALID_ARG, "%s: bad mode %d", fn, p->mode);
  PP_REQUIRE(p->oks_type >= 0 && p->oks_type <= 2, PP_ERR_INVALID_ARG, "%s: bad oks_type %d", fn, p->oks_type);
  return PP_OK;
}

int env_int(const char* name, int fallback) { return pp_env_int(name, fallback); }

template <typename T>
int launch_loss(const LossArgs& a, bool fwd, bool grad, cudaStream_t st) {
  const pp_loss_params& p = a.p;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  // bands of band_h rows: (band_h + 4) rows of `output` and, with a gradient, 2 x (band_h + 2) rows of P / Q
  const size_t row = sizeof(float) * static_cast<size_t>(p.W + 2);
  const size_t budget = static_cast<size_t>(pp_smem_optin()) - 4096;
  const size_t fixed = row * (grad ? 8 : 4);
  PP_REQUIRE(fixed + row * (grad ? 3 : 1) <= budget, PP_ERR_UNSUPPORTED_SHAPE,
             "pp_oks_loss: rows of %d pixels are too wide for shared memory", p.W);
  int band_h = static_cast<int>(std::min<size_t>(p.H, (budget - fixed) / (row * (grad ? 3 : 1))));
  if (const int cap = env_int("PP_LOSS_BAND", 0); cap > 0) band_h = std::min(band_h, cap);   // test hook: force several bands
  const size_t smem = fixed + row * (grad ? 3 : 1) * band_h;
  int per_sm = 1;
  if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(oks_loss_kernel<T>), kLossThreads, smem, &per_sm)) return rc;
  const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * per_sm));
  oks_loss_kernel<T><<<grid, kLossThreads, smem, st>>>(a, fwd, grad, band_h);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

// ---- fast path dispatch (pp_loss_fast.cuh) ---------------------------------------------------
bool fast_path_ok(const pp_loss_params& p, const void* output, const void* target, const void* pixel_weights,
                  const void* mask, const void* grad) {
  const int e = p.dtype == PP_F32 ? 4 : 2;
  return p.mode == PP_LOSS_PIXEL_MEAN && !pixel_weights && !mask && !p.skip_empty_channel && p.W % 4 == 0 &&
         p.W / 4 <= 128 && (static_cast<int64_t>(p.H) * p.W * e) % 16 == 0 && pp_aligned16(output) &&
         pp_aligned16(target) && pp_aligned16(grad) &&
         static_cast<int64_t>(p.H) * p.W * e + 512 <= pp_smem_optin();
}

template <typename T, bool kFwd, bool kGrad, int kTgt>
int launch_fast_tt(const pp_loss_fast::FastArgs& a, int threads, size_t smem, cudaStream_t st, int* grid_out) {
  int per_sm = 1;
  auto kern = pp_loss_fast::oks_loss_fast_kernel<T, kFwd, kGrad, kTgt>;
  if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(kern), threads, smem, &per_sm)) return rc;
  if (const int cap = env_int("PP_LOSS_CTAS", 0); cap > 0) per_sm = std::min(per_sm, cap);
  const int64_t units = (a.N + a.G - 1) / a.G;
  int grid = static_cast<int>(std::min<int64_t>(units, static_cast<int64_t>(pp_sm_count()) * per_sm));
  if (const int cap = env_int("PP_LOSS_GRID", 0); cap > 0) grid = std::min(grid, cap);   // test hook: many units per CTA
  kern<<<grid, threads, smem, st>>>(a);
  PP_CUDA_OK(cudaGetLastError());
  *grid_out = grid;
  return PP_OK;
}

#ifdef PP_EXPERIMENTS
template <bool kFwd, bool kGrad>
int launch_pair_t(const pp_loss_fast::FastArgs& a, int threads, size_t smem, cudaStream_t st, int* grid_out) {
  int per_sm = 1;
  auto kern = pp_loss_pair::oks_loss_pair_kernel<kFwd, kGrad>;
  if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(kern), threads, smem, &per_sm)) return rc;
  if (const int cap = env_int("PP_LOSS_CTAS", 0); cap > 0) per_sm = std::min(per_sm, cap);
  const int64_t units = (a.N + a.G - 1) / a.G;
  int grid = static_cast<int>(std::min<int64_t>(units, static_cast<int64_t>(pp_sm_count()) * per_sm));
  if (const int cap = env_int("PP_LOSS_GRID", 0); cap > 0) grid = std::min(grid, cap);   // test hook: many units per CTA
  kern<<<grid, threads, smem, st>>>(a);
  PP_CUDA_OK(cudaGetLastError());
  *grid_out = grid;
  return PP_OK;
}
#endif

template <typename T, bool kFwd, bool kGrad>
int launch_fast_t(const pp_loss_fast::FastArgs& a, int threads, size_t smem, cudaStream_t st, int* grid_out) {
  if (a.keypoints) return launch_fast_tt<T, kFwd, kGrad, pp_loss_fast::kTgtEncode>(a, threads, smem, st, grid_out);
  if (a.tgt_off != 0) return launch_fast_tt<T, kFwd, kGrad, pp_loss_fast::kTgtSmem>(a, threads, smem, st, grid_out);
  return launch_fast_tt<T, kFwd, kGrad, pp_loss_fast::kTgtGlobal>(a, threads, smem, st, grid_out);
}

// the encoder's inputs for the encode-inside-loss variant (target == nullptr)
struct EncodeSource {
  const pp_encode_params* ep = nullptr;
  const void* keypoints = nullptr;
  const float* visible = nullptr;
  const double* two_s = nullptr;
  float* weights_out = nullptr;
  uint8_t* in_image = nullptr;
  uint8_t* annotated = nullptr;
};

// Runs the fused mean-mode kernel; *grid_out = number of per-CTA partial sums written (forward).
int launch_fast(const pp_loss_params& p, const void* output, const void* target, const float* kp_weights, void* grad,
                double* partials, int32_t* range_flag, const float* upstream, float host_scale, bool fwd,
                cudaStream_t st, int* grid_out, const EncodeSource* enc = nullptr) {
  pp_loss_fast::FastArgs a{};
  if (enc) {
    a.keypoints = enc->keypoints; a.visible = enc->visible; a.two_s = enc->two_s;
    a.weights_out = enc->weights_out; a.in_image = enc->in_image; a.annotated = enc->annotated;
    a.K = p.K; a.kp_dim = enc->ep->keypoint_dim; a.kp_f64 = enc->ep->keypoint_dtype == PP_F64;
    a.scale_x = enc->ep->scale_x; a.scale_y = enc->ep->scale_y; a.input_w = enc->ep->input_w; a.input_h = enc->ep->input_h;
  }
  const int e = p.dtype == PP_F32 ? 4 : 2;
  a.output = output; a.target = target; a.kp_weights = kp_weights; a.grad = grad; a.partials = partials;
  a.range_flag = range_flag; a.upstream = upstream; a.host_scale = host_scale;
  a.N = static_cast<long long>(p.B) * p.K;
  a.H = p.H; a.W = p.W;
  a.strips = p.W / 4;
  // a thread owns a strip of 4 columns x T rows (plus a 2-row halo on each side that it recomputes):
  // T ~ 16 keeps the halo overhead at 25 %; G consecutive heatmaps share a CTA so that it has ~96 threads
  const int want_T = env_int("PP_LOSS_T", 8);
  int segs = std::max(1, (p.H + want_T / 2) / want_T);
  a.T = (p.H + segs - 1) / segs;
  a.segs = (p.H + a.T - 1) / a.T;
  while (a.strips * a.segs > 256) { a.T *= 2; a.segs = (p.H + a.T - 1) / a.T; }
  const int per = a.strips * a.segs;
  a.G = std::max(1, std::min(4, env_int("PP_LOSS_G", (96 + per / 2) / per)));
  while (a.G > 1 && per * a.G > 256) --a.G;
  if (static_cast<long long>(a.G) > static_cast<long long>(p.B) * p.K) a.G = 1;
  a.w_s = static_cast<float>(p.smoothing_weight);
  a.w_g = static_cast<float>(p.gaussian_weight);
  a.w_o = static_cast<float>(1.0 - p.smoothing_weight - p.gaussian_weight);
  a.lw = static_cast<float>(p.loss_weight);
  // oks = a_o o + a_t t - o t ;  d oks / d o = d_a + d_b t   (loss.py:92-99)
  if (p.oks_type == 0) { a.a_o = 1.f; a.a_t = 0.f; a.d_a = 1.f; a.d_b = -1.f; }
  else if (p.oks_type == 1) { a.a_o = 0.f; a.a_t = 1.f; a.d_a = 0.f; a.d_b = -1.f; }
  else { a.a_o = 0.5f; a.a_t = 0.5f; a.d_a = 0.5f; a.d_b = -1.f; }
  a.inv_count = static_cast<float>(1.0 / (static_cast<double>(a.N) * p.H * p.W));
  a.plane_bytes = static_cast<unsigned>(static_cast<int64_t>(p.H) * p.W * e);
  const bool g = grad != nullptr;

#ifdef PP_EXPERIMENTS
  // float32: two heatmaps per thread with packed FADD2 / FMUL2 / FFMA2 arithmetic (pp_loss_pair.cuh), whenever a
  // unit of 2 x pairs heatmaps (output + target) fits shared memory
  if (p.dtype == PP_F32 && a.N >= 2 && per <= 256 && env_int("PP_LOSS_PAIR", 0)) {
    pp_loss_fast::FastArgs b = a;
    int pairs = std::max(1, std::min(2, env_int("PP_LOSS_G", (96 + per / 2) / per)));
    while (pairs > 1 && per * pairs > 256) --pairs;
    for (;; --pairs) {
      b.G = 2 * pairs;
      b.tgt_off = (16 + b.G * b.plane_bytes + 16 + 127) / 128 * 128;
      b.stage_bytes = (b.tgt_off + b.G * b.plane_bytes + 127) / 128 * 128;
      if (b.stage_bytes <= static_cast<size_t>(pp_smem_optin()) || pairs == 1) break;
    }
    if (b.stage_bytes <= static_cast<size_t>(pp_smem_optin())) {
      b.stages = env_int("PP_LOSS_STAGES", 2);
      if (static_cast<size_t>(b.stages) * b.stage_bytes > static_cast<size_t>(pp_smem_optin())) b.stages = 1;
      const size_t smem = static_cast<size_t>(b.stages) * b.stage_bytes;
      const int threads = per * pairs;
      if (fwd && g) return launch_pair_t<true, true>(b, threads, smem, st, grid_out);
      if (fwd) return launch_pair_t<true, false>(b, threads, smem, st, grid_out);
      return launch_pair_t<false, true>(b, threads, smem, st, grid_out);
    }
  }
#endif

  const size_t fac_bytes = enc ? sizeof(float) * 2 * (p.W + p.H + 4) : 0;   // per heatmap slot of a unit
  if (enc) {
    // only the output planes are staged; the target is formed from the keypoint's factors
    for (;; --a.G) {
      a.tgt_off = 0;
      a.stage_bytes = (16 + a.G * a.plane_bytes + 16 + 127) / 128 * 128;
      if (a.stage_bytes + a.G * fac_bytes <= static_cast<size_t>(pp_smem_optin()) || a.G == 1) break;
    }
  } else {
    for (;; --a.G) {
      a.tgt_off = (16 + a.G * a.plane_bytes + 16 + 127) / 128 * 128;
      a.stage_bytes = (a.tgt_off + a.G * a.plane_bytes + 127) / 128 * 128;
      if (a.stage_bytes <= static_cast<size_t>(pp_smem_optin()) || a.G == 1) break;
    }
    if (a.stage_bytes > static_cast<size_t>(pp_smem_optin())) {
      // very large maps: only the output plane is staged, the target is read from global memory (tgt_off = 0)
      a.tgt_off = 0;
      a.stage_bytes = (16 + a.plane_bytes + 16 + 127) / 128 * 128;
    }
  }
  // two stages (the next unit lands while the current one is processed) whenever they fit
  a.stages = env_int("PP_LOSS_STAGES", 2);
  if (static_cast<size_t>(a.stages) * a.stage_bytes + a.G * fac_bytes > static_cast<size_t>(pp_smem_optin())) a.stages = 1;
  a.fac_off = static_cast<unsigned>(static_cast<size_t>(a.stages) * a.stage_bytes);
  const size_t smem = a.fac_off + a.G * fac_bytes;
  const int threads = a.strips * a.segs * a.G;
  if (p.dtype == PP_F32) {
    if (fwd && g) return launch_fast_t<float, true, true>(a, threads, smem, st, grid_out);
    if (fwd) return launch_fast_t<float, true, false>(a, threads, smem, st, grid_out);
    return launch_fast_t<float, false, true>(a, threads, smem, st, grid_out);
  }
  if (fwd && g) return launch_fast_t<__nv_bfloat16, true, true>(a, threads, smem, st, grid_out);
  if (fwd) return launch_fast_t<__nv_bfloat16, true, false>(a, threads, smem, st, grid_out);
  return launch_fast_t<__nv_bfloat16, false, true>(a, threads, smem, st, grid_out);
}

}  // namespace

extern "C" {

int64_t pp_oks_loss_scratch_bytes(const pp_loss_params* p) {
  if (!p) return 0;
  return static_cast<int64_t>(sizeof(double)) * std::max<int64_t>(1, static_cast<int64_t>(p->B) * p->K);
}

int pp_oks_loss_forward(const pp_loss_params* p, const void* output, const void* target, const float* keypoint_weights,
                        const void* pixel_weights, const void* mask, void* loss_map, float* loss_kpt,
                        float* loss_scalar, int32_t* peak_index, void* grad, float grad_scale,
                        int32_t* target_out_of_range, void* scratch, int64_t scratch_bytes, const pp_mailbox* publish,
                        pp_stream_t stream) {
  if (int rc = check_loss_params("pp_oks_loss_forward", p)) return rc;
  if (int rc = check_publish("pp_oks_loss_forward", publish, static_cast<int64_t>(p->B) * p->K)) return rc;
  PP_REQUIRE(!publish || p->mode == PP_LOSS_PIXEL_MEAN, PP_ERR_INVALID_ARG, "pp_oks_loss_forward: `publish` needs PP_LOSS_PIXEL_MEAN");
  const pp_mailbox mbv = publish ? *publish : pp_mailbox{};
  if (static_cast<int64_t>(p->B) * p->K == 0) return PP_OK;
  PP_REQUIRE(output && target, PP_ERR_INVALID_ARG, "pp_oks_loss_forward: null heatmaps");
  PP_REQUIRE(scratch && scratch_bytes >= pp_oks_loss_scratch_bytes(p), PP_ERR_SCRATCH,
             "pp_oks_loss_forward: scratch too small (%lld < %lld)", static_cast<long long>(scratch_bytes),
             static_cast<long long>(pp_oks_loss_scratch_bytes(p)));
  if (p->mode == PP_LOSS_PER_PIXEL) PP_REQUIRE(loss_map, PP_ERR_INVALID_ARG, "pp_oks_loss_forward: loss_map required");
  if (p->mode == PP_LOSS_PER_KEYPOINT) PP_REQUIRE(loss_kpt, PP_ERR_INVALID_ARG, "pp_oks_loss_forward: loss_kpt required");
  if (p->mode != PP_LOSS_PER_PIXEL) PP_REQUIRE(loss_scalar, PP_ERR_INVALID_ARG, "pp_oks_loss_forward: loss_scalar required");
  PP_REQUIRE(grad == nullptr || p->mode == PP_LOSS_PIXEL_MEAN, PP_ERR_INVALID_ARG,
             "pp_oks_loss_forward: the fused gradient exists for PP_LOSS_PIXEL_MEAN only");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t N = static_cast<int64_t>(p->B) * p->K;
  if (N == 0) return PP_OK;

  if (fast_path_ok(*p, output, target, pixel_weights, mask, grad)) {
    if (target_out_of_range) PP_CUDA_OK(cudaMemsetAsync(target_out_of_range, 0, sizeof(int32_t), st));
    int parts = 0;
    if (int rc = launch_fast(*p, output, target, keypoint_weights, grad, static_cast<double*>(scratch),
                             target_out_of_range, nullptr, grad_scale, true, st, &parts))
      return rc;
    PP_CUDA_OK(launch_finalize(static_cast<double*>(scratch), parts, 1.0 / (static_cast<double>(N) * p->H * p->W), loss_scalar,
                               mbv, N, publish != nullptr, true, st));
    return PP_OK;
  }

  LossArgs a{};
  a.p = *p;
  a.output = output; a.target = target; a.kp_weights = keypoint_weights; a.pix_weights = pixel_weights; a.mask = mask;
  a.loss_map = loss_map; a.loss_kpt = loss_kpt; a.peak_out = peak_index;
  a.partials = static_cast<double*>(scratch);
  a.grad = grad; a.upstream_kind = kUpScalar; a.host_scale = grad_scale; a.upstream = nullptr;
  a.range_flag = target_out_of_range;
  if (target_out_of_range) PP_CUDA_OK(cudaMemsetAsync(target_out_of_range, 0, sizeof(int32_t), st));
  int rc = (p->dtype == PP_F32) ? launch_loss<float>(a, true, grad != nullptr, st)
                                : launch_loss<__nv_bfloat16>(a, true, grad != nullptr, st);
  if (rc) return rc;
  if (p->mode != PP_LOSS_PER_PIXEL) {
    const double scale = (p->mode == PP_LOSS_PIXEL_MEAN) ? 1.0 / (static_cast<double>(N) * p->H * p->W) : 1.0 / static_cast<double>(N);
    PP_CUDA_OK(launch_finalize(a.partials, N, scale, loss_scalar, mbv, N, publish != nullptr, false, st));
  }
  return PP_OK;
}

int pp_oks_loss_forward_encoded(const pp_loss_params* p, const pp_encode_params* ep, const void* output,
                                const void* keypoints, const float* visible, const double* two_s,
                                const float* keypoint_weights, float* loss_scalar, void* grad, float grad_scale,
                                float* weights_out, uint8_t* in_image, uint8_t* annotated, void* scratch,
                                int64_t scratch_bytes, const pp_mailbox* publish, pp_stream_t stream) {
  if (int rc = check_loss_params("pp_oks_loss_forward_encoded", p)) return rc;
  if (int rc = check_publish("pp_oks_loss_forward_encoded", publish, static_cast<int64_t>(p->B) * p->K)) return rc;
  const pp_mailbox mbv = publish ? *publish : pp_mailbox{};
  PP_REQUIRE(ep != nullptr && ep->B == p->B && ep->K == p->K && ep->H == p->H && ep->W == p->W, PP_ERR_INVALID_ARG,
             "pp_oks_loss_forward_encoded: encode and loss parameters disagree on the shape");
  PP_REQUIRE(ep->keypoint_dim >= 2 && (ep->keypoint_dtype == PP_F32 || ep->keypoint_dtype == PP_F64) &&
                 ep->scale_x != 0.0f && ep->scale_y != 0.0f,
             PP_ERR_INVALID_ARG, "pp_oks_loss_forward_encoded: bad encode parameters");
  const int64_t N = static_cast<int64_t>(p->B) * p->K;
  if (N == 0) return PP_OK;
  PP_REQUIRE(output && keypoints && two_s && loss_scalar, PP_ERR_INVALID_ARG, "pp_oks_loss_forward_encoded: null argument");
  PP_REQUIRE(scratch && scratch_bytes >= pp_oks_loss_scratch_bytes(p), PP_ERR_SCRATCH,
             "pp_oks_loss_forward_encoded: scratch too small");
  PP_REQUIRE(p->mode == PP_LOSS_PIXEL_MEAN && fast_path_ok(*p, output, output, nullptr, nullptr, grad) && p->W + p->H <= 4096,
             PP_ERR_UNSUPPORTED_SHAPE,
             "pp_oks_loss_forward_encoded: needs PP_LOSS_PIXEL_MEAN without skip_empty_channel, W %% 4 == 0, 16-byte aligned "
             "planes that fit shared memory; encode with pp_encode and call pp_oks_loss_forward otherwise");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EncodeSource enc;
  enc.ep = ep; enc.keypoints = keypoints; enc.visible = visible; enc.two_s = two_s;
  enc.weights_out = weights_out; enc.in_image = in_image; enc.annotated = annotated;
  int parts = 0;
  if (int rc = launch_fast(*p, output, nullptr, keypoint_weights, grad, static_cast<double*>(scratch), nullptr, nullptr,
                           grad_scale, true, st, &parts, &enc))
    return rc;
  PP_CUDA_OK(launch_finalize(static_cast<double*>(scratch), parts, 1.0 / (static_cast<double>(N) * p->H * p->W), loss_scalar, mbv,
                             N, publish != nullptr, true, st));
  return PP_OK;
}

int pp_oks_loss_backward(const pp_loss_params* p, const void* output, const void* target, const float* keypoint_weights,
                         const void* pixel_weights, const void* mask, const void* upstream, int32_t upstream_kind,
                         const int32_t* peak_index, void* grad, void* scratch, int64_t scratch_bytes,
                         pp_stream_t stream) {
  (void)scratch; (void)scratch_bytes;
  if (int rc = check_loss_params("pp_oks_loss_backward", p)) return rc;
  if (static_cast<int64_t>(p->B) * p->K == 0) return PP_OK;
  PP_REQUIRE(output && target && upstream && grad, PP_ERR_INVALID_ARG, "pp_oks_loss_backward: null argument");
  PP_REQUIRE(p->mode != PP_LOSS_PER_KEYPOINT || peak_index, PP_ERR_INVALID_ARG,
             "pp_oks_loss_backward: peak_index required for PP_LOSS_PER_KEYPOINT");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (static_cast<int64_t>(p->B) * p->K == 0) return PP_OK;
  if (upstream_kind == PP_UPSTREAM_SCALAR && fast_path_ok(*p, output, target, pixel_weights, mask, grad)) {
    int parts = 0;
    return launch_fast(*p, output, target, keypoint_weights, grad, nullptr, nullptr, static_cast<const float*>(upstream),
                       1.0f, false, st, &parts);
  }
  LossArgs a{};
  a.p = *p;
  a.output = output; a.target = target; a.kp_weights = keypoint_weights; a.pix_weights = pixel_weights; a.mask = mask;
  a.grad = grad; a.upstream = upstream; a.peak_in = peak_index;
  a.host_scale = 1.0f;
  PP_REQUIRE(upstream_kind == PP_UPSTREAM_SCALAR || upstream_kind == PP_UPSTREAM_FULL, PP_ERR_INVALID_ARG,
             "pp_oks_loss_backward: bad upstream_kind %d", upstream_kind);
  PP_REQUIRE(!(upstream_kind == PP_UPSTREAM_FULL && p->mode == PP_LOSS_PIXEL_MEAN), PP_ERR_INVALID_ARG,
             "pp_oks_loss_backward: PP_LOSS_PIXEL_MEAN has a scalar upstream");
  a.upstream_kind = upstream_kind == PP_UPSTREAM_SCALAR ? kUpScalar : p->mode == PP_LOSS_PER_PIXEL ? kUpPerPixel : kUpPerKeypoint;
  return (p->dtype == PP_F32) ? launch_loss<float>(a, false, true, st) : launch_loss<__nv_bfloat16>(a, false, true, st);
}

int pp_scale_inplace(void* data, int32_t dtype, int64_t numel, const float* scale_dev, pp_stream_t stream) {
  PP_REQUIRE(data && scale_dev && numel >= 0, PP_ERR_INVALID_ARG, "pp_scale_inplace: bad argument");
  if (numel == 0) return PP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(std::min<int64_t>((numel + 255) / 256, static_cast<int64_t>(pp_sm_count()) * 8));
  if (dtype == PP_F32) scale_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(data), numel, scale_dev);
  else if (dtype == PP_BF16) scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(data), numel, scale_dev);
  else { pp_set_error("pp_scale_inplace: unsupported dtype %d", dtype); return PP_ERR_INVALID_ARG; }
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // extern "C"

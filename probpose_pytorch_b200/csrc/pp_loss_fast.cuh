// Fast path of the OKS heatmap loss: mean-per-pixel mode with per-heatmap weights (what
// ProbPoseLoss does with OKSHeatmapLoss, loss.py:428-431), forward + backward in one pass.
//
// HBM-bound by design: read `output` and `target` once, write `grad` once = 3 H W e bytes per heatmap.
// To stay under the ~66 instructions / pixel that the HBM roofline allows at fp32, the 3x3 Sobel
// stencil and its adjoint are evaluated with a register-rolling row pipeline:
//
//   * a persistent CTA pulls whole `output` heatmaps (contiguous H*W*e bytes) into shared memory with
//     1-D bulk async copies (TMA, cp.async.bulk + mbarrier), double buffered, so the next heatmap lands
//     while the current one is processed;
//   * a thread owns a strip 4 pixels wide and T rows tall.  Per row it reads 8 values of `output`
//     (the strip plus a 2-pixel halo) from shared memory, forms the separable row factors of the
//     Sobel pair (d = a[x-1]-a[x+1], s = a[x-1]+2a[x]+a[x+1]) for 6 columns, combines three rows into
//     gx, gy, scales them into P = 2 c gx, Q = 2 c gy, forms the row factors of the adjoint stencil
//     and, three rows later, emits one 128-bit gradient store.  The 1-pixel ring of gx/gy a strip needs
//     from its neighbours is recomputed instead of exchanged, so there is no shared-memory write and no
//     barrier inside a heatmap;
//   * `target` rides in the same TMA stage (second bulk copy on the same mbarrier): no thread ever
//     waits on a global load.
//
// Encode-inside-loss (kTgt == kTgtEncode, pp_oks_loss_forward_encoded): the target is the OKS probability map of a
// keypoint, t[y][x] = ex[x] ey[y] (generate_probmaps, codec.py:56-66, separable), so it is never read -- and never
// written by an encode kernel either: the CTA evaluates the W + H float64 factors of the next heatmap while it
// processes the current one, keeps them in shared memory as float32 and forms t with one multiply per pixel.  The
// pass then moves 2 H W e bytes per heatmap (read `output`, write `grad`) instead of 3 + the encode's 1.
#pragma once

#include "pp_common.cuh"

namespace pp_loss_fast {

using namespace pp;

struct FastArgs {
  const void* output;
  const void* target;
  const float* kp_weights;   // (N) or null
  void* grad;                // (N, H, W) or null
  double* partials;          // one per CTA (forward)
  int32_t* range_flag;       // or null
  const float* upstream;     // device scalar or null
  float host_scale;
  long long N;
  int H, W, strips, segs, T, stages;
  int G;                     // consecutive heatmaps processed together by one CTA (one TMA copy)
  float w_s, w_o, w_g, lw;
  float a_o, a_t;            // oks(o, t) = a_o o + a_t t - o t
  float d_a, d_b;            // d oks / d o = d_a + d_b t
  float inv_count;           // 1 / (N H W)
  unsigned plane_bytes, stage_bytes, tgt_off;   // tgt_off: byte offset of the target planes inside a stage
  // encode-inside-loss: the target of heatmap (b, k) is the probability map of keypoint (b, k)
  const void* keypoints;     // (N, kp_dim) input-image space, float32 or float64 (kp_f64)
  const float* visible;      // (N) or null (all ones)
  const double* two_s;       // (K) divisor 2 s per keypoint (codec.py:65)
  float* weights_out;        // (N) keypoint weights as the encoder defines them (codec.py:46,68), or null
  uint8_t* in_image;         // (N) or null (codec.py:189-200)
  uint8_t* annotated;        // (N) or null (codec.py:187)
  int K, kp_dim, kp_f64;
  float scale_x, scale_y, input_w, input_h;
  unsigned fac_off;          // byte offset of the factor buffers (after the stages): 2 x G x (W + H + 4) floats
};

enum TgtMode : int { kTgtGlobal = 0, kTgtSmem = 1, kTgtEncode = 2 };

struct RowState {
  float hd[3][6], hs[3][6];  // row factors of the Sobel pair, rows r-2, r-1, r
  float dP[3][4], sQ[3][4];  // row factors of the adjoint stencil, rows r-3, r-2, r-1
};

template <typename T>
__device__ __forceinline__ void load_strip8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load_strip8<float>(const float* p, float (&v)[8]) {
  const float2 l = *reinterpret_cast<const float2*>(p - 2);
  const float4 c = *reinterpret_cast<const float4*>(p);
  const float2 r = *reinterpret_cast<const float2*>(p + 4);
  v[0] = l.x; v[1] = l.y; v[2] = c.x; v[3] = c.y; v[4] = c.z; v[5] = c.w; v[6] = r.x; v[7] = r.y;
}
template <>
__device__ __forceinline__ void load_strip8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint32_t l = *reinterpret_cast<const uint32_t*>(p - 2);
  const uint2 c = *reinterpret_cast<const uint2*>(p);
  const uint32_t r = *reinterpret_cast<const uint32_t*>(p + 4);
  v[0] = __uint_as_float(l << 16); v[1] = __uint_as_float(l & 0xffff0000u);
  v[2] = __uint_as_float(c.x << 16); v[3] = __uint_as_float(c.x & 0xffff0000u);
  v[4] = __uint_as_float(c.y << 16); v[5] = __uint_as_float(c.y & 0xffff0000u);
  v[6] = __uint_as_float(r << 16); v[7] = __uint_as_float(r & 0xffff0000u);
}

template <typename T>
__device__ __forceinline__ void load_strip4(const T* p, float (&v)[4]);   // shared memory, strip-aligned
template <>
__device__ __forceinline__ void load_strip4<float>(const float* p, float (&v)[4]) {
  const float4 w = *reinterpret_cast<const float4*>(p);
  v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
}
template <>
__device__ __forceinline__ void load_strip4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(w.x << 16); v[1] = __uint_as_float(w.x & 0xffff0000u);
  v[2] = __uint_as_float(w.y << 16); v[3] = __uint_as_float(w.y & 0xffff0000u);
}

template <typename T>
__device__ __forceinline__ void store_grad4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store_grad4<float>(float* p, const float (&v)[4]) {
  stg_stream_128(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
}
template <>
__device__ __forceinline__ void store_grad4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  const uint32_t x = *reinterpret_cast<uint32_t*>(&a), y = *reinterpret_cast<uint32_t*>(&b);
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}

struct Coef {
  float c2;            // 2 lw w_s u m   (P = c2 gx, Q = c2 gy)
  float k_a, k_b, k_g; // direct gradient = k_a + k_b t + k_g (o - t)
  float a_o, a_t;
  bool has_mse;
};

struct Sums {
  float se, so, sm, tmin, tmax;
};

// One row step of the pipeline; PH = step index mod 3 selects the rotating register slots.
template <typename T>
__device__ __forceinline__ void load_strip4_global(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load_strip4_global<float>(const float* p, float (&v)[4]) {
  const uint4 w = ldg_stream_128(p);
  v[0] = __uint_as_float(w.x); v[1] = __uint_as_float(w.y); v[2] = __uint_as_float(w.z); v[3] = __uint_as_float(w.w);
}
template <>
__device__ __forceinline__ void load_strip4_global<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 w;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p));
  v[0] = __uint_as_float(w.x << 16); v[1] = __uint_as_float(w.x & 0xffff0000u);
  v[2] = __uint_as_float(w.y << 16); v[3] = __uint_as_float(w.y & 0xffff0000u);
}

// kTgt: kTgtSmem -- the target plane was staged in shared memory with the output plane (the normal case);
// kTgtGlobal -- (maps so large that only one plane fits) it is read from global memory where it is consumed;
// kTgtEncode -- it is the product of the keypoint's separable factors: tx[4] (this strip's columns, registers) x ey[row].
template <typename T, bool kFwd, bool kGrad, int kTgt, int PH>
__device__ __forceinline__ void row_step(RowState& st, Sums& sums, const Coef& cf, int q, int y0, int y1, int H, int W,
                                         int x0, bool left_ok, bool right_ok, const T* __restrict__ plane,
                                         const T* __restrict__ tgt, T* __restrict__ grad, const float (&tx)[4],
                                         const float* __restrict__ ey) {
  constexpr int cur = PH, p1 = (PH + 2) % 3, p2 = (PH + 1) % 3;  // rows r, r-1, r-2 (and r-3 == cur for dP/sQ)
  const int r = y0 - 2 + q;

  // 1. row r of `output` -> row factors of the Sobel pair at columns x0-1 .. x0+4
  float av[8];
  if (r >= 0 && r < H) {
    load_strip8<T>(plane + r * W + x0, av);
    if (!left_ok) { av[0] = 0.0f; av[1] = 0.0f; }
    if (!right_ok) { av[6] = 0.0f; av[7] = 0.0f; }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = 0.0f;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    st.hd[cur][j] = av[j] - av[j + 2];
    st.hs[cur][j] = (av[j] + av[j + 2]) + 2.0f * av[j + 1];
  }

  // 2. gx, gy at row r-1, columns x0-1 .. x0+4
  const int rm = r - 1;
  float gx[6], gy[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    gx[j] = (st.hd[p2][j] + st.hd[cur][j]) + 2.0f * st.hd[p1][j];
    gy[j] = st.hs[p2][j] - st.hs[cur][j];
  }
  if (kFwd && rm >= y0 && rm < y1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sums.se += gx[1 + i] * gx[1 + i] + gy[1 + i] * gy[1 + i];
  }
  if (kGrad) {
    const float cr = (rm >= 0 && rm < H) ? cf.c2 : 0.0f;
    float P[6], Q[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { P[j] = cr * gx[j]; Q[j] = cr * gy[j]; }
    if (!left_ok) { P[0] = 0.0f; Q[0] = 0.0f; }     // column -1 is outside the map
    if (!right_ok) { P[5] = 0.0f; Q[5] = 0.0f; }    // column W is outside the map
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      st.dP[p1][i] = P[i] - P[i + 2];
      st.sQ[p1][i] = (Q[i] + Q[i + 2]) + 2.0f * Q[i + 1];
    }
  }

  // 3. emit row r-2
  const int ro = r - 2;
  if (ro >= y0 && ro < y1) {
    float g[4], ov[4], tv[4];
    load_strip4<T>(plane + ro * W + x0, ov);
    if (kTgt == kTgtSmem) load_strip4<T>(tgt + ro * W + x0, tv);
    else if (kTgt == kTgtGlobal) load_strip4_global<T>(tgt + ro * W + x0, tv);
    else {
      const float fy = ey[ro];
#pragma unroll
      for (int i = 0; i < 4; ++i) tv[i] = tx[i] * fy;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float o = ov[i], t = tv[i];
      if (kFwd) {
        sums.so += cf.a_o * o + cf.a_t * t - o * t;
        if (cf.has_mse) { const float d = o - t; sums.sm += d * d; }
        if (kTgt != kTgtEncode) {   // an encoded target lies in [0, 1] by construction
          sums.tmin = fminf(sums.tmin, t);
          sums.tmax = fmaxf(sums.tmax, t);
        }
      }
      if (kGrad) {
        float direct = cf.k_a + cf.k_b * t;
        if (cf.has_mse) direct += cf.k_g * (o - t);
        const float sx = (st.dP[cur][i] + st.dP[p1][i]) + 2.0f * st.dP[p2][i];  // rows r-3, r-1, r-2
        const float sy = st.sQ[cur][i] - st.sQ[p1][i];                           // rows r-3, r-1
        g[i] = (direct - sx) - sy;
      }
    }
    if (kGrad) store_grad4<T>(grad + ro * W + x0, g);
  }
}

// float64 -> heatmap space like NumPy does (float32 keypoints / float32 scale stay float32, codec.py:180)
__device__ __forceinline__ double fast_to_heatmap_space(const void* kp, int f64, long long idx, float scale) {
  if (f64) return __ddiv_rn(static_cast<const double*>(kp)[idx], static_cast<double>(scale));
  return static_cast<double>(__fdiv_rn(static_cast<const float*>(kp)[idx], scale));
}

// Encode-inside-loss: the W + H separable factors of heatmap h (generate_probmaps, codec.py:45-68) as float32 into
// f[0 .. W + H), the keypoint weight into f[W + H], and the flag outputs of ProbMap.encode / ArgMaxProbMap.encode
// (codec.py:187-200).  Called by the `per` threads of a heatmap slot (thread `local`).  Out of line on purpose: the
// float64 divisions and the exponential are a few hundred instructions that would otherwise sit in the middle of the
// persistent row loop (inlined, the kernel was twice the size of the plain one and 20 % slower than it).
struct EncodeArgs {   // what encode_factors needs of FastArgs, passed by value (a reference would move the kernel's whole
                      // parameter block into local memory)
  const void* keypoints;
  const float* visible;
  const double* two_s;
  const float* kp_weights;
  float* weights_out;
  uint8_t* in_image;
  uint8_t* annotated;
  int W, H, K, kp_dim, kp_f64;
  float scale_x, scale_y, input_w, input_h;
};

__device__ __noinline__ void encode_factors(const EncodeArgs a, long long h, float* __restrict__ f, int local, int per) {
  const int W = a.W, H = a.H;
  const float vis = a.visible ? a.visible[h] : 1.0f;
  const bool labelled = !(vis < 0.5f);   // codec.py:53
  const double kx = fast_to_heatmap_space(a.keypoints, a.kp_f64, h * a.kp_dim, a.scale_x);
  const double ky = fast_to_heatmap_space(a.keypoints, a.kp_f64, h * a.kp_dim + 1, a.scale_y);
  const double div = a.two_s[h % a.K];
  // The reference evaluates exp(-(dx^2 + dy^2) / (2 s)) in float64 and stores float32 (codec.py:56-66).  Here the
  // exponent is formed in float64 and the exponential taken in float32 per factor: relative error of a factor
  // <= 2.4e-7 |exponent| + 1 ulp: the ABSOLUTE error of a target value t is <= 2.4e-7 t |ln t| <= 9e-8 -- far inside
  // the 1e-5 budget of the loss and its gradient -- at a fifth of the instructions of a float64 exp + division.
  const float inv_div = 1.0f / static_cast<float>(div);
  for (int j = local; j < W + H; j += per) {
    const double d = (j < W) ? (static_cast<double>(j) - kx) : (static_cast<double>(j - W) - ky);
    f[j] = labelled ? expf(-(static_cast<float>(d * d) * inv_div)) : 0.0f;   // unlabelled channels stay zero (codec.py:45)
  }
  if (local != 0) return;
  float wgt = vis;   // unlabelled keypoints keep their visibility value (codec.py:46,53-54)
  if (labelled) {    // (float64 map).max() > 0 (codec.py:68): the maximum sits at the grid point nearest to the keypoint
    const double xn = fmin(fmax(rint(kx), 0.0), static_cast<double>(W - 1));
    const double yn = fmin(fmax(rint(ky), 0.0), static_cast<double>(H - 1));
    const double dx = xn - kx, dy = yn - ky;
    // exp(-q) > 0 in float64  <=>  q < 745.13321910194...: the smallest subnormal is exp(-744.44), results round to
    // it down to exp(-745.13).  A compare instead of a float64 exp keeps this small.
    wgt = (dx * dx + dy * dy) / div < 745.1332191019411 ? 1.0f : 0.0f;
  }
  f[W + H] = a.kp_weights ? a.kp_weights[h] : wgt;   // explicit weights (learn_heatmaps_from_zeros etc.) win
  if (a.weights_out) a.weights_out[h] = wgt;
  if (a.annotated) a.annotated[h] = vis > 0.0f;
  if (a.in_image) {
    bool in;
    if (a.kp_f64) {
      const double x = static_cast<const double*>(a.keypoints)[h * a.kp_dim], y = static_cast<const double*>(a.keypoints)[h * a.kp_dim + 1];
      in = x >= 0.0 && x < static_cast<double>(a.input_w) && y >= 0.0 && y < static_cast<double>(a.input_h);
    } else {
      const float x = static_cast<const float*>(a.keypoints)[h * a.kp_dim], y = static_cast<const float*>(a.keypoints)[h * a.kp_dim + 1];
      in = x >= 0.0f && x < a.input_w && y >= 0.0f && y < a.input_h;
    }
    a.in_image[h] = in;
  }
}

template <typename T, bool kFwd, bool kGrad, int kTgt>
__global__ void __launch_bounds__(256)
oks_loss_fast_kernel(FastArgs a) {
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ double red[8];
  __shared__ int red_flag;

  // the one-block kernel that sums this grid's partial sums is launched programmatically dependent on it (pp_loss.cu,
  // launch_finalize): let it become resident now -- it waits (griddepcontrol.wait) until this grid has completed and
  // flushed -- instead of paying its launch latency after this grid's last CTA
  asm volatile("griddepcontrol.launch_dependents;");
  const int tid = threadIdx.x;
  const int H = a.H, W = a.W;
  const long long HW = static_cast<long long>(H) * W;
  const int per = a.strips * a.segs;           // threads per heatmap
  const int g = tid / per, local = tid - g * per;
  const int sx = local % a.strips, sy = local / a.strips;
  const int x0 = sx * 4, y0 = sy * a.T, y1 = min(y0 + a.T, H);
  const bool left_ok = sx > 0, right_ok = sx < a.strips - 1;
  const T* out = static_cast<const T*>(a.output);
  const T* tgt_all = static_cast<const T*>(a.target);
  auto tgt_stage_of = [&](int s) { return reinterpret_cast<const T*>(stage_mem + static_cast<size_t>(s) * a.stage_bytes + a.tgt_off); };
  // encode-inside-loss: factor buffer b (alternating per unit) of heatmap slot g: ex[W], ey[H], then the keypoint weight
  const int FS = W + H + 4;
  auto fac_of = [&](int b, int gslot) { return reinterpret_cast<float*>(stage_mem + a.fac_off) + (static_cast<size_t>(b) * a.G + gslot) * FS; };
  auto encode_unit = [&](long long un, int b) {
    if (g < a.G && un * a.G + g < a.N) {
      const EncodeArgs e{a.keypoints, a.visible, a.two_s, a.kp_weights, a.weights_out, a.in_image, a.annotated,
                         a.W, a.H, a.K, a.kp_dim, a.kp_f64, a.scale_x, a.scale_y, a.input_w, a.input_h};
      encode_factors(e, un * a.G + g, fac_of(b, g), local, per);
    }
  };
  T* grad_all = static_cast<T*>(a.grad);
  // 16 bytes of slack in front of / behind each stage keep the halo reads of the first / last strip in bounds
  auto stage_of = [&](int s) { return reinterpret_cast<const T*>(stage_mem + static_cast<size_t>(s) * a.stage_bytes + 16); };
  const long long units = (a.N + a.G - 1) / a.G;
  auto unit_bytes = [&](long long u) {
    const long long left = a.N - u * a.G;
    return static_cast<unsigned>((left < a.G ? left : a.G) * a.plane_bytes);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    red_flag = 0;
  }
  __syncthreads();

  float u = a.host_scale * a.inv_count;
  if (kGrad && a.upstream) u *= a.upstream[0];

  long long unit = blockIdx.x;
  auto fetch_unit = [&](long long un, int s) {   // `output` and `target` of a unit land in stage s
    const unsigned bytes = unit_bytes(un);
    mbar_expect_tx(&bars[s], kTgt == kTgtSmem ? 2 * bytes : bytes);
    tma_load_1d(const_cast<T*>(stage_of(s)), out + un * a.G * HW, bytes, &bars[s]);
    if (kTgt == kTgtSmem) tma_load_1d(const_cast<T*>(tgt_stage_of(s)), tgt_all + un * a.G * HW, bytes, &bars[s]);
  };
  if (tid == 0 && unit < units) fetch_unit(unit, 0);
  if (kTgt == kTgtEncode) {
    if (unit < units) encode_unit(unit, 0);
    __syncthreads();
  }

  double acc = 0.0;
  Sums sums{0.f, 0.f, 0.f, INFINITY, -INFINITY};
  const int nsteps = (y1 - y0) + 4;

  for (int it = 0; unit < units; unit += gridDim.x, ++it) {
    const int s = (a.stages == 2) ? (it & 1) : 0;
    const long long nxt = unit + gridDim.x;
    if (a.stages == 2 && tid == 0 && nxt < units) fetch_unit(nxt, s ^ 1);  // prefetch into the other stage
    const long long hm = unit * a.G + g;
    const bool active = g < a.G && hm < a.N && sy < a.segs;
    float m = 1.0f;
    float tx[4] = {0.f, 0.f, 0.f, 0.f};
    const float* ey = nullptr;
    if (kTgt == kTgtEncode) {
      if (active) {
        const float* f = fac_of(it & 1, g);
        m = f[W + H];
#pragma unroll
        for (int i = 0; i < 4; ++i) tx[i] = f[x0 + i];
        ey = f + W;
      }
    } else if (active && a.kp_weights) {
      m = a.kp_weights[hm];
    }
    Coef cf;
    cf.c2 = 2.0f * a.lw * a.w_s * u * m;
    cf.k_a = a.lw * u * m * a.w_o * a.d_a;
    cf.k_b = a.lw * u * m * a.w_o * a.d_b;
    cf.k_g = 2.0f * a.lw * u * m * a.w_g;
    cf.a_o = a.a_o; cf.a_t = a.a_t;
    cf.has_mse = a.w_g != 0.0f;

    T* grad = kGrad ? grad_all + hm * HW : nullptr;
    RowState st;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { st.hd[i][j] = 0.f; st.hs[i][j] = 0.f; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { st.dP[i][j] = 0.f; st.sQ[i][j] = 0.f; }
    }
    sums.se = sums.so = sums.sm = 0.f;

    mbar_wait(&bars[s], (a.stages == 2) ? ((it >> 1) & 1) : (it & 1));
    const T* plane = stage_of(s) + static_cast<size_t>(g) * HW;
    const T* tgt = kTgt == kTgtSmem ? tgt_stage_of(s) + static_cast<size_t>(g) * HW : kTgt == kTgtGlobal ? tgt_all + hm * HW : nullptr;
    if (active) {
      for (int q = 0; q < nsteps; q += 3) {
        row_step<T, kFwd, kGrad, kTgt, 0>(st, sums, cf, q, y0, y1, H, W, x0, left_ok, right_ok, plane, tgt, grad, tx, ey);
        if (q + 1 < nsteps)
          row_step<T, kFwd, kGrad, kTgt, 1>(st, sums, cf, q + 1, y0, y1, H, W, x0, left_ok, right_ok, plane, tgt, grad, tx, ey);
        if (q + 2 < nsteps)
          row_step<T, kFwd, kGrad, kTgt, 2>(st, sums, cf, q + 2, y0, y1, H, W, x0, left_ok, right_ok, plane, tgt, grad, tx, ey);
      }
      if (kFwd) {
        // per-pixel loss = (w_s e + w_o oks + w_g mse) m lw (loss.py:122-127, 143), summed per strip
        const float part = (a.w_s * sums.se + a.w_o * sums.so + a.w_g * sums.sm) * (m * a.lw);
        acc += static_cast<double>(part);
      }
    }
    // the next unit's target factors go to the other factor buffer (last read one iteration ago); the barrier
    // below publishes them
    if (kTgt == kTgtEncode && nxt < units) encode_unit(nxt, (it + 1) & 1);
    __syncthreads();  // every thread is done with stage s before it is refilled
    if (a.stages == 1 && tid == 0 && nxt < units) fetch_unit(nxt, 0);
  }

  if (kFwd) {
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    if (kTgt != kTgtEncode && a.range_flag && (sums.tmin < 0.0f || sums.tmax > 1.0f)) red_flag = 1;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
      a.partials[blockIdx.x] = t;
      if (a.range_flag && red_flag) atomicOr(a.range_flag, 1);
    }
  }
}

}  // namespace pp_loss_fast

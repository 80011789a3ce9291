// Heatmap decoders.
//
//  * pp_decode_expected  : get_heatmap_expected_value (heatmap.py:291-395) + ProbMap.decode scaling
//                          (codec.py:214-239).  OKS-kernel reflect-mode convolution, argmax, quadratic
//                          sub-pixel refinement, score read from the unconvolved map.
//  * pp_heatmap_maximum  : get_heatmap_maximum (heatmap.py:13-52).
//  * pp_decode_argmax_dark: ArgMaxProbMap.decode (codec.py:515-543) = argmax + gaussian_blur
//                          (codec.py:284-313) + refine_keypoints_dark_udp (codec.py:315-375).
//  * pp_heatmap_tail     : clamp(x / temperature, 0, 1) (head.py:526-532); also fusable into the
//                          decoders' load (pp_decode_params.apply_tail).
//
// Layout: one CTA owns one heatmap at a time (grid-stride over the B*K heatmaps).  The heatmap is
// read from HBM exactly once with 128-bit streaming loads into a padded shared-memory plane; all
// further passes run out of shared memory / registers.  Algorithmic HBM bytes = H*W*sizeof(T).
//
// Bit-exact argmax through the convolution.  scipy accumulates the d x d kernel in double and stores
// float32, and the argmax is taken on those float32 values (heatmap.py:362-369).  Doing that for
// every pixel is fp64-pipe bound, so the kernel
//   (1) runs a separable float32 convolution (register sliding window, 2d FFMA/pixel) -> P(p),
//   (2) bounds |P(p) - exact(p)| <= E = gamma * max|h| with gamma = (2d + 8) * 2^-23 (a 4x margin over
//       the worst-case float32 rounding of the two passes incl. tap rounding; sum of taps == 1),
//   (3) re-evaluates in float64, with the reference's own d x d table, only the pixels with
//       P(p) >= max P - (2E + 4u max|h|): the float32-rounded exact maximum can only be one of those,
//   (4) picks the maximum of the exact float32 values with the lowest flat index on ties (NumPy
//       argmax), and evaluates the four neighbours exactly for the sub-pixel fit.
#include <algorithm>
#include <cstdio>

#include "pp_common.cuh"

namespace {

using namespace pp;

constexpr int kPadL = 16;        // left/right pad of the staged plane (>= any radius, multiple of 4)
constexpr int kTile = 8;         // outputs per thread along the filtered axis
constexpr int kMaxCand = 64;     // candidate list; more than that -> ballot scan
constexpr int kMaxWarps = 32;

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
// Row stride (in floats) that is a multiple of 4 and whose quarter is odd: 128-bit shared-memory
// accesses by threads on consecutive rows are then bank-conflict free.
__host__ __device__ inline int conflict_free_stride(int w) {
  int s = round_up(w, 4);
  if (((s >> 2) & 1) == 0) s += 4;
  return s;
}

struct PlaneGeom {
  int W8, SP, ST;        // padded width, staged-plane stride, row-filtered-plane stride
  int raw_floats, tmp_floats, out_floats;
};
__host__ __device__ inline PlaneGeom plane_geom(int H, int W, int tmp_extra_rows) {
  PlaneGeom g;
  g.W8 = round_up(W, kTile);
  g.SP = conflict_free_stride(kPadL + g.W8 + kPadL);
  g.ST = conflict_free_stride(g.W8);
  g.raw_floats = H * g.SP;
  g.tmp_floats = (H + tmp_extra_rows) * g.ST;
  g.out_floats = round_up(H * W, 4);
  return g;
}

// ---------------------------------------------------------------------------
// staging: global -> padded shared plane, with the optional head tail fused in
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float apply_tail(float v, bool tail, float temperature) {
  if (!tail) return v;
  v = __fdiv_rn(v, temperature);
  v = Elem<T>::to_f32(Elem<T>::from_f32(v));  // torch rounds x / t to the tensor dtype before the clamp
  return fminf(fmaxf(v, 0.0f), 1.0f);
}

struct LoadStats {
  float vmax, vmin;
  int imax;  // flat index of the first maximum
};

template <typename T>
__device__ __forceinline__ LoadStats stage_plane(const T* __restrict__ src, float* __restrict__ raw, int H, int W,
                                                 int SP, bool vector_ok, bool tail, float temperature) {
  constexpr int V = Elem<T>::kVec;
  LoadStats s{-INFINITY, INFINITY, 0x7fffffff};
  const int HW = H * W;
  if (vector_ok) {
    const int WV = W / V;
    for (int i = threadIdx.x; i < HW / V; i += blockDim.x) {
      const int y = i / WV, xv = i - y * WV;
      float f[V];
      unpack(ldg_stream_128(src + i * V), f, T());
#pragma unroll
      for (int j = 0; j < V; ++j) {
        f[j] = apply_tail<T>(f[j], tail, temperature);
        if (f[j] > s.vmax) { s.vmax = f[j]; s.imax = i * V + j; }
        s.vmin = fminf(s.vmin, f[j]);
      }
      float4* dst = reinterpret_cast<float4*>(raw + y * SP + kPadL + xv * V);
#pragma unroll
      for (int j = 0; j < V / 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int y = i / W, x = i - y * W;
      const float v = apply_tail<T>(Elem<T>::to_f32(src[i]), tail, temperature);
      if (v > s.vmax) { s.vmax = v; s.imax = i; }
      s.vmin = fminf(s.vmin, v);
      raw[y * SP + kPadL + x] = v;
    }
  }
  return s;
}

// Fill the left/right pads of every row: reflect (scipy 'reflect') or zeros.  All columns that a
// partial tile can touch are written so that no uninitialised shared memory is ever read.
__device__ __forceinline__ void fill_pads(float* raw, int H, int W, int SP, bool reflect) {
  const int right = SP - kPadL - W;
  const int per_row = kPadL + right;
  for (int i = threadIdx.x; i < H * per_row; i += blockDim.x) {
    const int y = i / per_row, j = i - y * per_row;
    const int col = (j < kPadL) ? (j - kPadL) : (W + j - kPadL);
    float* row = raw + y * SP + kPadL;
    row[col] = reflect ? row[reflect_index(col, W)] : 0.0f;
  }
}

// ---------------------------------------------------------------------------
// separable float32 filter passes (register sliding window, R = radius)
// ---------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void row_pass(const float* __restrict__ raw, float* __restrict__ tmp,
                                         const float* __restrict__ taps, int H, int W8, int SP, int ST) {
  constexpr int PADR = (R + 3) & ~3;
  constexpr int WIN = kTile + 2 * PADR;
  float g[R + 1];
#pragma unroll
  for (int j = 0; j <= R; ++j) g[j] = taps[j];  // symmetric: tap[j] == tap[2R - j]
  const int tasks = (W8 / kTile) * H;
  for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
    const int xb = t / H, y = t - xb * H;  // y fastest: conflict-free 128-bit accesses
    const float4* src = reinterpret_cast<const float4*>(raw + y * SP + kPadL + xb * kTile - PADR);
    float in[WIN];
#pragma unroll
    for (int c = 0; c < WIN / 4; ++c) {
      const float4 v = src[c];
      in[4 * c] = v.x; in[4 * c + 1] = v.y; in[4 * c + 2] = v.z; in[4 * c + 3] = v.w;
    }
    float acc[kTile];
#pragma unroll
    for (int o = 0; o < kTile; ++o) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j <= 2 * R; ++j) a = fmaf(g[j <= R ? j : 2 * R - j], in[PADR - R + o + j], a);
      acc[o] = a;
    }
    float4* dst = reinterpret_cast<float4*>(tmp + y * ST + xb * kTile);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// tmp rows are addressed through `row_of(y)`: reflect for the OKS convolution, or an offset into a
// zero-bordered plane for the zero-padded blur.
template <int R, bool kReflect>
__device__ __forceinline__ void col_pass(const float* __restrict__ tmp, float* __restrict__ out,
                                         const float* __restrict__ taps, int H, int W, int W8, int ST) {
  float g[R + 1];
#pragma unroll
  for (int j = 0; j <= R; ++j) g[j] = taps[j];
  const int tasks = W8 * ((H + kTile - 1) / kTile);
  for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
    const int yb = t / W8, x = t - yb * W8;
    if (x >= W) continue;
    const int y0 = yb * kTile;
    float in[kTile + 2 * R];
#pragma unroll
    for (int j = 0; j < kTile + 2 * R; ++j) {
      const int yy = y0 - R + j;
      // zero-padded variant: tmp has R zero rows above and >= R + kTile zero rows below
      in[j] = kReflect ? tmp[reflect_index(yy, H) * ST + x] : tmp[(yy + R) * ST + x];
    }
#pragma unroll
    for (int o = 0; o < kTile; ++o) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j <= 2 * R; ++j) a = fmaf(g[j <= R ? j : 2 * R - j], in[o + j], a);
      if (y0 + o < H) out[(y0 + o) * W + x] = a;
    }
  }
}

// Runtime-radius fallbacks (any radius <= kPadL): one pixel per thread-iteration.
__device__ __forceinline__ void row_pass_generic(const float* raw, float* tmp, const float* taps, int r, int H, int W,
                                                 int SP, int ST) {
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float* p = raw + y * SP + kPadL + x - r;
    float a = 0.0f;
    for (int j = 0; j <= 2 * r; ++j) a = fmaf(taps[j], p[j], a);
    tmp[y * ST + x] = a;
  }
}
template <bool kReflect>
__device__ __forceinline__ void col_pass_generic(const float* tmp, float* out, const float* taps, int r, int H, int W,
                                                 int ST) {
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    float a = 0.0f;
    for (int j = 0; j <= 2 * r; ++j) {
      const int yy = y - r + j;
      const float v = kReflect ? tmp[reflect_index(yy, H) * ST + x] : ((yy >= 0 && yy < H) ? tmp[yy * ST + x] : 0.0f);
      a = fmaf(taps[j], v, a);
    }
    out[i] = a;
  }
}

template <bool kReflect>
__device__ __forceinline__ void separable_filter(const float* raw, float* tmp, float* out, const float* taps, int r,
                                                 int H, int W, const PlaneGeom& g) {
  // kReflect == false expects `tmp` to point R rows into a zero-bordered plane only for the
  // specialised radius (5); the generic path bounds-checks instead.
  switch (kReflect ? r : (r == 5 ? 5 : 0)) {
#define PP_CASE(R)                                                   \
  case R:                                                            \
    row_pass<R>(raw, tmp + (kReflect ? 0 : R * g.ST), taps, H, g.W8, g.SP, g.ST); \
    __syncthreads();                                                 \
    col_pass<R, kReflect>(tmp, out, taps, H, W, g.W8, g.ST);         \
    break;
    PP_CASE(1) PP_CASE(2) PP_CASE(3) PP_CASE(4) PP_CASE(5) PP_CASE(6) PP_CASE(7) PP_CASE(8) PP_CASE(9)
#undef PP_CASE
    default:
      row_pass_generic(raw, tmp, taps, r, H, W, g.SP, g.ST);
      __syncthreads();
      col_pass_generic<kReflect>(tmp, out, taps, r, H, W, g.ST);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// block reductions through a small shared scratch
// ---------------------------------------------------------------------------
struct BlockScratch {
  float fv[kMaxWarps];
  float fw[kMaxWarps];
  int iv[kMaxWarps];
  int cand[kMaxCand];
  int cand_count;
  int best_idx;
  float best_val;
  float nb[4];
};

__device__ __forceinline__ float block_max(float v, BlockScratch& s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) s.fv[warp] = v;
  __syncthreads();
  float r = s.fv[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, s.fv[i]);
  return r;
}
__device__ __forceinline__ float block_min(float v, BlockScratch& s) { return -block_max(-v, s); }

__device__ __forceinline__ void block_argmax(float& v, int& idx, BlockScratch& s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  warp_argmax(v, idx);
  __syncthreads();
  if (lane == 0) { s.fv[warp] = v; s.iv[warp] = idx; }
  __syncthreads();
  v = s.fv[0]; idx = s.iv[0];
  for (int i = 1; i < nw; ++i) argmax_combine(v, idx, s.fv[i], s.iv[i]);
}

// ---------------------------------------------------------------------------
// exact evaluation of one convolved pixel: double accumulation of the d x d table, float32 result
// (what scipy.ndimage.convolve stores, heatmap.py:362-364).  Warp-collective; all lanes return the
// same value.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float exact_conv_at(const float* __restrict__ raw, int SP, int H, int y, int x, int r,
                                               const double* __restrict__ w2d) {
  const int d = 2 * r + 1, n = d * d, lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) {
    const int ti = i / d, tj = i - ti * d;
    const int yy = reflect_index(y + ti - r, H);
    acc = fma(w2d[i], static_cast<double>(raw[yy * SP + kPadL + x + tj - r]), acc);
  }
  return static_cast<float>(warp_sum(acc));
}

// ---------------------------------------------------------------------------
// expected-OKS decoder
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512)
decode_expected_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps, float* __restrict__ locs,
                       float* __restrict__ vals, int32_t* __restrict__ argmax, double* __restrict__ keypoints,
                       bool vector_ok, bool only_marked) {
  extern __shared__ __align__(16) float smem[];
  __shared__ BlockScratch bs;
  __shared__ float taps[PP_OKS_TAPS];

  const int H = p.H, W = p.W, HW = H * W;
  const PlaneGeom g = plane_geom(H, W, 0);
  float* raw = smem;
  float* tmp = raw + g.raw_floats;
  float* conv = tmp + g.tmp_floats;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;

  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    // second launch after the pruned kernel: only the heatmaps it marked (NaN in locs) are left
    if (only_marked && !isnan(locs[hm * 2])) continue;
    const int k = static_cast<int>(hm % p.K);
    const int r = tab.radius[k];
    const double* w2d = tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;
    if (threadIdx.x < PP_OKS_TAPS) taps[threadIdx.x] = tab.taps_f32[k * PP_OKS_TAPS + threadIdx.x];
    if (threadIdx.x == 0) { bs.cand_count = 0; bs.best_idx = 0x7fffffff; bs.best_val = -INFINITY; }

    LoadStats st = stage_plane<T>(heatmaps + hm * HW, raw, H, W, g.SP, vector_ok, p.apply_tail != 0, p.temperature);
    const float vmax = block_max(st.vmax, bs);
    const float vmin = block_min(st.vmin, bs);   // (contains the barriers that publish raw/taps/bs)

    int best = 0;
    float best_val = 0.0f;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    bool interior = false;

    if (vmax == vmin) {
      // Constant map: every convolved pixel is the same float (same taps, same values, same order),
      // so the first index wins; (0,0) is a border pixel -> no sub-pixel shift.
      best = 0;
    } else {
      fill_pads(raw, H, W, g.SP, /*reflect=*/true);
      __syncthreads();
      separable_filter<true>(raw, tmp, conv, taps, r, H, W, g);

      float pmax = -INFINITY;
      for (int i = threadIdx.x; i < HW; i += blockDim.x) pmax = fmaxf(pmax, conv[i]);
      pmax = block_max(pmax, bs);
      const float amax = fmaxf(fabsf(vmax), fabsf(vmin));
      const float gamma = static_cast<float>(2 * (2 * r + 1) + 8) * 1.1920929e-7f;  // (2d + 8) * 2^-23
      const float thr = pmax - (2.0f * gamma + 4.0f * 5.9604645e-8f) * amax;

      for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        if (conv[i] >= thr) {
          const int slot = atomicAdd(&bs.cand_count, 1);
          if (slot < kMaxCand) bs.cand[slot] = i;
        }
      }
      __syncthreads();
      const int count = bs.cand_count;
      float wv = -INFINITY;
      int wi = 0x7fffffff;
      if (count <= kMaxCand) {
        for (int c = warp; c < count; c += nw) {
          const int i = bs.cand[c];
          const float e = exact_conv_at(raw, g.SP, H, i / W, i % W, r, w2d);
          argmax_combine(wv, wi, e, i);
        }
      } else {  // plateau: too many near-maximal pixels -> every warp scans its share of the plane
        for (int base = warp * 32; base < HW; base += nw * 32) {
          const int i = base + lane;
          unsigned m = __ballot_sync(0xffffffffu, i < HW && conv[i] >= thr);
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int q = base + b;
            const float e = exact_conv_at(raw, g.SP, H, q / W, q % W, r, w2d);
            argmax_combine(wv, wi, e, q);
          }
        }
      }
      if (lane == 0) { bs.fw[warp] = wv; bs.iv[warp] = wi; }
      __syncthreads();
      best_val = bs.fw[0]; best = bs.iv[0];
      for (int i = 1; i < nw; ++i) argmax_combine(best_val, best, bs.fw[i], bs.iv[i]);

      const int bx = best % W, by = best / W;
      interior = bx > 0 && bx < W - 1 && by > 0 && by < H - 1;  // heatmap.py:120-125
      if (interior) {
        for (int q = warp; q < 4; q += nw) {
          const int dx = (q == 0) ? -1 : (q == 1) ? 1 : 0;
          const int dy = (q == 2) ? -1 : (q == 3) ? 1 : 0;
          const float e = exact_conv_at(raw, g.SP, H, by + dy, bx + dx, r, w2d);
          if (lane == 0) bs.nb[q] = e;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) nb[q] = bs.nb[q];
      }
    }

    if (threadIdx.x == 0) {
      const int bx = best % W, by = best / W;
      float fx = static_cast<float>(bx), fy = static_cast<float>(by);
      if (interior) {  // _get_subpixel_maximums, float32, op order of heatmap.py:136-165
        const float l = nb[0], rr = nb[1], u = nb[2], dn = nb[3], c = best_val;
        const float gx = __fdiv_rn(__fsub_rn(rr, l), 2.0f);
        const float gy = __fdiv_rn(__fsub_rn(dn, u), 2.0f);
        float hxx = __fsub_rn(__fadd_rn(rr, l), __fmul_rn(2.0f, c));
        float hyy = __fsub_rn(__fadd_rn(dn, u), __fmul_rn(2.0f, c));
        if (hxx == 0.0f) hxx = 1e-6f;
        if (hyy == 0.0f) hyy = 1e-6f;
        fx = __fadd_rn(fx, __fdiv_rn(-gx, hxx));
        fy = __fadd_rn(fy, __fdiv_rn(-gy, hyy));
      }
      locs[hm * 2] = fx;
      locs[hm * 2 + 1] = fy;
      vals[hm] = raw[by * g.SP + kPadL + bx];  // the unconvolved map (heatmap.py:375-379)
      if (argmax) argmax[hm] = best;
      if (keypoints) {  // float32 / int -> float64, then * input_size (codec.py:237)
        keypoints[hm * 2] = static_cast<double>(fx) / static_cast<double>(W - 1) * p.input_w;
        keypoints[hm * 2 + 1] = static_cast<double>(fy) / static_cast<double>(H - 1) * p.input_h;
      }
    }
    __syncthreads();  // shared planes / scratch are reused by the next heatmap
  }
}

#include "pp_decode_fast.cuh"
#ifdef PP_EXPERIMENTS   // measured-and-rejected variants live in tools/experiments, outside the default build
#include "../../tools/experiments/pp_decode_dense.cuh"
#endif
#include "pp_decode_warp.cuh"

// ---------------------------------------------------------------------------
// generic exact path: full convolved map (return_heatmap=True, or maps too large for shared memory)
// one thread per pixel, sequential row-major double accumulation without contraction = scipy's order.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_exact_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps, float* __restrict__ conv_out) {
  const int H = p.H, W = p.W, HW = H * W;
  const int64_t total = static_cast<int64_t>(p.B) * p.K * HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t hm = i / HW;
    const int pix = static_cast<int>(i - hm * HW);
    const int y = pix / W, x = pix - y * W;
    const int k = static_cast<int>(hm % p.K);
    const int r = tab.radius[k], d = 2 * r + 1;
    const double* w2d = tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;
    const T* src = heatmaps + hm * HW;
    double acc = 0.0;
    for (int ti = 0; ti < d; ++ti) {
      const int yy = reflect_index(y + ti - r, H);
      for (int tj = 0; tj < d; ++tj) {
        const int xx = reflect_index(x + tj - r, W);
        const float v = apply_tail<T>(Elem<T>::to_f32(src[yy * W + xx]), p.apply_tail != 0, p.temperature);
        acc = __dadd_rn(acc, __dmul_rn(w2d[ti * d + tj], static_cast<double>(v)));
      }
    }
    conv_out[i] = static_cast<float>(acc);
  }
}

// argmax + sub-pixel + score from an already convolved map in global memory (any size)
template <typename T>
__global__ void __launch_bounds__(256)
argmax_from_conv_kernel(pp_decode_params p, const T* __restrict__ heatmaps, const float* __restrict__ conv,
                        float* __restrict__ locs, float* __restrict__ vals, int32_t* __restrict__ argmax,
                        double* __restrict__ keypoints) {
  __shared__ BlockScratch bs;
  const int H = p.H, W = p.W, HW = H * W;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const float* c = conv + hm * HW;
    float v = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) argmax_combine(v, idx, c[i], i);
    block_argmax(v, idx, bs);
    if (threadIdx.x == 0) {
      const int bx = idx % W, by = idx / W;
      float fx = static_cast<float>(bx), fy = static_cast<float>(by);
      if (bx > 0 && bx < W - 1 && by > 0 && by < H - 1) {
        const float l = c[idx - 1], rr = c[idx + 1], u = c[idx - W], dn = c[idx + W], cc = c[idx];
        const float gx = __fdiv_rn(__fsub_rn(rr, l), 2.0f);
        const float gy = __fdiv_rn(__fsub_rn(dn, u), 2.0f);
        float hxx = __fsub_rn(__fadd_rn(rr, l), __fmul_rn(2.0f, cc));
        float hyy = __fsub_rn(__fadd_rn(dn, u), __fmul_rn(2.0f, cc));
        if (hxx == 0.0f) hxx = 1e-6f;
        if (hyy == 0.0f) hyy = 1e-6f;
        fx = __fadd_rn(fx, __fdiv_rn(-gx, hxx));
        fy = __fadd_rn(fy, __fdiv_rn(-gy, hyy));
      }
      locs[hm * 2] = fx;
      locs[hm * 2 + 1] = fy;
      vals[hm] = apply_tail<T>(Elem<T>::to_f32(heatmaps[hm * HW + idx]), p.apply_tail != 0, p.temperature);
      if (argmax) argmax[hm] = idx;
      if (keypoints) {
        keypoints[hm * 2] = static_cast<double>(fx) / static_cast<double>(W - 1) * p.input_w;
        keypoints[hm * 2 + 1] = static_cast<double>(fy) / static_cast<double>(H - 1) * p.input_h;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// get_heatmap_maximum
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
heatmap_maximum_kernel(const T* __restrict__ heatmaps, int64_t N, int H, int W, float* __restrict__ locs,
                       float* __restrict__ vals, int32_t* __restrict__ argmax, bool vector_ok) {
  __shared__ BlockScratch bs;
  constexpr int V = Elem<T>::kVec;
  const int HW = H * W;
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const T* src = heatmaps + hm * HW;
    float v = -INFINITY;
    int idx = 0x7fffffff;
    if (vector_ok) {
      for (int i = threadIdx.x; i < HW / V; i += blockDim.x) {
        float f[V];
        unpack(ldg_stream_128(src + i * V), f, T());
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (f[j] > v) { v = f[j]; idx = i * V + j; }
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) argmax_combine(v, idx, Elem<T>::to_f32(src[i]), i);
    }
    block_argmax(v, idx, bs);
    if (threadIdx.x == 0) {
      const bool empty = !(v > 0.0f);  // locs[vals <= 0] = -1 (heatmap.py:46)
      locs[hm * 2] = empty ? -1.0f : static_cast<float>(idx % W);
      locs[hm * 2 + 1] = empty ? -1.0f : static_cast<float>(idx / W);
      vals[hm] = v;
      if (argmax) argmax[hm] = idx;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// argmax + DARK-UDP decoder
// ---------------------------------------------------------------------------
// 2x2 symmetric pseudo-inverse through the eigen-decomposition (== numpy.linalg.pinv for symmetric
// input, rcond = 1e-15), float64 (codec.py:371).
__device__ __forceinline__ void pinv_sym2(double a, double b, double c, double& ia, double& ib, double& ic) {
  const double half_tr = 0.5 * (a + c), half_df = 0.5 * (a - c);
  const double rad = sqrt(half_df * half_df + b * b);
  const double l1 = half_tr + rad, l2 = half_tr - rad;
  double vx, vy;  // unit eigenvector of l1
  if (rad == 0.0) { vx = 1.0; vy = 0.0; }
  else if (half_df >= 0.0) { vx = half_df + rad; vy = b; }
  else { vx = b; vy = rad - half_df; }
  const double nrm = sqrt(vx * vx + vy * vy);
  if (nrm > 0.0) { vx /= nrm; vy /= nrm; } else { vx = 1.0; vy = 0.0; }
  const double cutoff = 1e-15 * fmax(fabs(l1), fabs(l2));
  const double i1 = fabs(l1) > cutoff ? 1.0 / l1 : 0.0;
  const double i2 = fabs(l2) > cutoff ? 1.0 / l2 : 0.0;
  // A^+ = i1 v v^T + i2 w w^T with w = (-vy, vx)
  ia = i1 * vx * vx + i2 * vy * vy;
  ib = (i1 - i2) * vx * vy;
  ic = i1 * vy * vy + i2 * vx * vx;
}

#include "pp_dark_fast.cuh"
#include "pp_decode_mma.cuh"

template <typename T>
__global__ void __launch_bounds__(512)
decode_dark_kernel(pp_decode_params p, const float* __restrict__ blur_taps, int ksize, const T* __restrict__ heatmaps,
                   float* __restrict__ peaks, float* __restrict__ scores, float* __restrict__ refined,
                   double* __restrict__ keypoints, bool vector_ok) {
  extern __shared__ __align__(16) float smem[];
  __shared__ BlockScratch bs;
  __shared__ float taps[PP_MAX_BLUR_KSIZE];

  const int H = p.H, W = p.W, HW = H * W;
  const int r = ksize / 2;
  // the specialised (radius 5) column pass reads a zero-bordered tmp plane: r rows above, r + kTile below
  const int extra = (r == 5) ? (2 * r + kTile) : 0;
  const PlaneGeom g = plane_geom(H, W, extra);
  float* raw = smem;
  float* tmp = raw + g.raw_floats;
  float* blur = tmp + g.tmp_floats;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;

  for (int i = threadIdx.x; i < ksize; i += blockDim.x) taps[i] = blur_taps[i];
  for (int i = threadIdx.x; i < g.tmp_floats; i += blockDim.x) tmp[i] = 0.0f;  // zero border rows stay zero
  __syncthreads();

  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    LoadStats st = stage_plane<T>(heatmaps + hm * HW, raw, H, W, g.SP, vector_ok, p.apply_tail != 0, p.temperature);
    float top = st.vmax;
    int at = st.imax;
    block_argmax(top, at, bs);
    const bool empty = !(top > 0.0f);
    float fx = -1.0f, fy = -1.0f;

    if (!empty) {
      fill_pads(raw, H, W, g.SP, /*reflect=*/false);
      __syncthreads();
      separable_filter<false>(raw, tmp, blur, taps, r, H, W, g);
      float bmax = -INFINITY;
      for (int i = threadIdx.x; i < HW; i += blockDim.x) bmax = fmaxf(bmax, blur[i]);
      bmax = block_max(bmax, bs);

      if (threadIdx.x == 0) {
        const int px = at % W, py = at / W;
        // heatmaps[k] *= origin_max / (max(blurred) + 1e-12), float32 (codec.py:312)
        const float ratio = __fdiv_rn(top, __fadd_rn(bmax, 1e-12f));
        auto lg = [&](int yy, int xx) -> float {  // edge-padded, clipped, log (codec.py:343-347)
          yy = min(max(yy, 0), H - 1);
          xx = min(max(xx, 0), W - 1);
          float v = __fmul_rn(blur[yy * W + xx], ratio);
          v = fminf(fmaxf(v, 1e-3f), 50.0f);
          return static_cast<float>(log(static_cast<double>(v)));
        };
        const float c = lg(py, px), xp = lg(py, px + 1), xm = lg(py, px - 1), yp = lg(py + 1, px), ym = lg(py - 1, px);
        const float pp_ = lg(py + 1, px + 1), mm = lg(py - 1, px - 1);
        // float32 derivatives in the reference's operation order (codec.py:361-368)
        const float dx = __fmul_rn(0.5f, __fsub_rn(xp, xm));
        const float dy = __fmul_rn(0.5f, __fsub_rn(yp, ym));
        const float dxx = __fadd_rn(__fsub_rn(xp, __fmul_rn(2.0f, c)), xm);
        const float dyy = __fadd_rn(__fsub_rn(yp, __fmul_rn(2.0f, c)), ym);
        float t = __fsub_rn(pp_, xp);
        t = __fsub_rn(t, yp);
        t = __fadd_rn(t, c);
        t = __fadd_rn(t, c);
        t = __fsub_rn(t, xm);
        t = __fsub_rn(t, ym);
        t = __fadd_rn(t, mm);
        const float dxy = __fmul_rn(0.5f, t);
        const double eps = 1.1920928955078125e-07;  // np.finfo(np.float32).eps
        double ia, ib, ic;
        pinv_sym2(static_cast<double>(dxx) + eps, static_cast<double>(dxy), static_cast<double>(dyy) + eps, ia, ib, ic);
        const double sx = ia * static_cast<double>(dx) + ib * static_cast<double>(dy);
        const double sy = ib * static_cast<double>(dx) + ic * static_cast<double>(dy);
        // float32 keypoints -= float64 shift, stored back as float32 (codec.py:372-373)
        fx = static_cast<float>(static_cast<double>(static_cast<float>(px)) - sx);
        fy = static_cast<float>(static_cast<double>(static_cast<float>(py)) - sy);
      }
    }
    if (threadIdx.x == 0) {
      if (peaks) {
        peaks[hm * 2] = empty ? -1.0f : static_cast<float>(at % W);
        peaks[hm * 2 + 1] = empty ? -1.0f : static_cast<float>(at / W);
      }
      scores[hm] = top;
      refined[hm * 2] = fx;
      refined[hm * 2 + 1] = fy;
      if (keypoints) {
        keypoints[hm * 2] = static_cast<double>(fx) / static_cast<double>(W - 1) * p.input_w;
        keypoints[hm * 2 + 1] = static_cast<double>(fy) / static_cast<double>(H - 1) * p.input_h;
      }
    }
    __syncthreads();
  }
}

// Generic DARK path for maps that do not fit shared memory: everything from global memory (L1/L2).
// blurred(y, x) is evaluated per pixel as "row sums first, then the column combination", i.e. with the
// operation order of the separable passes, so the values are identical to the shared-memory kernels'.
template <typename T>
__device__ __forceinline__ float blur_at(const T* __restrict__ src, int H, int W, int y, int x, const float* taps,
                                         int r, bool tail, float temp) {
  float acc = 0.0f;
  for (int i = 0; i <= 2 * r; ++i) {
    const int yy = y + i - r;
    float rowsum = 0.0f;
    if (yy >= 0 && yy < H) {
      for (int j = 0; j <= 2 * r; ++j) {
        const int xx = x + j - r;
        const float v = (xx >= 0 && xx < W) ? apply_tail<T>(Elem<T>::to_f32(src[yy * W + xx]), tail, temp) : 0.0f;
        rowsum = fmaf(taps[j], v, rowsum);
      }
    }
    acc = fmaf(taps[i], rowsum, acc);
  }
  return acc;
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_dark_generic_kernel(pp_decode_params p, const float* __restrict__ blur_taps, int ksize,
                           const T* __restrict__ heatmaps, float* __restrict__ peaks, float* __restrict__ scores,
                           float* __restrict__ refined, double* __restrict__ keypoints) {
  __shared__ BlockScratch bs;
  __shared__ float taps[PP_MAX_BLUR_KSIZE];
  __shared__ float stencil[8];
  const int H = p.H, W = p.W, HW = H * W, r = ksize / 2;
  const bool tail = p.apply_tail != 0;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  for (int i = threadIdx.x; i < ksize; i += blockDim.x) taps[i] = blur_taps[i];
  __syncthreads();
  for (int64_t hm = blockIdx.x; hm < N; hm += gridDim.x) {
    const T* src = heatmaps + hm * HW;
    float top = -INFINITY;
    int at = 0x7fffffff;
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
      argmax_combine(top, at, apply_tail<T>(Elem<T>::to_f32(src[i]), tail, p.temperature), i);
    block_argmax(top, at, bs);
    const bool empty = !(top > 0.0f);
    float fx = -1.0f, fy = -1.0f;
    const int px = at % W, py = at / W;
    if (!empty) {
      float bmax = -INFINITY;
      for (int i = threadIdx.x; i < HW; i += blockDim.x)
        bmax = fmaxf(bmax, blur_at<T>(src, H, W, i / W, i % W, taps, r, tail, p.temperature));
      bmax = block_max(bmax, bs);
      if (threadIdx.x < 7) {
        const int dxs[7] = {0, 1, -1, 0, 0, 1, -1}, dys[7] = {0, 0, 0, 1, -1, 1, -1};
        const int yy = min(max(py + dys[threadIdx.x], 0), H - 1), xx = min(max(px + dxs[threadIdx.x], 0), W - 1);
        stencil[threadIdx.x] = blur_at<T>(src, H, W, yy, xx, taps, r, tail, p.temperature);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const float ratio = __fdiv_rn(top, __fadd_rn(bmax, 1e-12f));
        float b[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) b[q] = dark_log(stencil[q], ratio);
        dark_shift(b, px, py, fx, fy);
      }
    }
    if (threadIdx.x == 0) {
      if (peaks) {
        peaks[hm * 2] = empty ? -1.0f : static_cast<float>(px);
        peaks[hm * 2 + 1] = empty ? -1.0f : static_cast<float>(py);
      }
      scores[hm] = top;
      refined[hm * 2] = fx;
      refined[hm * 2 + 1] = fy;
      if (keypoints) {
        keypoints[hm * 2] = static_cast<double>(fx) / static_cast<double>(W - 1) * p.input_w;
        keypoints[hm * 2 + 1] = static_cast<double>(fy) / static_cast<double>(H - 1) * p.input_h;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// head tail, stand-alone
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
tail_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t numel, float temperature, bool vector_ok) {
  constexpr int V = Elem<T>::kVec;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nvec = vector_ok ? numel / V : 0;
  for (int64_t i = tid; i < nvec; i += stride) {
    float f[V];
    unpack(ldg_stream_128(x + i * V), f, T());
#pragma unroll
    for (int j = 0; j < V; ++j) f[j] = apply_tail<T>(f[j], true, temperature);
    stg_stream_128(y + i * V, pack(f, T()));
  }
  for (int64_t i = nvec * V + tid; i < numel; i += stride)
    y[i] = Elem<T>::from_f32(apply_tail<T>(Elem<T>::to_f32(x[i]), true, temperature));
}

template <typename T>
__global__ void __launch_bounds__(256)
tail_backward_kernel(const T* __restrict__ x, const T* __restrict__ gy, T* __restrict__ gx, int64_t numel,
                     float temperature, bool vector_ok) {
  constexpr int V = Elem<T>::kVec;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nvec = vector_ok ? numel / V : 0;
  auto one = [&](float xv, float g) -> float {
    // z is rounded to the tensor dtype like the forward; clamp passes the gradient on [0, 1] inclusive
    const float z = Elem<T>::to_f32(Elem<T>::from_f32(__fdiv_rn(xv, temperature)));
    return (z >= 0.0f && z <= 1.0f) ? Elem<T>::to_f32(Elem<T>::from_f32(__fdiv_rn(g, temperature))) : 0.0f;
  };
  for (int64_t i = tid; i < nvec; i += stride) {
    float a[V], b[V];
    unpack(ldg_stream_128(x + i * V), a, T());
    unpack(ldg_stream_128(gy + i * V), b, T());
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = one(a[j], b[j]);
    stg_stream_128(gx + i * V, pack(a, T()));
  }
  for (int64_t i = nvec * V + tid; i < numel; i += stride)
    gx[i] = Elem<T>::from_f32(one(Elem<T>::to_f32(x[i]), Elem<T>::to_f32(gy[i])));
}

int pick_threads(int H, int W) {
  // one register-tiled task = kTile outputs; aim for whole rounds of tasks per pass
  const int tasks = round_up(W, kTile) / kTile * H;
  int best = 128, best_waste = 1 << 30;
  for (int t = 128; t <= 512; t += 32) {
    const int rounds = (tasks + t - 1) / t;
    const int waste = rounds * t - tasks;
    if (waste < best_waste || (waste == best_waste && t <= 256)) { best = t; best_waste = waste; }
  }
  return best;
}

// Shared-memory layout of the persistent decoder kernels; false when the shape / alignment rules them out.
template <typename T>
bool fast_geometry(const pp_decode_params& p, const void* heatmaps, int max_radius, FastGeom* out, size_t* smem_bytes) {
  FastGeom geo{};
  geo.plane_bytes = static_cast<unsigned>(sizeof(T) * static_cast<size_t>(p.H) * p.W);
  geo.W8 = round_up(p.W, kTile);
  geo.full_stride = conflict_free_stride(kFMarg + geo.W8 + kFMarg);
  geo.work_off = (geo.plane_bytes + 127) / 128 * 128;
  // + 32 floats of slack: the last window of the last row reads a little past its row (zero taps only)
  geo.work_floats = static_cast<unsigned>(
      round_up(static_cast<int>(std::max<size_t>(kFTileFloats, static_cast<size_t>(p.H) * geo.full_stride)) + 32, 4));
  geo.taskmax_off = geo.work_off + static_cast<unsigned>(sizeof(float)) * geo.work_floats;
  // per-task maxima of the full path; the same bytes later hold kFThreads doubles of partial sums
  geo.w2d_off = geo.taskmax_off + static_cast<unsigned>(std::max<size_t>(sizeof(float) * round_up((geo.W8 / kTile) * p.H, 4),
                                                                         sizeof(double) * kFThreads));
  const size_t fsmem = geo.w2d_off + sizeof(double) * PP_OKS_TAPS * PP_OKS_TAPS;
  const bool ok = geo.plane_bytes % 16 == 0 && pp_aligned16(heatmaps) && p.W % Elem<T>::kVec == 0 &&
                  (!p.apply_tail || p.temperature > 0.0f) && fsmem + 4096 <= static_cast<size_t>(pp_smem_optin()) &&
                  p.W > max_radius + 1 && p.H > max_radius + 1 && static_cast<int64_t>(p.H) * p.W < (1 << 20);
  geo.div_WV = div_magic(static_cast<unsigned>(std::max(1, p.W / Elem<T>::kVec)));
  geo.div_W = div_magic(static_cast<unsigned>(p.W));
  geo.div_H = div_magic(static_cast<unsigned>(p.H));
  *out = geo;
  *smem_bytes = fsmem;
  return ok;
}

// Shared-memory slot of the warp-per-heatmap decoder (pp_decode_warp.cuh); false when the shape rules it out.
template <typename T>
bool warp_geometry(const pp_decode_params& p, const void* heatmaps, WarpGeom* out) {
  WarpGeom geo{};
  geo.plane_bytes = static_cast<unsigned>(sizeof(T) * static_cast<size_t>(p.H) * p.W);
  geo.tmp_off = (geo.plane_bytes + 127) / 128 * 128;
  geo.taps_off = geo.tmp_off + static_cast<unsigned>(sizeof(wf2)) * kWTmpPairs;
  geo.cand_off = geo.taps_off + 2u * static_cast<unsigned>(sizeof(wf2)) * kWTaps;
  geo.exch_off = (geo.cand_off + static_cast<unsigned>(sizeof(int)) * (2 * kWCand + 4) + 15) / 16 * 16;
  geo.slot_bytes = (geo.exch_off + static_cast<unsigned>(sizeof(TeamExchange)) + 127) / 128 * 128;
  constexpr int kMinSide = 2 * PP_MAX_OKS_RADIUS + 6;   // single reflections only, with the 4-row task overhang
  const bool ok = geo.plane_bytes % 16 == 0 && pp_aligned16(heatmaps) && p.W % Elem<T>::kVec == 0 &&
                  (!p.apply_tail || p.temperature > 0.0f) && p.W >= kMinSide && p.H >= kMinSide &&
                  2 * (round_up(p.W, 8) + kWTaps + 2) <= kWTmpPairs &&   // a band holds at least two row pairs
                  static_cast<int64_t>(p.H) * p.W < (1 << 20) && static_cast<int64_t>(p.B) * p.K < (1ll << 31);
  geo.div_WV = div_magic(static_cast<unsigned>(std::max(1, p.W / Elem<T>::kVec)));
  geo.div_W = div_magic(static_cast<unsigned>(p.W));
  *out = geo;
  return ok;
}

#ifdef PP_EXPERIMENTS
// Shared-memory layout of the dense decoder; false when the shape / alignment rules it out.
template <typename T>
bool dense_geometry(const pp_decode_params& p, const void* heatmaps, DenseGeom* out, size_t* smem_bytes) {
  DenseGeom geo{};
  geo.plane_bytes = static_cast<unsigned>(sizeof(T) * static_cast<size_t>(p.H) * p.W);
  geo.raw_stride = (geo.plane_bytes + 127) / 128 * 128;
  geo.Wp = round_up(p.W, 8);
  geo.Hp = round_up(p.H, 8);
  geo.PS = conflict_free_stride(kDMargin + geo.Wp + kDMargin);
  geo.QS = conflict_free_stride(geo.Wp);
  geo.p_off = 2 * geo.raw_stride;
  const size_t p_floats = static_cast<size_t>(geo.PS) * geo.Hp + 32;
  geo.q_off = static_cast<unsigned>((geo.p_off + sizeof(float) * p_floats + 127) / 128 * 128);
  geo.q_floats = static_cast<unsigned>(static_cast<size_t>(geo.QS) * (geo.Hp + 2 * kDMargin) + 32);
  geo.w2d_off = static_cast<unsigned>((geo.q_off + sizeof(float) * geo.q_floats + 127) / 128 * 128);
  geo.part_off = geo.w2d_off + 2u * static_cast<unsigned>(sizeof(double)) * PP_OKS_TAPS * PP_OKS_TAPS;
  const size_t smem = geo.part_off + sizeof(double) * kDThreads;
  geo.div_WV = div_magic(static_cast<unsigned>(std::max(1, p.W / Elem<T>::kVec)));
  geo.div_W = div_magic(static_cast<unsigned>(p.W));
  geo.div_H = div_magic(static_cast<unsigned>(p.H));
  geo.div_Wp = div_magic(static_cast<unsigned>(geo.Wp));
  geo.div_Wp4 = div_magic(static_cast<unsigned>(geo.Wp / 4));
  const bool ok = geo.plane_bytes % 16 == 0 && pp_aligned16(heatmaps) && p.W % Elem<T>::kVec == 0 &&
                  (!p.apply_tail || p.temperature > 0.0f) && 2 * (smem + 2048) <= static_cast<size_t>(pp_smem_optin()) &&
                  p.W > PP_MAX_OKS_RADIUS + 1 && p.H > PP_MAX_OKS_RADIUS + 1 && static_cast<int64_t>(p.H) * p.W < (1 << 20) &&
                  static_cast<int64_t>(p.B) * p.K < (1ll << 31);
  *out = geo;
  *smem_bytes = smem;
  return ok;
}
#endif

// Launch of the tensor-core kernel for the shapes it is instantiated for; PP_ERR_UNSUPPORTED_SHAPE otherwise.
template <typename T, int H, int W, int WPC, int MINB, bool kDark>
int mma_launch_shape(const pp_decode_params& p, const pp_oks_table& tab, const T* hm, float* locs, float* vals,
                     int32_t* argmax, double* keypoints, unsigned* scratch, float* dbg, const MmaDarkArgs& dk, cudaStream_t st) {
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  MmaGeom geo{};
  geo.plane_bytes = static_cast<unsigned>(sizeof(T) * H * W);
  geo.cand_off = (geo.plane_bytes + 15) / 16 * 16;
  geo.slot_bytes = (geo.cand_off + static_cast<unsigned>(sizeof(int)) * (2 * kWCand + 4) + 127) / 128 * 128;
  geo.div_W = div_magic(static_cast<unsigned>(W));
  geo.static_split = pp_env_int("PP_DECODE_STATIC", 0) ? 1u : 0u;
  const size_t smem = static_cast<size_t>(WPC) * geo.slot_bytes;
  auto kern = ((dbg || dk.dbg_times) && !kDark) ? decode_expected_mma_kernel<T, H, W, WPC, MINB, true, false>
                              : decode_expected_mma_kernel<T, H, W, WPC, MINB, false, kDark>;
  int per = 0;
  if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(kern), 32 * WPC, smem, &per)) return rc;
  if (const int cap = pp_env_int("PP_DECODE_CTAS", 0); cap > 0) per = std::min(per, cap);
  int grid = static_cast<int>(std::min<int64_t>((N + WPC - 1) / WPC, static_cast<int64_t>(pp_sm_count()) * per));
  if (const int cap = pp_env_int("PP_DECODE_GRID", 0); cap > 0) grid = std::min(grid, cap);   // test hook: many heatmaps per warp
  PP_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(unsigned) * kMmaScratchHead, st));
  if (pp_env_int("PP_DEBUG", 0))
    fprintf(stderr, "[pp] decode_expected_mma_kernel %dx%d dark=%d grid=%d threads=%d smem=%zu ctas/sm=%d\n", H, W, int(kDark), grid, 32 * WPC, smem, per);
  kern<<<grid, 32 * WPC, smem, st>>>(p, tab, hm, locs, vals, argmax, keypoints, geo, scratch, dbg, dk);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

template <typename T, bool kDark>
int mma_launch(const pp_decode_params& p, const pp_oks_table& tab, const T* hm, float* locs, float* vals, int32_t* argmax,
               double* keypoints, unsigned* scratch, float* dbg, const MmaDarkArgs& dk, cudaStream_t st) {
  if (!pp_aligned16(hm) || (p.apply_tail && !(p.temperature > 0.0f)) || p.K > kMmaMaxK) return PP_ERR_UNSUPPORTED_SHAPE;
  if (p.H == 64 && p.W == 48) {
    if (!kDark && pp_env_int("PP_DECODE_MINB", 4) == 3)   // experiment: 12 warps per SM with up to 168 registers (slower)
      return mma_launch_shape<T, 64, 48, 4, 3, false>(p, tab, hm, locs, vals, argmax, keypoints, scratch, dbg, dk, st);
    return mma_launch_shape<T, 64, 48, 4, 4, kDark>(p, tab, hm, locs, vals, argmax, keypoints, scratch, dbg, dk, st);
  }
  if (p.H == 96 && p.W == 72) return mma_launch_shape<T, 96, 72, 4, 2, kDark>(p, tab, hm, locs, vals, argmax, keypoints, scratch, dbg, dk, st);
  return PP_ERR_UNSUPPORTED_SHAPE;
}

thread_local int g_last_expected_kernel = -1;   // PP_DECODE_KERNEL_* of this thread's most recent pp_decode_expected

template <typename T>
int launch_decode_expected(const pp_decode_params& p, const pp_oks_table& tab, const void* heatmaps, float* locs,
                           float* vals, int32_t* argmax, double* keypoints, float* conv_out, void* scratch,
                           int64_t scratch_bytes, cudaStream_t st) {
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const T* hm = static_cast<const T*>(heatmaps);
  const PlaneGeom g = plane_geom(p.H, p.W, 0);
  const size_t smem = sizeof(float) * (static_cast<size_t>(g.raw_floats) + g.tmp_floats + g.out_floats);
  const bool fits = smem <= static_cast<size_t>(pp_smem_optin());
  if (conv_out != nullptr || !fits) {
    g_last_expected_kernel = PP_DECODE_KERNEL_EXACT;
    // exact full-map path; needs a convolved-map buffer
    PP_REQUIRE(conv_out != nullptr, PP_ERR_SCRATCH,
               "pp_decode_expected: %dx%d maps do not fit in shared memory; pass conv_out as work space", p.H, p.W);
    const int64_t total = N * p.H * p.W;
    const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(pp_sm_count()) * 16));
    conv_exact_kernel<T><<<grid, 256, 0, st>>>(p, tab, hm, conv_out);
    PP_CUDA_OK(cudaGetLastError());
    const int grid2 = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * 8));
    argmax_from_conv_kernel<T><<<grid2, 256, 0, st>>>(p, hm, conv_out, locs, vals, argmax, keypoints);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  }
  const bool vec = (p.W % Elem<T>::kVec == 0) && pp_aligned16(heatmaps);
  const int threads = pick_threads(p.H, p.W);

  // team-per-heatmap kernel (pp_decode_warp.cuh): G warps own a heatmap from its bulk copy to its outputs.  With
  // G = 1 it has a better throughput than the CTA-per-heatmap kernel below but the longest single-heatmap latency, so
  // among the two it takes over once every team has at least two heatmaps to work on (measured cross-over between
  // 2176 and 4352 heatmaps of 64x48, profiles/r01s_summary.md).  PP_DECODE_WARP=0 / 1 forces the choice,
  // PP_DECODE_TEAM selects G (1 or 2).  It is also the kernel that decodes the heatmaps the tensor-core kernel hands
  // on (`list`).
  WarpGeom wgeo{};
  const bool warp_ok = warp_geometry<T>(p, heatmaps, &wgeo);
  constexpr int kTPC = 2;
  int wG = 1, wper = 0;
  size_t wsmem = 0;
  const void* wfn = nullptr;
  if (warp_ok) {
    wgeo.rowoff_off = static_cast<unsigned>(kTPC) * wgeo.slot_bytes;
    wgeo.full_taps = pp_env_int("PP_DECODE_FULLTAPS", 0) ? 1u : 0u;
    wsmem = wgeo.rowoff_off + sizeof(int) * static_cast<size_t>(p.H + 2 * kWRowPad);
    // warps per heatmap: one while a dozen heatmaps fit on an SM (64x48 float32: 12); larger planes leave room for
    // fewer heatmaps, and two warps each then keep the SM's schedulers fed (96x72 float32, 6 heatmaps per SM:
    // 202 us with two warps, 237 us with one, 232 us for the CTA-per-heatmap kernel; profiles/r01s_summary.md)
    wG = pp_env_int("PP_DECODE_TEAM", 0);
    if (wG != 1 && wG != 2) {
      const int64_t slots = (static_cast<int64_t>(pp_smem_optin()) + 1024) / static_cast<int64_t>(wsmem + 1024) * kTPC;
      wG = slots >= 10 ? 1 : 2;
    }
    wfn = wG == 1 ? reinterpret_cast<const void*>(decode_expected_warp_kernel<T, 1, kTPC>)
                  : reinterpret_cast<const void*>(decode_expected_warp_kernel<T, 2, kTPC>);
  }
  const bool warp_fits = warp_ok && wsmem + 2048 <= static_cast<size_t>(pp_smem_optin()) &&
                         pp_configure_kernel(wfn, 32 * wG * kTPC, wsmem, &wper) == PP_OK && wper * kTPC >= 6;
  auto launch_warp = [&](int grid, unsigned* counter, const int* list, const unsigned* list_count) -> int {
    if (pp_env_int("PP_DEBUG", 0))
      fprintf(stderr, "[pp] decode_expected_warp_kernel G=%d grid=%d threads=%d smem=%zu ctas/sm=%d list=%d\n", wG, grid,
              32 * wG * kTPC, wsmem, wper, list != nullptr);
    if (wG == 1)
      decode_expected_warp_kernel<T, 1, kTPC><<<grid, 32 * wG * kTPC, wsmem, st>>>(p, tab, hm, locs, vals, argmax, keypoints, wgeo, counter, list, list_count);
    else
      decode_expected_warp_kernel<T, 2, kTPC><<<grid, 32 * wG * kTPC, wsmem, st>>>(p, tab, hm, locs, vals, argmax, keypoints, wgeo, counter, list, list_count);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  };

  // tensor-core kernel (pp_decode_mma.cuh): one warp per heatmap, the whole separable prefilter as two banded
  // Toeplitz GEMMs on mma.sync; needs the per-channel operand tables (pp_oks_mma_table_build) for exactly this
  // shape, scratch for the hand-over list and the team kernel above for the heatmaps on that list.
  // PP_DECODE_MMA=0 disables it; PP_DECODE_WARP=0 / 1 (a forced choice among the general kernels) does too.
  const int want_warp = pp_env_int("PP_DECODE_WARP", -1);
  if (warp_fits && want_warp < 0 && pp_env_int("PP_DECODE_MMA", 1) && tab.mma_tables && tab.mma_index &&
      tab.mma_H == p.H && tab.mma_W == p.W && pp_aligned16(tab.mma_tables) && scratch &&
      scratch_bytes >= static_cast<int64_t>(sizeof(unsigned)) * (kMmaScratchHead + N) && N < (1ll << 31)) {
    int rc = mma_launch<T, false>(p, tab, hm, locs, vals, argmax, keypoints, static_cast<unsigned*>(scratch), nullptr, MmaDarkArgs{}, st);
    if (rc == PP_OK) {
      // second, usually empty, launch: the heatmaps the tensor-core kernel could not rank (plateaus, no dynamic range)
      unsigned* words = static_cast<unsigned*>(scratch);
      const int cap = pp_env_int("PP_DECODE_RETRY_GRID", 0);
      const int rgrid = static_cast<int>(std::min<int64_t>((N + kTPC - 1) / kTPC, cap > 0 ? cap : pp_sm_count()));
      g_last_expected_kernel = PP_DECODE_KERNEL_MMA;
      return launch_warp(rgrid, words + 1, reinterpret_cast<const int*>(words + kMmaScratchHead), words + 2);
    }
    if (rc != PP_ERR_UNSUPPORTED_SHAPE) return rc;
  }

  if (want_warp != 0 && warp_fits && (want_warp == 1 || N >= 2ll * pp_sm_count() * wper * kTPC)) {
    int per = wper;
    if (const int cap = pp_env_int("PP_DECODE_CTAS", 0); cap > 0) per = std::min(per, cap);
    int wgrid = static_cast<int>(std::min<int64_t>((N + kTPC - 1) / kTPC, static_cast<int64_t>(pp_sm_count()) * per));
    if (const int cap = pp_env_int("PP_DECODE_GRID", 0); cap > 0) wgrid = std::min(wgrid, cap);   // test hook: many heatmaps per team
    unsigned* counter = (scratch && scratch_bytes >= 4) ? static_cast<unsigned*>(scratch) : nullptr;
    if (counter) PP_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
    g_last_expected_kernel = PP_DECODE_KERNEL_TEAM;
    return launch_warp(wgrid, counter, nullptr, nullptr);
  }

#ifdef PP_EXPERIMENTS
  // dense kernel (pp_decode_dense.cuh): the whole separable prefilter, specialised per radius; at least two CTAs
  // per SM must fit, otherwise the pruned kernel below takes over
  DenseGeom dgeo{};
  size_t dsmem = 0;
  if (pp_env_int("PP_DECODE_DENSE", 0) && dense_geometry<T>(p, heatmaps, &dgeo, &dsmem)) {
    int dper = 1;
    if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(decode_expected_dense_kernel<T>), kDThreads, dsmem, &dper))
      return rc;
    if (const int cap = pp_env_int("PP_DECODE_CTAS", 0); cap > 0) dper = std::min(dper, cap);
    const int dgrid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * dper));
    unsigned* counter = (scratch && scratch_bytes >= 4 && N < (1ll << 31)) ? static_cast<unsigned*>(scratch) : nullptr;
    if (counter) PP_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
    g_last_expected_kernel = PP_DECODE_KERNEL_DENSE;
    decode_expected_dense_kernel<T><<<dgrid, kDThreads, dsmem, st>>>(p, tab, hm, locs, vals, argmax, keypoints, dgeo, counter);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  }
#endif

  // pruned kernel: TMA-staged plane, convolution pruned to the neighbourhood of {h >= L} when possible
  FastGeom geo{};
  size_t fsmem = 0;
  const bool fast = fast_geometry<T>(p, heatmaps, PP_MAX_OKS_RADIUS, &geo, &fsmem);
  if (fast) {
    int fper = 1;
    if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(decode_expected_fast_kernel<T>), kFThreads, fsmem, &fper))
      return rc;
    if (const int cap = pp_env_int("PP_DECODE_CTAS", 0); cap > 0) fper = std::min(fper, cap);
    int fgrid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * fper));
    if (const int cap = pp_env_int("PP_DECODE_GRID", 0); cap > 0) fgrid = std::min(fgrid, cap);   // test hook: many heatmaps per CTA
    unsigned* counter = (scratch && scratch_bytes >= 4 && N < (1ll << 31)) ? static_cast<unsigned*>(scratch) : nullptr;
    if (counter) PP_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
    g_last_expected_kernel = PP_DECODE_KERNEL_CTA;
    decode_expected_fast_kernel<T><<<fgrid, kFThreads, fsmem, st>>>(p, tab, hm, locs, vals, argmax, keypoints, geo, counter);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  }
  int per_sm = 1;
  if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(decode_expected_kernel<T>), threads, smem, &per_sm)) return rc;
  const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * per_sm));
  g_last_expected_kernel = PP_DECODE_KERNEL_GENERIC;
  decode_expected_kernel<T><<<grid, threads, smem, st>>>(p, tab, hm, locs, vals, argmax, keypoints, vec, false);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

thread_local int g_last_dark_kernel = -1;   // 1: CTA-per-heatmap kernel, 5: tensor-core kernel, 0: the other fallbacks

template <typename T>
int launch_decode_dark(const pp_decode_params& p, const float* taps, int ksize, const void* blur_mma_table,
                       const void* heatmaps, float* peaks, float* scores, float* refined, double* keypoints, void* scratch,
                       int64_t scratch_bytes, cudaStream_t st) {
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const int r = ksize / 2;
  {
    FastGeom geo{};
    size_t fsmem = 0;
    if (r <= PP_MAX_OKS_RADIUS && fast_geometry<T>(p, heatmaps, r, &geo, &fsmem)) {
      int fper = 1;
      if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(decode_dark_fast_kernel<T>), kFThreads, fsmem, &fper))
        return rc;
      // tensor-core kernel (pp_decode_mma.cuh, kDark): needs the blur's operand table for this shape, scratch for the
      // hand-over list and the kernel below for the heatmaps on that list.  PP_DARK_MMA=0 disables it.
      if (blur_mma_table && pp_aligned16(blur_mma_table) && ksize <= 16 && pp_env_int("PP_DARK_MMA", 1) && scratch &&
          scratch_bytes >= static_cast<int64_t>(sizeof(unsigned)) * (kMmaScratchHead + N) && N < (1ll << 31)) {
        pp_oks_table tab{};
        tab.mma_tables = blur_mma_table;
        MmaDarkArgs dk{taps, ksize, peaks};
        const int rc = mma_launch<T, true>(p, tab, static_cast<const T*>(heatmaps), refined, scores, nullptr, keypoints,
                                           static_cast<unsigned*>(scratch), nullptr, dk, st);
        if (rc == PP_OK) {
          unsigned* words = static_cast<unsigned*>(scratch);
          const int cap = pp_env_int("PP_DECODE_RETRY_GRID", 0);
          const int rgrid = static_cast<int>(std::min<int64_t>(N, cap > 0 ? cap : pp_sm_count()));
          g_last_dark_kernel = 5;
          decode_dark_fast_kernel<T><<<rgrid, kFThreads, fsmem, st>>>(p, taps, ksize, static_cast<const T*>(heatmaps), peaks, scores,
                                                                    refined, keypoints, geo, words + 1,
                                                                    reinterpret_cast<const int*>(words + kMmaScratchHead), words + 2);
          PP_CUDA_OK(cudaGetLastError());
          return PP_OK;
        }
        if (rc != PP_ERR_UNSUPPORTED_SHAPE) return rc;
      }
      int fgrid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * fper));
      if (const int cap = pp_env_int("PP_DARK_GRID", 0); cap > 0) fgrid = std::min(fgrid, cap);   // test hook: many heatmaps per CTA
      unsigned* counter = (scratch && scratch_bytes >= 4 && N < (1ll << 31)) ? static_cast<unsigned*>(scratch) : nullptr;
      if (counter) PP_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
      g_last_dark_kernel = 1;
      decode_dark_fast_kernel<T><<<fgrid, kFThreads, fsmem, st>>>(p, taps, ksize, static_cast<const T*>(heatmaps), peaks,
                                                                scores, refined, keypoints, geo, counter, nullptr, nullptr);
      PP_CUDA_OK(cudaGetLastError());
      return PP_OK;
    }
  }
  g_last_dark_kernel = 0;
  const PlaneGeom g = plane_geom(p.H, p.W, r == 5 ? 2 * r + kTile : 0);
  const size_t smem = sizeof(float) * (static_cast<size_t>(g.raw_floats) + g.tmp_floats + g.out_floats);
  if (smem + 4096 > static_cast<size_t>(pp_smem_optin())) {   // too large for shared memory: global-memory path
    const int ggrid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * 8));
    decode_dark_generic_kernel<T><<<ggrid, 256, 0, st>>>(p, taps, ksize, static_cast<const T*>(heatmaps), peaks, scores,
                                                         refined, keypoints);
    PP_CUDA_OK(cudaGetLastError());
    return PP_OK;
  }
  const bool vec = (p.W % Elem<T>::kVec == 0) && pp_aligned16(heatmaps);
  const int threads = pick_threads(p.H, p.W);
  int per_sm = 1;
  if (int rc = pp_configure_kernel(reinterpret_cast<const void*>(decode_dark_kernel<T>), threads, smem, &per_sm)) return rc;
  const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * per_sm));
  decode_dark_kernel<T><<<grid, threads, smem, st>>>(p, taps, ksize, static_cast<const T*>(heatmaps), peaks, scores,
                                                     refined, keypoints, vec);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int check_decode_params(const char* fn, const pp_decode_params* p) {
  PP_REQUIRE(p != nullptr, PP_ERR_INVALID_ARG, "%s: null params", fn);
  PP_REQUIRE(p->B >= 0 && p->K > 0 && p->H > 1 && p->W > 1, PP_ERR_INVALID_ARG, "%s: bad shape B=%d K=%d H=%d W=%d", fn,
             p->B, p->K, p->H, p->W);
  PP_REQUIRE(p->heatmap_dtype == PP_F32 || p->heatmap_dtype == PP_BF16, PP_ERR_INVALID_ARG, "%s: unsupported dtype %d",
             fn, p->heatmap_dtype);
  PP_REQUIRE(!p->apply_tail || p->temperature != 0.0f, PP_ERR_INVALID_ARG, "%s: zero temperature", fn);
  PP_REQUIRE(static_cast<int64_t>(p->H) * p->W < (1 << 30), PP_ERR_UNSUPPORTED_SHAPE, "%s: heatmap too large", fn);
  return PP_OK;
}

}  // namespace

#ifdef PP_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int pp_debug_phase_cycles(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16) != cudaSuccess) return -3;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" {

int64_t pp_decode_expected_scratch_bytes(void) { return 16; }

int64_t pp_decode_expected_scratch_bytes_for(const pp_decode_params* p) {
  if (!p || p->B < 0 || p->K < 0) return 16;
  return static_cast<int64_t>(sizeof(unsigned)) * (kMmaScratchHead + static_cast<int64_t>(p->B) * p->K);
}

int64_t pp_oks_mma_table_bytes(int32_t U, int32_t H, int32_t W) {
  if (U <= 0) return 0;
  if (H == 64 && W == 48) return static_cast<int64_t>(U) * (MmaShape<64, 48>::kT1 + MmaShape<64, 48>::kT2) * 16;
  if (H == 96 && W == 72) return static_cast<int64_t>(U) * (MmaShape<96, 72>::kT1 + MmaShape<96, 72>::kT2) * 16;
  return 0;
}

int pp_oks_mma_table_build(const float* taps_f32, const int32_t* radius, int32_t U, int32_t H, int32_t W, void* out,
                           pp_stream_t stream) {
  PP_REQUIRE(taps_f32 && radius && out && U > 0, PP_ERR_INVALID_ARG, "pp_oks_mma_table_build: null argument");
  PP_REQUIRE(pp_oks_mma_table_bytes(U, H, W) > 0, PP_ERR_UNSUPPORTED_SHAPE,
             "pp_oks_mma_table_build: no tensor-core decoder for %dx%d maps", H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = std::min(U * 8, 1024);
  if (H == 64) build_mma_tables_kernel<64, 48><<<grid, 256, 0, st>>>(taps_f32, radius, U, static_cast<__half*>(out), PP_OKS_TAPS, 0, false);
  else build_mma_tables_kernel<96, 72><<<grid, 256, 0, st>>>(taps_f32, radius, U, static_cast<__half*>(out), PP_OKS_TAPS, 0, false);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_decode_expected(const pp_decode_params* p, const pp_oks_table* table, const void* heatmaps, float* locs,
                       float* vals, int32_t* argmax, double* keypoints, float* conv_out, void* scratch,
                       int64_t scratch_bytes, pp_stream_t stream) {
  if (int rc = check_decode_params("pp_decode_expected", p)) return rc;
  if (p->B == 0) return PP_OK;
  PP_REQUIRE(table && table->radius && table->taps_f32 && table->kernel2d && heatmaps && locs && vals,
             PP_ERR_INVALID_ARG, "pp_decode_expected: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->heatmap_dtype == PP_F32)
    return launch_decode_expected<float>(*p, *table, heatmaps, locs, vals, argmax, keypoints, conv_out, scratch, scratch_bytes, st);
  return launch_decode_expected<__nv_bfloat16>(*p, *table, heatmaps, locs, vals, argmax, keypoints, conv_out, scratch, scratch_bytes, st);
}

int pp_decode_expected_last_kernel(void) { return g_last_expected_kernel; }

// Test hook (not part of the ABI header): runs the tensor-core kernel and additionally writes its float16 / float32
// proposal values, mapped back to the heatmap's units, into `prefilter` (N, H, W) -- tests/test_gpu_parity.py checks
// the rigorous error bound kMmaErr against the oracle's exact convolution with it.  Heatmaps that the kernel hands on
// keep their `prefilter` plane untouched and their outputs unwritten.
__attribute__((visibility("default"))) int pp_debug_decode_mma_prefilter(const pp_decode_params* p, const pp_oks_table* table,
                                                                         const void* heatmaps, float* locs, float* vals,
                                                                         int32_t* argmax, float* prefilter, void* scratch,
                                                                         int64_t scratch_bytes, pp_stream_t stream) {
  if (int rc = check_decode_params("pp_debug_decode_mma_prefilter", p)) return rc;
  PP_REQUIRE(table && table->mma_tables && table->mma_index && heatmaps && locs && vals && prefilter && scratch &&
                 scratch_bytes >= pp_decode_expected_scratch_bytes_for(p),
             PP_ERR_INVALID_ARG, "pp_debug_decode_mma_prefilter: null argument / scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->heatmap_dtype == PP_F32)
    return mma_launch<float, false>(*p, *table, static_cast<const float*>(heatmaps), locs, vals, argmax, nullptr,
                                    static_cast<unsigned*>(scratch), prefilter, MmaDarkArgs{}, st);
  return mma_launch<__nv_bfloat16, false>(*p, *table, static_cast<const __nv_bfloat16*>(heatmaps), locs, vals, argmax, nullptr,
                                          static_cast<unsigned*>(scratch), prefilter, MmaDarkArgs{}, st);
}

// Debug entry point (not in the header): the tensor-core kernel with a per-heatmap timeline -- `times` (N, 8) uint64, see
// the kernel; tools/decode_timeline.py turns it into phase latencies and a picture of when the warps finish.
__attribute__((visibility("default"))) int pp_debug_decode_mma_timeline(const pp_decode_params* p, const pp_oks_table* table,
                                                                        const void* heatmaps, float* locs, float* vals,
                                                                        int32_t* argmax, unsigned long long* times, void* scratch,
                                                                        int64_t scratch_bytes, pp_stream_t stream) {
  if (int rc = check_decode_params("pp_debug_decode_mma_timeline", p)) return rc;
  PP_REQUIRE(table && table->mma_tables && table->mma_index && heatmaps && locs && vals && times && scratch &&
                 scratch_bytes >= pp_decode_expected_scratch_bytes_for(p) && p->heatmap_dtype == PP_F32,
             PP_ERR_INVALID_ARG, "pp_debug_decode_mma_timeline: null argument / scratch too small / not float32");
  MmaDarkArgs dk{};
  dk.dbg_times = times;
  return mma_launch<float, false>(*p, *table, static_cast<const float*>(heatmaps), locs, vals, argmax, nullptr,
                                  static_cast<unsigned*>(scratch), nullptr, dk, static_cast<cudaStream_t>(stream));
}

int64_t pp_decode_expected_workspace_floats(const pp_decode_params* p) {
  if (!p || p->H < 1 || p->W < 1) return 0;
  const PlaneGeom g = plane_geom(p->H, p->W, 0);
  const size_t smem = sizeof(float) * (static_cast<size_t>(g.raw_floats) + g.tmp_floats + g.out_floats);
  if (smem <= static_cast<size_t>(pp_smem_optin())) return 0;
  return static_cast<int64_t>(p->B) * p->K * p->H * p->W;
}

int pp_heatmap_maximum(const void* heatmaps, int32_t heatmap_dtype, int64_t N, int32_t H, int32_t W, float* locs,
                       float* vals, int32_t* argmax, pp_stream_t stream) {
  PP_REQUIRE(N >= 0 && H > 0 && W > 0 && static_cast<int64_t>(H) * W < (1 << 30), PP_ERR_INVALID_ARG,
             "pp_heatmap_maximum: bad shape N=%lld H=%d W=%d", static_cast<long long>(N), H, W);
  if (N == 0) return PP_OK;
  PP_REQUIRE(heatmaps && locs && vals, PP_ERR_INVALID_ARG, "pp_heatmap_maximum: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(std::min<int64_t>(N, static_cast<int64_t>(pp_sm_count()) * 16));
  if (heatmap_dtype == PP_F32) {
    const bool vec = (static_cast<int64_t>(H) * W % 4 == 0) && pp_aligned16(heatmaps);
    heatmap_maximum_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(heatmaps), N, H, W, locs, vals, argmax, vec);
  } else if (heatmap_dtype == PP_BF16) {
    const bool vec = (static_cast<int64_t>(H) * W % 8 == 0) && pp_aligned16(heatmaps);
    heatmap_maximum_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(heatmaps), N, H, W, locs, vals, argmax, vec);
  } else {
    pp_set_error("pp_heatmap_maximum: unsupported dtype %d", heatmap_dtype);
    return PP_ERR_INVALID_ARG;
  }
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_blur_mma_table_build(const float* blur_taps, int32_t blur_ksize, int32_t H, int32_t W, void* out, pp_stream_t stream) {
  PP_REQUIRE(blur_taps && out && blur_ksize % 2 == 1 && blur_ksize >= 3 && blur_ksize <= 16, PP_ERR_INVALID_ARG,
             "pp_blur_mma_table_build: bad argument (odd kernel size in [3, 15])");
  PP_REQUIRE(pp_oks_mma_table_bytes(1, H, W) > 0, PP_ERR_UNSUPPORTED_SHAPE,
             "pp_blur_mma_table_build: no tensor-core decoder for %dx%d maps", H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (H == 64) build_mma_tables_kernel<64, 48><<<8, 256, 0, st>>>(blur_taps, nullptr, 1, static_cast<__half*>(out), blur_ksize, blur_ksize / 2, true);
  else build_mma_tables_kernel<96, 72><<<8, 256, 0, st>>>(blur_taps, nullptr, 1, static_cast<__half*>(out), blur_ksize, blur_ksize / 2, true);
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_decode_argmax_dark_last_kernel(void) { return g_last_dark_kernel; }

int pp_decode_argmax_dark(const pp_decode_params* p, const float* blur_taps, int32_t blur_ksize, const void* blur_mma_table,
                          const void* heatmaps, float* peaks, float* scores, float* refined, double* keypoints, void* scratch,
                          int64_t scratch_bytes, pp_stream_t stream) {
  if (int rc = check_decode_params("pp_decode_argmax_dark", p)) return rc;
  PP_REQUIRE(blur_ksize % 2 == 1 && blur_ksize >= 3 && blur_ksize <= PP_MAX_BLUR_KSIZE, PP_ERR_INVALID_ARG,
             "pp_decode_argmax_dark: blur kernel size %d must be odd and in [3, %d]", blur_ksize, PP_MAX_BLUR_KSIZE);
  if (p->B == 0) return PP_OK;
  PP_REQUIRE(blur_taps && heatmaps && scores && refined, PP_ERR_INVALID_ARG, "pp_decode_argmax_dark: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->heatmap_dtype == PP_F32)
    return launch_decode_dark<float>(*p, blur_taps, blur_ksize, blur_mma_table, heatmaps, peaks, scores, refined, keypoints, scratch, scratch_bytes, st);
  return launch_decode_dark<__nv_bfloat16>(*p, blur_taps, blur_ksize, blur_mma_table, heatmaps, peaks, scores, refined, keypoints, scratch, scratch_bytes, st);
}

int pp_heatmap_tail(const void* x, void* y, int32_t dtype, int64_t numel, float temperature, pp_stream_t stream) {
  PP_REQUIRE(numel >= 0 && temperature != 0.0f, PP_ERR_INVALID_ARG, "pp_heatmap_tail: bad numel/temperature");
  if (numel == 0) return PP_OK;
  PP_REQUIRE(x && y, PP_ERR_INVALID_ARG, "pp_heatmap_tail: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = pp_aligned16(x) && pp_aligned16(y);
  const int grid = static_cast<int>(std::min<int64_t>((numel + 2047) / 2048, static_cast<int64_t>(pp_sm_count()) * 16));
  if (dtype == PP_F32)
    tail_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), numel, temperature, vec);
  else if (dtype == PP_BF16)
    tail_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), numel, temperature, vec);
  else {
    pp_set_error("pp_heatmap_tail: unsupported dtype %d", dtype);
    return PP_ERR_INVALID_ARG;
  }
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

int pp_heatmap_tail_backward(const void* x, const void* grad_y, void* grad_x, int32_t dtype, int64_t numel,
                             float temperature, pp_stream_t stream) {
  PP_REQUIRE(numel >= 0 && temperature != 0.0f, PP_ERR_INVALID_ARG, "pp_heatmap_tail_backward: bad numel/temperature");
  if (numel == 0) return PP_OK;
  PP_REQUIRE(x && grad_y && grad_x, PP_ERR_INVALID_ARG, "pp_heatmap_tail_backward: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = pp_aligned16(x) && pp_aligned16(grad_y) && pp_aligned16(grad_x);
  const int grid = static_cast<int>(std::min<int64_t>((numel + 2047) / 2048, static_cast<int64_t>(pp_sm_count()) * 16));
  if (dtype == PP_F32)
    tail_backward_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<const float*>(grad_y),
                                                      static_cast<float*>(grad_x), numel, temperature, vec);
  else if (dtype == PP_BF16)
    tail_backward_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                              static_cast<const __nv_bfloat16*>(grad_y),
                                                              static_cast<__nv_bfloat16*>(grad_x), numel, temperature, vec);
  else {
    pp_set_error("pp_heatmap_tail_backward: unsupported dtype %d", dtype);
    return PP_ERR_INVALID_ARG;
  }
  PP_CUDA_OK(cudaGetLastError());
  return PP_OK;
}

}  // extern "C"

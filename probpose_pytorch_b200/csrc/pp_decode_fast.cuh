// Fast path of the expected-OKS decoder (included into pp_decode.cu, inside its anonymous namespace).
//
// The full separable convolution costs ~2 x 13 FFMA per pixel on COCO sigmas -- more than the ~2100
// warp instructions per 64x48 heatmap that the HBM roofline allows -- so this kernel prunes it, exactly:
//
//   the OKS kernel is non-negative and sums to 1, hence for every pixel p
//        R(p) = float32(sum_q w(q) h(p+q))  <=  max over the kernel window of h,
//   and  max_p R(p) >= R(p0) =: L  for any pixel p0 (we take the raw maximum).  Therefore the argmax of
//   R -- and every pixel tied with it -- lies within `radius` of the set S = {h >= L}.  For blob-shaped
//   heatmaps S is a handful of pixels around the peak.
//
// Per heatmap: (A) one TMA bulk copy of the contiguous plane into shared memory (prefetched during the
// previous heatmap), one scan for max/min/argmax; (B) L = exact R(p0) (double accumulation of the
// reference's d x d table); (C) bounding box of S; (D) gather the box dilated by the radius (+ reflect
// halo) into a small padded tile; (E/F) separable float32 prefilter on the tile only; (G) exact
// re-evaluation of the near-maximal pixels and of the four neighbours of the winner; (H) sub-pixel fit.
// Heatmaps whose box is larger than the tile (flat / multi-modal / noisy maps) are marked by writing
// NaN into locs[2*hm] and are finished by the full-plane kernel in a second launch.
#pragma once

constexpr int kFThreads = 128;
constexpr int kFRegion = 24;                    // largest output region side handled here
constexpr int kFMarg = 12;                      // tile column margin: >= radius + 1, multiple of 4
constexpr int kFTileW = kFMarg + kFRegion + kFMarg;             // 48
constexpr int kFTileStride = 52;                // multiple of 4 with odd quarter (conflict-free 128-bit rows)
constexpr int kFTileRows = kFRegion + 2 * (PP_MAX_OKS_RADIUS + 1);   // 44
constexpr int kFTmpRows = kFRegion + 2 * PP_MAX_OKS_RADIUS;          // 42
constexpr int kFTmpStride = 28;                 // 24 -> multiple of 4 with odd quarter
constexpr int kFT = 4;                          // outputs per task

struct FastShared {
  float red_f[2][4];
  int red_i[4];
  float taps[PP_OKS_TAPS];
  float nb[4];
  float L;
  int bbox[4];        // min x, max x, min y, max y of S
  int cand[kMaxCand];
  int cand_count;
};

template <typename T>
__device__ __forceinline__ float plane_value(const T* plane, int idx, bool tail, float temperature) {
  return apply_tail<T>(Elem<T>::to_f32(plane[idx]), tail, temperature);
}

// row pass on the gathered tile: tmp[ty][x] = sum_j tap[j] * tile[ty + 1][kFMarg + x - R + j]
template <int R>
__device__ __forceinline__ void fast_row_pass(const float* __restrict__ tile, float* __restrict__ tmp,
                                              const float* __restrict__ taps, int rows, int xblocks) {
  constexpr int PADR = (R + 3) & ~3;
  constexpr int WIN = kFT + 2 * PADR;
  float g[R + 1];
#pragma unroll
  for (int j = 0; j <= R; ++j) g[j] = taps[j];
  const int tasks = rows * xblocks;
  for (int t = threadIdx.x; t < tasks; t += kFThreads) {
    const int xb = t / rows, ty = t - xb * rows;
    const float4* src = reinterpret_cast<const float4*>(tile + (ty + 1) * kFTileStride + kFMarg + xb * kFT - PADR);
    float in[WIN];
#pragma unroll
    for (int c = 0; c < WIN / 4; ++c) {
      const float4 v = src[c];
      in[4 * c] = v.x; in[4 * c + 1] = v.y; in[4 * c + 2] = v.z; in[4 * c + 3] = v.w;
    }
    float acc[kFT];
#pragma unroll
    for (int o = 0; o < kFT; ++o) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j <= 2 * R; ++j) a = fmaf(g[j <= R ? j : 2 * R - j], in[PADR - R + o + j], a);
      acc[o] = a;
    }
    *reinterpret_cast<float4*>(tmp + ty * kFTmpStride + xb * kFT) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// column pass: conv(x, y) = sum_j tap[j] * tmp[y + j][x]; up to two tasks (kFT rows each) per thread,
// results stay in registers.
template <int R>
__device__ __forceinline__ void fast_col_pass(const float* __restrict__ tmp, const float* __restrict__ taps, int OW,
                                              int OH, float (&cv)[2][kFT]) {
  float g[R + 1];
#pragma unroll
  for (int j = 0; j <= R; ++j) g[j] = taps[j];
  const int yblocks = (OH + kFT - 1) / kFT;
  const int tasks = OW * yblocks;
#pragma unroll
  for (int slot = 0; slot < 2; ++slot) {
    const int t = threadIdx.x + slot * kFThreads;
#pragma unroll
    for (int o = 0; o < kFT; ++o) cv[slot][o] = -INFINITY;
    if (t < tasks) {
      const int yb = t / OW, x = t - yb * OW;
      float in[kFT + 2 * R];
#pragma unroll
      for (int j = 0; j < kFT + 2 * R; ++j) in[j] = tmp[min(yb * kFT + j, kFTmpRows - 1) * kFTmpStride + x];
#pragma unroll
      for (int o = 0; o < kFT; ++o) {
        float a = 0.0f;
#pragma unroll
        for (int j = 0; j <= 2 * R; ++j) a = fmaf(g[j <= R ? j : 2 * R - j], in[o + j], a);
        if (yb * kFT + o < OH) cv[slot][o] = a;
      }
    }
  }
}

// exact value of one convolved pixel from the gathered tile (reflect halo already materialised):
// double accumulation of the d x d table, float32 result.  Warp-collective.
__device__ __forceinline__ float exact_conv_tile(const float* __restrict__ tile, int ty, int tc, int r,
                                                 const double* __restrict__ w2d) {
  const int d = 2 * r + 1, lane = threadIdx.x & 31;
  double acc = 0.0;
  if (lane < d) {
    const float* p = tile + (ty - r) * kFTileStride + tc - r + lane;
    const double* w = w2d + lane;
    for (int ti = 0; ti < d; ++ti) acc = fma(w[ti * d], static_cast<double>(p[ti * kFTileStride]), acc);
  }
  return static_cast<float>(warp_sum(acc));
}

template <typename T>
__global__ void __launch_bounds__(kFThreads, 6)
decode_expected_fast_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps,
                            float* __restrict__ locs, float* __restrict__ vals, int32_t* __restrict__ argmax,
                            double* __restrict__ keypoints, unsigned plane_bytes, unsigned tile_off) {
  extern __shared__ __align__(128) unsigned char fsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ FastShared sh;

  const T* plane = reinterpret_cast<const T*>(fsm);
  float* tile = reinterpret_cast<float*>(fsm + tile_off);
  float* tmp = tile + kFTileRows * kFTileStride;

  constexpr int V = Elem<T>::kVec;
  const int H = p.H, W = p.W, HW = H * W;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tail = p.apply_tail != 0;
  const float temp = p.temperature;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  int64_t hm = blockIdx.x;
  if (tid == 0 && hm < N) {
    mbar_expect_tx(&bar, plane_bytes);
    tma_load_1d(fsm, heatmaps + hm * HW, plane_bytes, &bar);
  }

  for (int it = 0; hm < N; hm += gridDim.x, ++it) {
    const int k = static_cast<int>(hm % p.K);
    const int r = tab.radius[k];
    const double* w2d = tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;
    if (tid < PP_OKS_TAPS) sh.taps[tid] = tab.taps_f32[k * PP_OKS_TAPS + tid];
    if (tid == 0) {
      sh.bbox[0] = W; sh.bbox[1] = -1; sh.bbox[2] = H; sh.bbox[3] = -1;
      sh.cand_count = 0;
    }
    mbar_wait(&bar, it & 1);

    // ---- A: one scan of the raw plane: max (first index) and min.  The head tail is monotone
    // (temperature > 0 is required for this path), so the raw extrema are the extrema after the tail.
    float xmax = -INFINITY, xmin = INFINITY;
    int imax = 0x7fffffff;
    for (int i = tid; i < HW / V; i += kFThreads) {
      float f[V];
      const uint4 w = *reinterpret_cast<const uint4*>(plane + i * V);
      unpack(w, f, T());
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (f[j] > xmax) { xmax = f[j]; imax = i * V + j; }
        xmin = fminf(xmin, f[j]);
      }
    }
    warp_argmax(xmax, imax);
    xmin = -warp_max(-xmin);
    if (lane == 0) { sh.red_f[0][warp] = xmax; sh.red_f[1][warp] = xmin; sh.red_i[warp] = imax; }
    __syncthreads();
    xmax = sh.red_f[0][0]; xmin = sh.red_f[1][0]; imax = sh.red_i[0];
#pragma unroll
    for (int w = 1; w < kFThreads / 32; ++w) {
      argmax_combine(xmax, imax, sh.red_f[0][w], sh.red_i[w]);
      xmin = fminf(xmin, sh.red_f[1][w]);
    }
    const float vmax = apply_tail<T>(xmax, tail, temp), vmin = apply_tail<T>(xmin, tail, temp);

    bool finished = false;   // uniform across the CTA
    int best = 0;
    float best_val = 0.0f, score = vmax;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    bool interior = false, deferred = false;

    if (vmax == vmin) {
      // constant map (e.g. all zeros after the clamp): all convolved pixels are the same float, the
      // first index wins and (0,0) is a border pixel
      finished = true;
    }

    int ox0 = 0, oy0 = 0, OW = 0, OH = 0;
    if (!finished) {
      // ---- B: L = exact convolved value at the raw maximum p0
      const int py = imax / W, px = imax - py * W;
      if (warp == 0) {
        const int d = 2 * r + 1;
        double acc = 0.0;
        if (lane < d) {
          const int xx = reflect_index(px + lane - r, W);
          for (int ti = 0; ti < d; ++ti) {
            const int yy = reflect_index(py + ti - r, H);
            acc = fma(w2d[ti * d + lane], static_cast<double>(plane_value<T>(plane, yy * W + xx, tail, temp)), acc);
          }
        }
        const float e = static_cast<float>(warp_sum(acc));
        // a few ulp of slack: this evaluation and the tile evaluation sum the same terms in different orders
        if (lane == 0) sh.L = e - fabsf(e) * 2.4e-7f;
      }
      __syncthreads();
      const float L = sh.L;

      // ---- C: bounding box of S = {h >= L}
      int bx0 = W, bx1 = -1, by0 = H, by1 = -1;
      for (int i = tid; i < HW / V; i += kFThreads) {
        float f[V];
        const uint4 w = *reinterpret_cast<const uint4*>(plane + i * V);
        unpack(w, f, T());
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (apply_tail<T>(f[j], tail, temp) >= L) {
            const int idx = i * V + j, y = idx / W, x = idx - y * W;
            bx0 = min(bx0, x); bx1 = max(bx1, x); by0 = min(by0, y); by1 = max(by1, y);
          }
        }
      }
      if (__any_sync(0xffffffffu, bx1 >= 0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
          bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
          by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, o));
          by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, o));
        }
        if (lane == 0) {
          atomicMin(&sh.bbox[0], bx0); atomicMax(&sh.bbox[1], bx1);
          atomicMin(&sh.bbox[2], by0); atomicMax(&sh.bbox[3], by1);
        }
      }
      __syncthreads();
      // p0 itself is in S (h(p0) = max >= R(p0)), so the box is never empty
      ox0 = max(sh.bbox[0] - r, 0);
      oy0 = max(sh.bbox[2] - r, 0);
      OW = min(sh.bbox[1] + r, W - 1) - ox0 + 1;
      OH = min(sh.bbox[3] + r, H - 1) - oy0 + 1;
      if (OW > kFRegion || OH > kFRegion) {
        deferred = true;   // too spread out for the tile: leave it to the full-plane kernel
        finished = true;
      }
    }

    if (!finished) {
      // ---- D: gather the region (+ radius + 1 halo, reflect-extended, tail applied) into the tile
      const int rows = OH + 2 * (r + 1);
      const int c_lo = kFMarg - (r + 1), c_hi = kFMarg + OW + r;   // inclusive
      for (int ty = warp; ty < rows; ty += kFThreads / 32) {
        const int yy = reflect_index(oy0 - (r + 1) + ty, H);
        for (int c = c_lo + lane; c <= c_hi; c += 32) {
          const int xx = reflect_index(ox0 - kFMarg + c, W);
          tile[ty * kFTileStride + c] = plane_value<T>(plane, yy * W + xx, tail, temp);
        }
      }
    }
    __syncthreads();   // the plane is not read after this point
    {
      const int64_t nxt = hm + gridDim.x;
      if (tid == 0 && nxt < N) {   // the next heatmap streams in while this one is finished from the tile
        mbar_expect_tx(&bar, plane_bytes);
        tma_load_1d(fsm, heatmaps + nxt * HW, plane_bytes, &bar);
      }
    }

    if (!finished) {
      // ---- E/F: separable float32 prefilter on the tile
      float cv[2][kFT];
      const int trows = OH + 2 * r, xblocks = (OW + kFT - 1) / kFT;
      switch (r) {
#define PP_FCASE(R)                                         \
  case R:                                                   \
    fast_row_pass<R>(tile, tmp, sh.taps, trows, xblocks);   \
    __syncthreads();                                        \
    fast_col_pass<R>(tmp, sh.taps, OW, OH, cv);             \
    break;
        PP_FCASE(1) PP_FCASE(2) PP_FCASE(3) PP_FCASE(4) PP_FCASE(5) PP_FCASE(6) PP_FCASE(7) PP_FCASE(8) PP_FCASE(9)
#undef PP_FCASE
        default: break;
      }
      float pmax = -INFINITY;
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int o = 0; o < kFT; ++o) pmax = fmaxf(pmax, cv[s][o]);
      pmax = warp_max(pmax);
      if (lane == 0) sh.red_f[0][warp] = pmax;
      __syncthreads();
      pmax = fmaxf(fmaxf(sh.red_f[0][0], sh.red_f[0][1]), fmaxf(sh.red_f[0][2], sh.red_f[0][3]));
      const float amax = fmaxf(fabsf(vmax), fabsf(vmin));
      const float gamma = static_cast<float>(2 * (2 * r + 1) + 8) * 1.1920929e-7f;
      const float thr = pmax - (2.0f * gamma + 4.0f * 5.9604645e-8f) * amax;
      const int yblocks = (OH + kFT - 1) / kFT;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int t = tid + s * kFThreads;
        if (t < OW * yblocks) {
          const int yb = t / OW, x = t - yb * OW;
#pragma unroll
          for (int o = 0; o < kFT; ++o) {
            if (cv[s][o] >= thr) {   // cv is -inf for rows beyond the region
              const int slot = atomicAdd(&sh.cand_count, 1);
              if (slot < kMaxCand) sh.cand[slot] = (oy0 + yb * kFT + o) * W + ox0 + x;
            }
          }
        }
      }
      __syncthreads();
      const int count = sh.cand_count;
      if (count > kMaxCand) {
        deferred = true;   // plateau wider than the list: full-plane kernel
      } else {
        // ---- G: exact values of the near-maximal pixels
        float wv = -INFINITY;
        int wi = 0x7fffffff;
        for (int c = warp; c < count; c += kFThreads / 32) {
          const int idx = sh.cand[c], y = idx / W, x = idx - y * W;
          const float e = exact_conv_tile(tile, y - oy0 + r + 1, x - ox0 + kFMarg, r, w2d);
          argmax_combine(wv, wi, e, idx);
        }
        if (lane == 0) { sh.red_f[0][warp] = wv; sh.red_i[warp] = wi; }
        __syncthreads();
        best_val = sh.red_f[0][0]; best = sh.red_i[0];
#pragma unroll
        for (int w = 1; w < kFThreads / 32; ++w) argmax_combine(best_val, best, sh.red_f[0][w], sh.red_i[w]);
        const int by = best / W, bx = best - by * W;
        interior = bx > 0 && bx < W - 1 && by > 0 && by < H - 1;
        if (interior) {
          const int dx = (warp == 0) ? -1 : (warp == 1) ? 1 : 0;
          const int dy = (warp == 2) ? -1 : (warp == 3) ? 1 : 0;
          const float e = exact_conv_tile(tile, by + dy - oy0 + r + 1, bx + dx - ox0 + kFMarg, r, w2d);
          if (lane == 0) sh.nb[warp] = e;
          __syncthreads();
#pragma unroll
          for (int q = 0; q < 4; ++q) nb[q] = sh.nb[q];
        }
        score = tile[(by - oy0 + r + 1) * kFTileStride + bx - ox0 + kFMarg];
      }
    }

    // ---- H: outputs
    if (tid == 0) {
      if (deferred) {
        locs[hm * 2] = __int_as_float(0x7fc00000);   // NaN marks "finish me" for the full-plane kernel
      } else {
        const int by = best / W, bx = best - by * W;
        float fx = static_cast<float>(bx), fy = static_cast<float>(by);
        if (interior) {   // _get_subpixel_maximums, float32, op order of heatmap.py:136-165
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
        vals[hm] = score;
        if (argmax) argmax[hm] = best;
        if (keypoints) {
          keypoints[hm * 2] = static_cast<double>(fx) / static_cast<double>(W - 1) * p.input_w;
          keypoints[hm * 2 + 1] = static_cast<double>(fy) / static_cast<double>(H - 1) * p.input_h;
        }
      }
    }
    __syncthreads();   // tile / scratch are reused by the next heatmap
  }
}

// Main kernel of the argmax + DARK-UDP decoder (included into pp_decode.cu after pp_decode_fast.cuh, whose
// building blocks it shares).  Reference: ArgMaxProbMap.decode (codec.py:515-543), gaussian_blur
// (codec.py:284-313), refine_keypoints_dark_udp (codec.py:315-375).
//
// The refinement needs the blurred map at the 3 x 3 neighbourhood of the raw peak p0 and the GLOBAL maximum
// of the blurred map (the blur is rescaled so that the maximum is preserved, codec.py:312, and the rescale
// decides which pixels the clip to [1e-3, 50] touches).  The same pruning as in the expected-OKS decoder
// applies: the blur taps are non-negative and sum to 1 (zero padding only removes mass), so
//     blurred(p) <= max over the 11 x 11 window of h      (for positive maxima),
// and blurred(p0) >= L for a cheap lower bound L.  The blurred maximum therefore lies within the blur radius
// of S = {h >= L}: only the bounding box of S, dilated by the radius, is blurred (row pass, then column
// pass -- the order of cv2's separable filter and of the full-plane kernel), from a small zero-padded tile.
// Flat / noisy maps whose box does not fit the tile are blurred whole (column pass from the plane, row pass
// on a zero-padded float32 plane).
#pragma once

struct DarkShared {
  float red_f[2][kFWarps];
  __align__(16) float grow[kFTaps];
  __align__(16) float gcol[kFTaps];
  float L;
  int p0;
  int bbox[4];
  float stencil[8];   // blurred values at p0 and its six DARK neighbours
};

// float32 derivatives in the reference's operation order (codec.py:361-368), float64 2 x 2 pseudo-inverse
// (codec.py:371), float32 keypoint minus float64 shift stored as float32 (codec.py:372-373).
// b[] = rescaled, clipped, log-ed blur at: centre, x+1, x-1, y+1, y-1, (x+1,y+1), (x-1,y-1).
__device__ __forceinline__ void dark_shift(const float (&b)[7], int px, int py, float& fx, float& fy) {
  const float c = b[0], xp = b[1], xm = b[2], yp = b[3], ym = b[4], pp_ = b[5], mm = b[6];
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
  fx = static_cast<float>(static_cast<double>(static_cast<float>(px)) - sx);
  fy = static_cast<float>(static_cast<double>(static_cast<float>(py)) - sy);
}

// heatmaps[k] *= origin_max / (max(blurred) + 1e-12) (float32, codec.py:312); clip (codec.py:343); log (:344)
__device__ __forceinline__ float dark_log(float blurred, float ratio) {
  float v = __fmul_rn(blurred, ratio);
  v = fminf(fmaxf(v, 1e-3f), 50.0f);
  return static_cast<float>(log(static_cast<double>(v)));
}

template <typename T>
__global__ void __launch_bounds__(kFThreads, 6)
decode_dark_fast_kernel(pp_decode_params p, const float* __restrict__ blur_taps, int ksize,
                        const T* __restrict__ heatmaps, float* __restrict__ peaks, float* __restrict__ scores,
                        float* __restrict__ refined, double* __restrict__ keypoints, FastGeom geo,
                        unsigned* __restrict__ work_counter, const int* __restrict__ list,
                        const unsigned* __restrict__ list_count) {
  extern __shared__ __align__(128) unsigned char fsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ DarkShared sh;
  __shared__ long long next_item;

  const T* plane = reinterpret_cast<const T*>(fsm);
  T* plane_rw = reinterpret_cast<T*>(fsm);
  float* work = reinterpret_cast<float*>(fsm + geo.work_off);
  float* tile = work;
  float* tmp = work + kFTileRows * kFTileStride;

  constexpr int V = Elem<T>::kVec;
  const int H = p.H, W = p.W, HW = H * W, WV = W / V, FS = geo.full_stride;
  // `list` (with its device-side length): decode only the listed heatmaps -- the ones the tensor-core kernel
  // (pp_decode_mma.cuh, kDark) handed on
  const int64_t N = list ? static_cast<int64_t>(*list_count) : static_cast<int64_t>(p.B) * p.K;
  auto item_to_hm = [&](long long j) -> long long { return list ? static_cast<long long>(list[j]) : j; };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tail = p.apply_tail != 0;
  const float temp = p.temperature;
  const int step_y = fast_div(kFThreads, geo.div_WV), step_x = kFThreads - step_y * WV;
  const int first_y = fast_div(tid, geo.div_WV), first_x = tid - first_y * WV;
  const int r = ksize >> 1, d = ksize;
  const int padr = (r + 3) & ~3, shift = padr - r;
  const int nch_row = (shift + d + 7) >> 3, nch_col = (d + 7) >> 3;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (tid < kFTaps) {
    sh.gcol[tid] = tid < d ? blur_taps[tid] : 0.0f;
    sh.grow[tid] = (tid >= shift && tid < shift + d) ? blur_taps[tid - shift] : 0.0f;
  }
  for (int i = tid; i < static_cast<int>(geo.work_floats); i += kFThreads) work[i] = 0.0f;
  const bool dynamic = work_counter != nullptr;
  if (tid == 0) {
    next_item = dynamic ? static_cast<long long>(atomicAdd(work_counter, 1u)) : static_cast<long long>(blockIdx.x);
    if (next_item < N) {
      mbar_expect_tx(&bar, geo.plane_bytes);
      tma_load_1d(fsm, heatmaps + item_to_hm(next_item) * HW, geo.plane_bytes, &bar);
    }
  }
  __syncthreads();
  long long item = next_item;
  bool pads_dirty = false;   // the tile path scribbles over what the full path uses as zero pad columns

  for (int it = 0; item < N; ++it) {
    const long long hm = item_to_hm(item);
    if (tid == 0) { sh.bbox[0] = W; sh.bbox[1] = -1; sh.bbox[2] = H; sh.bbox[3] = -1; sh.p0 = 0x7fffffff; }
    mbar_wait(&bar, it & 1);

    // ---- A: optional head tail in place, then max / min and the first index of the maximum
    if (tail) {
      for (int i = tid; i < HW / V; i += kFThreads) {
        float f[V];
        uint4* vec = reinterpret_cast<uint4*>(plane_rw + i * V);
        unpack(*vec, f, T());
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = tail_value<T>(f[j], temp);
        *vec = pack(f, T());
      }
    }
    float xmax = -INFINITY, xmin = INFINITY;
#pragma unroll 2
    for (int i = tid; i < HW / V; i += kFThreads) {
      float f[V];
      unpack(*reinterpret_cast<const uint4*>(plane + i * V), f, T());
#pragma unroll
      for (int j = 0; j < V; ++j) { xmax = fmaxf(xmax, f[j]); xmin = fminf(xmin, f[j]); }
    }
    const float tmax = xmax;
    xmax = warp_max(xmax);
    xmin = -warp_max(-xmin);
    if (lane == 0) { sh.red_f[0][warp] = xmax; sh.red_f[1][warp] = xmin; }
    __syncthreads();
    xmax = fmaxf(fmaxf(sh.red_f[0][0], sh.red_f[0][1]), fmaxf(sh.red_f[0][2], sh.red_f[0][3]));
    xmin = fminf(fminf(sh.red_f[1][0], sh.red_f[1][1]), fminf(sh.red_f[1][2], sh.red_f[1][3]));
    if (tmax == xmax) {   // NumPy argmax: lowest flat index among the maxima
      int idx = 0x7fffffff;
      for (int i = tid; i < HW / V && idx == 0x7fffffff; i += kFThreads) {
        float f[V];
        unpack(*reinterpret_cast<const uint4*>(plane + i * V), f, T());
#pragma unroll
        for (int j = V - 1; j >= 0; --j) idx = (f[j] == xmax) ? i * V + j : idx;
      }
      atomicMin(&sh.p0, idx);
    }
    __syncthreads();
    const int p0 = sh.p0;
    const int py = fast_div(p0, geo.div_W), px = p0 - py * W;
    const float top = xmax;
    const bool empty = !(top > 0.0f);   // locs[vals <= 0] = -1 (heatmap.py:46): the sentinel is kept
    bool tile_path = false;
    int ox0 = 0, oy0 = 0, OW = 0, OH = 0;

    if (!empty) {
      // ---- B: lower bound L of the blurred maximum from the central 5 x 5 taps at p0
      if (warp == 0) {
        const int half = min(2, r), side = 2 * half + 1;
        double sw = 0.0, swh = 0.0;
        if (lane < side * side) {
          const int ti = lane / side, tj = lane - ti * side;
          const int yy = py - half + ti, xx = px - half + tj;
          const double w = static_cast<double>(sh.gcol[r - half + ti]) * static_cast<double>(sh.gcol[r - half + tj]);
          sw = w;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) swh = w * static_cast<double>(plane_value<T>(plane, yy * W + xx));
        }
        sw = warp_sum(sw); swh = warp_sum(swh);
        if (lane == 0) {
          const float e = static_cast<float>(swh + fmax(1.0 - sw, 0.0) * static_cast<double>(fminf(xmin, 0.0f)));
          sh.L = e - fabsf(e) * 1e-5f - 1e-37f;   // slack covers the float32 rounding of the blur and sum(taps) != 1
        }
      }
      __syncthreads();
      const float L = sh.L;

      // ---- C: bounding box of S = {h >= L}
      int bx0 = W, bx1 = -1, by0 = H, by1 = -1;
      {
        int y = first_y, xv = first_x;
#pragma unroll 2
        for (int i = tid; i < HW / V; i += kFThreads) {
          float f[V];
          unpack(*reinterpret_cast<const uint4*>(plane + i * V), f, T());
          float m = f[0];
#pragma unroll
          for (int j = 1; j < V; ++j) m = fmaxf(m, f[j]);
          if (m >= L) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              if (f[j] >= L) { bx0 = min(bx0, xv * V + j); bx1 = max(bx1, xv * V + j); }
            }
            by0 = min(by0, y); by1 = max(by1, y);
          }
          xv += step_x; y += step_y;
          if (xv >= WV) { xv -= WV; ++y; }
        }
      }
      {
        const bool mine = bx1 >= 0;
        const unsigned holders = __ballot_sync(0xffffffffu, mine);
        if (holders != 0u && __popc(holders) <= 6) {
          if (mine) {
            atomicMin(&sh.bbox[0], bx0); atomicMax(&sh.bbox[1], bx1);
            atomicMin(&sh.bbox[2], by0); atomicMax(&sh.bbox[3], by1);
          }
        } else if (holders != 0u) {
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
      }
      __syncthreads();
      // p0 is in S (h(p0) = max >= blurred(p0) >= L for a positive maximum); make sure of it anyway
      ox0 = max(min(sh.bbox[0], px) - r, 0);
      oy0 = max(min(sh.bbox[2], py) - r, 0);
      OW = min(max(sh.bbox[1], px) + r, W - 1) - ox0 + 1;
      OH = min(max(sh.bbox[3], py) + r, H - 1) - oy0 + 1;
      tile_path = OW <= kFRegion && OH <= kFRegion;
    }

    if (!empty && tile_path) {
      // ---- D: gather the region (+ radius halo) into the tile; outside the map the blur sees zeros
      const int rows = OH + 2 * (r + 1);
      const int c_lo = kFMarg - (r + 1), c_hi = kFMarg + OW + r;
      const int ncols = c_hi - c_lo + 1, total = rows * ncols;
      const unsigned mcols = div_magic(ncols);
#pragma unroll 4
      for (int e = tid; e < total; e += kFThreads) {
        const int ty = fast_div(e, mcols), c = c_lo + (e - ty * ncols);
        const int yy = oy0 - (r + 1) + ty, xx = ox0 - kFMarg + c;
        tile[ty * kFTileStride + c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? plane_value<T>(plane, yy * W + xx) : 0.0f;
      }
      pads_dirty = true;
    } else if (!empty) {
      // ---- D': full path.  The pad columns of the float32 plane must be zero (zero padding in x).
      if (pads_dirty) {
        const int right = FS - kFMarg - W;
        for (int i = tid; i < H * (kFMarg + right); i += kFThreads) {
          const int y = i / (kFMarg + right), j = i - y * (kFMarg + right);
          work[y * FS + (j < kFMarg ? j : W + j)] = 0.0f;
        }
        for (int i = tid + H * FS; i < static_cast<int>(geo.work_floats); i += kFThreads) work[i] = 0.0f;
        __syncthreads();
        pads_dirty = false;
      }
      const int yblocks = (H + 3) >> 2, W2 = W >> 1, nch4 = (d + 3) >> 2;
      const unsigned mW2 = div_magic(W2);
      for (int t = tid; t < W2 * yblocks; t += kFThreads) {
        const int yb = fast_div(t, mW2), x = (t - yb * W2) * 2;
        const int y0 = yb * 4;
        if (y0 - r >= 0 && y0 + 3 + r < H)
          full_col_task<T, true, true>(plane, work, sh.gcol, nch4, x, y0, r, H, W, FS);
        else
          full_col_task<T, false, true>(plane, work, sh.gcol, nch4, x, y0, r, H, W, FS);
      }
    }
    __syncthreads();
    const bool plane_free = empty || tile_path;
    auto fetch_next = [&]() {   // thread 0
      const long long j = dynamic ? static_cast<long long>(atomicAdd(work_counter, 1u)) : item + static_cast<long long>(gridDim.x);
      next_item = j;
      if (j < N) {
        fence_proxy_async();
        mbar_expect_tx(&bar, geo.plane_bytes);
        tma_load_1d(fsm, heatmaps + item_to_hm(j) * HW, geo.plane_bytes, &bar);
      }
    };
    if (plane_free && tid == 0) fetch_next();

    float fx = -1.0f, fy = -1.0f;
    if (!empty) {
      // ---- E/F: the blur (row pass, then column pass on the tile; column then row pass on the full plane)
      float bmax = -INFINITY;
      if (tile_path) {
        const int trows = OH + 2 * r, xblocks = (OW + 7) >> 3;
        const unsigned mrows = div_magic(trows);
        for (int t = tid; t < trows * xblocks; t += kFThreads) {
          const int xb = fast_div(t, mrows), ty = t - xb * trows;
          float acc[8];
          conv8_contiguous(tile + (ty + 1) * kFTileStride + kFMarg + xb * 8 - padr, sh.grow, nch_row, acc);
          float4* dst = reinterpret_cast<float4*>(tmp + ty * kFTmpStride + xb * 8);
          dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        __syncthreads();
        const int yblocks = (OH + 7) >> 3;
        if (tid < OW * yblocks) {
          const int cyb = fast_div(tid, div_magic(OW)), cx = tid - cyb * OW;
          const float* colp = tmp + cx;
          const int y0 = cyb * 8;
          float acc[8];
          conv8_gather([&](int j) -> float { return colp[min(y0 + j, kFTmpRows + 7) * kFTmpStride]; }, sh.gcol, nch_col,
                       acc);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            if (y0 + o < OH) {
              bmax = fmaxf(bmax, acc[o]);
              tile[(y0 + o) * kFRegion + cx] = acc[o];   // the tile is free now: keep the blurred region there
            }
          }
        }
      } else {
        const int tasks = (geo.W8 >> 3) * H;
        for (int t = tid; t < tasks; t += kFThreads) {
          const int xb = fast_div(t, geo.div_H), y = t - xb * H;
          float acc[8];
          conv8_contiguous(work + y * FS + kFMarg + xb * 8 - padr, sh.grow, nch_row, acc);
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (xb * 8 + o < W) bmax = fmaxf(bmax, acc[o]);
        }
      }
      bmax = warp_max(bmax);
      if (lane == 0) sh.red_f[0][warp] = bmax;
      __syncthreads();
      bmax = fmaxf(fmaxf(sh.red_f[0][0], sh.red_f[0][1]), fmaxf(sh.red_f[0][2], sh.red_f[0][3]));

      // ---- the seven stencil points around p0, edge-clamped (codec.py:346-359)
      if (tid < 7) {
        const int dxs[7] = {0, 1, -1, 0, 0, 1, -1}, dys[7] = {0, 0, 0, 1, -1, 1, -1};
        const int yy = min(max(py + dys[tid], 0), H - 1), xx = min(max(px + dxs[tid], 0), W - 1);
        float v;
        if (tile_path) {
          v = tile[(yy - oy0) * kFRegion + (xx - ox0)];
        } else {   // recompute from the column-passed plane with the row pass's operation order
          const float* c = work + yy * FS + kFMarg + xx - r;
          v = 0.0f;
          for (int j = 0; j < d; ++j) v = fmaf(sh.gcol[j], c[j], v);
        }
        sh.stencil[tid] = v;
      }
      __syncthreads();
      if (tid == 0) {
        const float ratio = __fdiv_rn(top, __fadd_rn(bmax, 1e-12f));
        float b[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) b[q] = dark_log(sh.stencil[q], ratio);
        dark_shift(b, px, py, fx, fy);
      }
    }

    if (tid == 0) {
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
    if (!plane_free) {
      if (tid == 0) fetch_next();
      __syncthreads();
    }
    item = next_item;
  }
}

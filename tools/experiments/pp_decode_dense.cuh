// Expected-OKS decoder, dense variant (included into pp_decode.cu after pp_decode_fast.cuh, same namespace).
//
// pp_decode_fast.cuh prunes the convolution to the neighbourhood of {h >= L}; that wins on clean blob-shaped
// maps but its control flow (two scans, a bound, a bounding box, two code paths, radius-generic filter loops)
// costs ~8.5 k warp instructions per 64x48 heatmap on a realistic mix, 85 % of them not multiply-adds, and a
// third of the predictions of a training batch (unlabelled / out-of-image keypoints: noise-only maps) take the
// unpruned path anyway.  This kernel does the opposite: no pruning, no data-dependent path, the float32
// separable prefilter over the whole map written as two perfectly regular, fully unrolled passes per radius
// (template parameter), so that nearly every instruction is an FFMA or a 128-bit shared-memory access:
//
//   (A) TMA bulk copy of the plane (double buffered: the next heatmap lands during this one);
//   (B) one sweep: head tail (optional) + conversion to float32 + copy into a padded plane P whose left / right
//       margins hold the reflected columns; min / max on the way (constant maps finish here);
//   (C) row pass  P -> Q: a task = 8 consecutive outputs of one row from an aligned 128-bit window; rows within
//       `radius` of the top / bottom edge are also stored to their mirror rows, so Q is reflect-padded in y;
//   (D) column pass Q -> R: a task = 8 consecutive rows of one column (lanes = consecutive columns);
//   (E) candidates = pixels whose prefilter value lies within the rigorous error band of its maximum;
//   (F) exact double-precision re-evaluation of the candidates and of the winner's neighbours with the
//       reference's d x d table (scipy's arithmetic), NumPy tie-break, float32 sub-pixel fit -- as in the pruned
//       kernel, whose helpers are reused.
#pragma once

constexpr int kDThreads = 192;
constexpr int kDWarps = kDThreads / 32;
constexpr int kDOrderMax = 512;   // channel order staged in shared memory up to this many channels
constexpr int kDMargin = 12;   // reflect margin of P (columns) and Q (rows): >= radius, multiple of 4

struct DenseGeom {
  unsigned plane_bytes;   // H * W * sizeof(T)
  unsigned raw_stride;    // bytes between the two raw stages
  unsigned p_off, q_off, w2d_off, part_off;   // w2d: two tables of PP_OKS_TAPS^2 doubles
  int PS, QS;             // row pitch of P / Q in floats (multiple of 4, odd number of 16-byte groups)
  int Wp, Hp;             // W, H rounded up to 8
  unsigned q_floats;
  unsigned div_WV, div_W, div_H, div_Wp, div_Wp4;
};

struct DenseShared {
  float red_max[kDWarps], red_min[kDWarps];
  int red_i[kDWarps];
  __align__(16) float taps[2][PP_OKS_TAPS + 5];   // per-channel tables are double buffered (prefetched)
  int ev_idx[5];
  float ev_val[5];
  int cand[kMaxCand];
  int cand_count;
};

// 8 outputs of a (2R+1)-tap filter from a register window; win[SH + o + j] is the j-th input of output o
template <int R, int NW, int SH>
__device__ __forceinline__ void fir8(const float (&g)[2 * R + 1], const float (&win)[NW], float (&acc)[8]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.0f;
#pragma unroll
  for (int j = 0; j <= 2 * R; ++j)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = fmaf(g[j], win[SH + o + j], acc[o]);
}

// passes (C) and (D) for one radius; returns this thread's maximum over the valid outputs
template <int R>
__device__ __forceinline__ float dense_passes(const DenseGeom& geo, const float* __restrict__ taps, float* __restrict__ P,
                                              float* __restrict__ Q, int H, int W) {
  constexpr int R4 = (R + 3) & ~3, NW = 8 + 2 * R4;
  const int tid = threadIdx.x, PS = geo.PS, QS = geo.QS;
  float g[2 * R + 1];
#pragma unroll
  for (int j = 0; j <= 2 * R; ++j) g[j] = taps[j];

  // (C) rows: task t -> (x block, row), consecutive lanes on consecutive rows
  const int row_tasks = (geo.Wp >> 3) * H;
  for (int t = tid; t < row_tasks; t += kDThreads) {
    const int xb = fast_div(t, geo.div_H), y = t - xb * H;
    const float4* src = reinterpret_cast<const float4*>(P + y * PS + kDMargin + xb * 8 - R4);
    float win[NW];
#pragma unroll
    for (int i = 0; i < NW / 4; ++i) {
      const float4 v = src[i];
      win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
    }
    float acc[8];
    fir8<R, NW, R4 - R>(g, win, acc);
    const float4 lo = make_float4(acc[0], acc[1], acc[2], acc[3]), hi = make_float4(acc[4], acc[5], acc[6], acc[7]);
    float4* dst = reinterpret_cast<float4*>(Q + (kDMargin + y) * QS + xb * 8);
    dst[0] = lo; dst[1] = hi;
    if (y < R) {                       // mirror row -1 - y
      float4* m = reinterpret_cast<float4*>(Q + (kDMargin - 1 - y) * QS + xb * 8);
      m[0] = lo; m[1] = hi;
    }
    if (y >= H - R) {                  // mirror row 2H - 1 - y
      float4* m = reinterpret_cast<float4*>(Q + (kDMargin + 2 * H - 1 - y) * QS + xb * 8);
      m[0] = lo; m[1] = hi;
    }
  }
  __syncthreads();

  // (D) columns: task t -> (row block, column), consecutive lanes on consecutive columns; R overwrites P
  float pmax = -INFINITY;
  const int col_tasks = geo.Wp * (geo.Hp >> 3);
  for (int t = tid; t < col_tasks; t += kDThreads) {
    const int yb = fast_div(t, geo.div_Wp), x = t - yb * geo.Wp;
    const int y0 = yb * 8;
    const float* col = Q + (kDMargin + y0 - R) * QS + x;
    float win[8 + 2 * R];
#pragma unroll
    for (int j = 0; j < 8 + 2 * R; ++j) win[j] = col[j * QS];
    float acc[8];
    fir8<R, 8 + 2 * R, 0>(g, win, acc);
    float* out = P + y0 * geo.Wp + x;
#pragma unroll
    for (int o = 0; o < 8; ++o) out[o * geo.Wp] = acc[o];
    if (x < W) {
      if (y0 + 8 <= H) {
        pmax = fmaxf(pmax, fmaxf(fmaxf(fmaxf(acc[0], acc[1]), fmaxf(acc[2], acc[3])), fmaxf(fmaxf(acc[4], acc[5]), fmaxf(acc[6], acc[7]))));
      } else {
#pragma unroll
        for (int o = 0; o < 8; ++o)
          if (y0 + o < H) pmax = fmaxf(pmax, acc[o]);
      }
    }
  }
  return pmax;
}

template <typename T>
__global__ void __launch_bounds__(kDThreads, 3)
decode_expected_dense_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps,
                             float* __restrict__ locs, float* __restrict__ vals, int32_t* __restrict__ argmax,
                             double* __restrict__ keypoints, DenseGeom geo, unsigned* __restrict__ work_counter) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ DenseShared sh;
  __shared__ int q_hm[2], q_k[2];
  __shared__ int s_order[kDOrderMax];

  float* P = reinterpret_cast<float*>(dsm + geo.p_off);
  float* Q = reinterpret_cast<float*>(dsm + geo.q_off);
  double* w2d = reinterpret_cast<double*>(dsm + geo.w2d_off);
  double* ev_part = reinterpret_cast<double*>(dsm + geo.part_off);

  constexpr int V = Elem<T>::kVec;
  const int H = p.H, W = p.W, HW = H * W, WV = W / V;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tail = p.apply_tail != 0;
  const float temp = p.temperature;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  // Q's rows beyond the mirrored band and P's slack are read by outputs that are discarded; they must be finite
  for (int i = tid; i < static_cast<int>(geo.q_floats); i += kDThreads) Q[i] = 0.0f;
  for (int i = tid; i < geo.PS * geo.Hp; i += kDThreads) P[i] = 0.0f;
  __syncthreads();

  // ---- work queue.  Items come from a global counter, channel-major and widest kernels first (see
  // pp_decode_fast.cuh).  Drawing a ticket is a global atomic (~1 us round trip) and turning it into a heatmap
  // takes two integer divisions, so the queue runs two items ahead: the ticket of item i+2 is drawn at the top of
  // iteration i, decoded by thread 0 at its end, and its plane is copied into the stage item i used as soon as
  // iteration i has ended.  Slots and stages are indexed by item parity.
  const bool dynamic = work_counter != nullptr;
  const bool order_smem = dynamic && tab.order != nullptr && p.K <= kDOrderMax;
  if (order_smem)
    for (int i = tid; i < p.K; i += kDThreads) s_order[i] = tab.order[i];
  unsigned static_next = blockIdx.x;   // thread 0: without a counter the items are strided statically
  auto draw = [&]() -> unsigned {      // thread 0
    if (dynamic) return atomicAdd(work_counter, 1u);
    const unsigned j = static_next;
    static_next += gridDim.x;
    return j;
  };
  auto publish = [&](int slot, unsigned j) {   // thread 0: describe item j (heatmap, channel) in a queue slot
    int hm_ = -1, k_ = -1;
    if (static_cast<int64_t>(j) < N) {
      if (dynamic) {
        const unsigned sl = j / static_cast<unsigned>(p.B), bi = j - sl * static_cast<unsigned>(p.B);
        k_ = tab.order ? (order_smem ? s_order[sl] : tab.order[sl]) : static_cast<int>(sl);
        hm_ = static_cast<int>(bi) * p.K + k_;
      } else {
        hm_ = static_cast<int>(j);
        k_ = static_cast<int>(j % static_cast<unsigned>(p.K));
      }
    }
    q_hm[slot] = hm_;
    q_k[slot] = k_;
  };
  auto start_copy = [&](int slot) {            // thread 0: plane of the item in `slot` -> raw stage `slot`
    if (q_hm[slot] >= 0) {
      fence_proxy_async();
      mbar_expect_tx(&bars[slot], geo.plane_bytes);
      tma_load_1d(dsm + static_cast<size_t>(slot) * geo.raw_stride, heatmaps + static_cast<int64_t>(q_hm[slot]) * HW,
                  geo.plane_bytes, &bars[slot]);
    }
  };
  __syncthreads();   // s_order
  if (tid == 0) {
    publish(0, draw());
    start_copy(0);
    publish(1, draw());
    start_copy(1);
  }
  __syncthreads();
  int64_t hm = q_hm[0];
  // per-channel tables (float32 taps, the reference's d x d doubles): staged for the first item here, afterwards
  // loaded into registers one heatmap ahead and parked in the other buffer, off the critical path
  int k = q_k[0], tb = 0, r = 1, d = 3;
  if (hm >= 0) {
    r = tab.radius[k];
    d = 2 * r + 1;
    const double* src = tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;
    for (int i = tid; i < PP_OKS_TAPS * PP_OKS_TAPS; i += kDThreads) w2d[i] = src[i];
    if (tid < PP_OKS_TAPS) sh.taps[0][tid] = tid < d ? tab.taps_f32[k * PP_OKS_TAPS + tid] : 0.0f;
  }
  __syncthreads();

  for (int it = 0; hm >= 0; ++it) {
    const int s = it & 1;
    const T* plane = reinterpret_cast<const T*>(dsm + static_cast<size_t>(s) * geo.raw_stride);
    T* plane_rw = reinterpret_cast<T*>(dsm + static_cast<size_t>(s) * geo.raw_stride);
    const double* w2d_cur = w2d + tb * (PP_OKS_TAPS * PP_OKS_TAPS);
    const float* taps_cur = sh.taps[tb];
    unsigned ticket = 0;
    if (tid == 0) {
      if (it > 0) start_copy(s ^ 1);   // item it+1 into the stage that item it-1 has left (described at the end of it-1)
      ticket = draw();                 // item it+2; decoded at the end of this iteration
      sh.cand_count = 0;
    }
    mbar_wait(&bars[s], (it >> 1) & 1);

    // ---- B: tail + float32 + reflect-padded copy, min / max
    float xmax = -INFINITY, xmin = INFINITY;
    {
      int y = fast_div(tid, geo.div_WV), xv = tid - y * WV;
      const int step_y = fast_div(kDThreads, geo.div_WV), step_x = kDThreads - step_y * WV;
      for (int i = tid; i < HW / V; i += kDThreads) {
        float f[V];
        uint4* vec = reinterpret_cast<uint4*>(plane_rw + i * V);
        unpack(*vec, f, T());
        if (tail) {
#pragma unroll
          for (int j = 0; j < V; ++j) f[j] = tail_value<T>(f[j], temp);
          *vec = pack(f, T());           // the exact evaluation reads the raw stage
        }
#pragma unroll
        for (int j = 0; j < V; ++j) { xmax = fmaxf(xmax, f[j]); xmin = fminf(xmin, f[j]); }
        float* row = P + y * geo.PS + kDMargin;
        const int x0 = xv * V;
#pragma unroll
        for (int j = 0; j < V; j += 4) *reinterpret_cast<float4*>(row + x0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        xv += step_x; y += step_y;
        if (xv >= WV) { xv -= WV; ++y; }
      }
    }
    // reflected margins, r columns on each side, straight from the raw stage (every lane busy; the in-place
    // tail must have reached the whole stage first)
    if (tail) __syncthreads();
    {
      const int sh_slots = (2 * r <= 16) ? 4 : 5, slots = 1 << sh_slots;
      for (int e = tid; e < (H << sh_slots); e += kDThreads) {
        const int y = e >> sh_slots, j = e & (slots - 1);
        if (j < 2 * r) {
          const bool right = j >= r;
          const int jj = right ? j - r : j;
          P[y * geo.PS + kDMargin + (right ? W + jj : -1 - jj)] = plane_value<T>(plane, y * W + (right ? W - 1 - jj : jj));
        }
      }
    }
    xmax = warp_max(xmax);
    xmin = -warp_max(-xmin);
    if (lane == 0) { sh.red_max[warp] = xmax; sh.red_min[warp] = xmin; }
    __syncthreads();
    float vmax = sh.red_max[0], vmin = sh.red_min[0];
#pragma unroll
    for (int w = 1; w < kDWarps; ++w) { vmax = fmaxf(vmax, sh.red_max[w]); vmin = fminf(vmin, sh.red_min[w]); }

    // the tables of the next heatmap, if its channel differs, travel into registers during this heatmap
    const int k_next = q_k[s ^ 1];
    const bool new_tables = k_next >= 0 && k_next != k;
    double w_next[2] = {0.0, 0.0};
    float tap_next = 0.0f;
    int r_next = r;
    if (new_tables) {
      const double* src = tab.kernel2d + static_cast<size_t>(k_next) * PP_OKS_TAPS * PP_OKS_TAPS;
      w_next[0] = src[tid];
      if (tid + kDThreads < PP_OKS_TAPS * PP_OKS_TAPS) w_next[1] = src[tid + kDThreads];
      r_next = tab.radius[k_next];
      if (tid < PP_OKS_TAPS) tap_next = tab.taps_f32[k_next * PP_OKS_TAPS + tid];
    }

    const bool constant = vmax == vmin;   // first index wins, border pixel: no refinement
    int best = 0;
    float best_val = 0.0f, score = vmax;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    bool interior = false;

    if (!constant) {
      // ---- C / D: the two filter passes, specialised per radius (uniform across the CTA)
      float pmax;
      switch (r) {
        case 1: pmax = dense_passes<1>(geo, taps_cur, P, Q, H, W); break;
        case 2: pmax = dense_passes<2>(geo, taps_cur, P, Q, H, W); break;
        case 3: pmax = dense_passes<3>(geo, taps_cur, P, Q, H, W); break;
        case 4: pmax = dense_passes<4>(geo, taps_cur, P, Q, H, W); break;
        case 5: pmax = dense_passes<5>(geo, taps_cur, P, Q, H, W); break;
        case 6: pmax = dense_passes<6>(geo, taps_cur, P, Q, H, W); break;
        case 7: pmax = dense_passes<7>(geo, taps_cur, P, Q, H, W); break;
        case 8: pmax = dense_passes<8>(geo, taps_cur, P, Q, H, W); break;
        default: pmax = dense_passes<9>(geo, taps_cur, P, Q, H, W); break;
      }
      pmax = warp_max(pmax);
      if (lane == 0) sh.red_max[warp] = pmax;
      __syncthreads();
      pmax = sh.red_max[0];
#pragma unroll
      for (int w = 1; w < kDWarps; ++w) pmax = fmaxf(pmax, sh.red_max[w]);
      const float amax = fmaxf(fabsf(vmax), fabsf(vmin));
      const float gamma = static_cast<float>(2 * d + 8) * 1.1920929e-7f;   // (2d + 8) * 2^-23
      const float thr = pmax - (2.0f * gamma + 4.0f * 5.9604645e-8f) * amax;

      // ---- E: candidates
      {
        const int Wp4 = geo.Wp >> 2;
        for (int i = tid; i < H * Wp4; i += kDThreads) {
          const float4 v = reinterpret_cast<const float4*>(P)[i];
          if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) >= thr) {
            const int y = fast_div(i, geo.div_Wp4), x = (i - y * Wp4) * 4;
            const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (x + j < W && f[j] >= thr) {
                const int slot = atomicAdd(&sh.cand_count, 1);
                if (slot < kMaxCand) sh.cand[slot] = y * W + x + j;
              }
            }
          }
        }
      }
      __syncthreads();
      const int count = sh.cand_count;

      // ---- F: exact values.  Up to five pixels at once: group g (25 threads) strides the taps of pixel ev_idx[g].
      auto run_evals = [&](int n) {
        __syncthreads();   // sh.ev_idx is visible
        const int g = tid / kFGroup, gl = tid - g * kFGroup;
        if (g < n) {
          const int idx = sh.ev_idx[g], y = fast_div(idx, geo.div_W), x = idx - y * W;
          const int nt = d * d, q25 = kFGroup / d, r25 = kFGroup - q25 * d;
          int ti = gl / d, tj = gl - ti * d;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll 1
          for (int i = gl; i < nt; i += 2 * kFGroup) {
            a0 = fma(w2d_cur[i], static_cast<double>(plane_value<T>(plane, reflect1(y + ti - r, H) * W + reflect1(x + tj - r, W))), a0);
            tj += r25; ti += q25;
            if (tj >= d) { tj -= d; ++ti; }
            if (i + kFGroup < nt)
              a1 = fma(w2d_cur[i + kFGroup], static_cast<double>(plane_value<T>(plane, reflect1(y + ti - r, H) * W + reflect1(x + tj - r, W))), a1);
            tj += r25; ti += q25;
            if (tj >= d) { tj -= d; ++ti; }
          }
          ev_part[tid] = a0 + a1;
        }
        __syncthreads();
        if (tid < n) {
          double sum = 0.0;
#pragma unroll 5
          for (int j = 0; j < kFGroup; ++j) sum += ev_part[tid * kFGroup + j];
          sh.ev_val[tid] = static_cast<float>(sum);
        }
        __syncthreads();
      };
      auto set_neighbours = [&](int first_slot, int by, int bx) {   // left, right, up, down
        if (tid < 4) {
          const int dx = (tid == 0) ? -1 : (tid == 1) ? 1 : 0;
          const int dy = (tid == 2) ? -1 : (tid == 3) ? 1 : 0;
          sh.ev_idx[first_slot + tid] = (by + dy) * W + bx + dx;
        }
      };

      bool have_nb = false;
      if (count == 1) {
        best = sh.cand[0];
        const int by = fast_div(best, geo.div_W), bx = best - by * W;
        interior = bx > 0 && bx < W - 1 && by > 0 && by < H - 1;
        if (interior) {
          if (tid == 0) sh.ev_idx[0] = best;
          set_neighbours(1, by, bx);
          run_evals(5);
          best_val = sh.ev_val[0];
#pragma unroll
          for (int q = 0; q < 4; ++q) nb[q] = sh.ev_val[1 + q];
        }
        have_nb = true;
      } else if (count <= kMaxCand) {
        best_val = -INFINITY; best = 0x7fffffff;
        for (int base = 0; base < count; base += 5) {
          const int n = min(5, count - base);
          if (tid < n) sh.ev_idx[tid] = sh.cand[base + tid];
          run_evals(n);
          for (int j = 0; j < n; ++j) argmax_combine(best_val, best, sh.ev_val[j], sh.cand[base + j]);
        }
      } else {
        // plateaus / heavily quantised maps: every warp walks its share of the prefilter plane
        ExactCtx<T> ectx;
        ectx.plane = plane; ectx.tile = nullptr; ectx.w2d = w2d_cur;
        ectx.H = H; ectx.W = W; ectx.r = r; ectx.d = d; ectx.oy0 = 0; ectx.ox0 = 0; ectx.tile_path = false;
        float wv = -INFINITY;
        int wi = 0x7fffffff;
        for (int base = warp * 32; base < HW; base += kDWarps * 32) {
          const int q = base + lane;
          const int y = fast_div(min(q, HW - 1), geo.div_W), x = min(q, HW - 1) - y * W;
          const bool want = q < HW && P[y * geo.Wp + x] >= thr;
          unsigned msk = __ballot_sync(0xffffffffu, want);
          while (msk) {
            const int b = __ffs(msk) - 1;
            msk &= msk - 1;
            const int yy = __shfl_sync(0xffffffffu, y, b), xx = __shfl_sync(0xffffffffu, x, b);
            argmax_combine(wv, wi, exact_eval<T>(ectx, yy, xx), yy * W + xx);
          }
        }
        if (lane == 0) { sh.red_max[warp] = wv; sh.red_i[warp] = wi; }
        __syncthreads();
        best_val = sh.red_max[0]; best = sh.red_i[0];
#pragma unroll
        for (int w = 1; w < kDWarps; ++w) argmax_combine(best_val, best, sh.red_max[w], sh.red_i[w]);
      }
      const int by = fast_div(best, geo.div_W), bx = best - by * W;
      if (!have_nb) {
        interior = bx > 0 && bx < W - 1 && by > 0 && by < H - 1;
        if (interior) {
          set_neighbours(0, by, bx);
          run_evals(4);
#pragma unroll
          for (int q = 0; q < 4; ++q) nb[q] = sh.ev_val[q];
        }
      }
      score = plane_value<T>(plane, best);
    }

    // ---- outputs: float32 sub-pixel fit in the reference's operation order (heatmap.py:136-165)
    if (tid == 0 || tid == 32) {
      const int by = fast_div(best, geo.div_W), bx = best - by * W;
      const bool is_y = tid == 32;
      float f = static_cast<float>(is_y ? by : bx);
      if (interior) {
        const float lo = is_y ? nb[2] : nb[0], hi = is_y ? nb[3] : nb[1], c = best_val;
        const float g = __fdiv_rn(__fsub_rn(hi, lo), 2.0f);
        float h = __fsub_rn(__fadd_rn(hi, lo), __fmul_rn(2.0f, c));
        if (h == 0.0f) h = 1e-6f;
        f = __fadd_rn(f, __fdiv_rn(-g, h));
      }
      locs[hm * 2 + (is_y ? 1 : 0)] = f;
      if (keypoints)
        keypoints[hm * 2 + (is_y ? 1 : 0)] =
            static_cast<double>(f) / static_cast<double>(is_y ? H - 1 : W - 1) * (is_y ? p.input_h : p.input_w);
      if (is_y) {
        vals[hm] = score;
        if (argmax) argmax[hm] = best;
      }
    }
    if (new_tables) {   // park the next channel's tables in the other buffer (nobody reads it during this iteration)
      double* w2d_nxt = w2d + (tb ^ 1) * (PP_OKS_TAPS * PP_OKS_TAPS);
      w2d_nxt[tid] = w_next[0];
      if (tid + kDThreads < PP_OKS_TAPS * PP_OKS_TAPS) w2d_nxt[tid + kDThreads] = w_next[1];
      if (tid < PP_OKS_TAPS) sh.taps[tb ^ 1][tid] = tid < 2 * r_next + 1 ? tap_next : 0.0f;
      tb ^= 1;
      r = r_next;
      d = 2 * r + 1;
    }
    k = k_next;
    hm = q_hm[s ^ 1];                   // item it+1 (written one iteration ago)
    if (tid == 0) publish(s, ticket);   // item it+2 takes over this item's slot; its copy starts after the barrier
    __syncthreads();                    // planes, tables and raw stage s are free
  }
}

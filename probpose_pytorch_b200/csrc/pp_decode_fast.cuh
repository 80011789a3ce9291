// Main kernel of the expected-OKS decoder (included into pp_decode.cu, inside its anonymous namespace).
//
// Exact pruning.  The full separable convolution costs ~2 x 13 FFMA per pixel on COCO sigmas -- more
// than the ~2100 warp instructions per 64x48 heatmap that the HBM roofline allows.  But the OKS kernel
// is non-negative and sums to 1, hence for every pixel p
//        R(p) = float32(sum_q w(q) h(p+q))  <=  max over the kernel window of h,
// and    max_p R(p) >= R(p0) =: L            for any pixel p0 (we take the raw maximum).
// Therefore the argmax of R -- and every pixel tied with it -- lies within `radius` of S = {h >= L}.
// For blob-shaped heatmaps S is a handful of pixels around the peak.
//
// Per heatmap, one persistent CTA of 128 threads:
//   (A) the contiguous plane arrives in shared memory by one TMA bulk copy (issued during the previous
//       heatmap); one vector scan gives max / first argmax / min.  Constant maps finish here.
//   (B) L = exact R(p0): double accumulation of the reference's d x d table, float32 result.
//   (C) second vector scan: bounding box of S.
//   tile path (box dilated by the radius fits 24 x 24):
//   (D) gather the dilated box + reflect halo into a padded tile, head tail applied; the plane is then
//       free and the next heatmap's TMA is issued;  (E/F) separable float32 prefilter on the tile only;
//   full path (flat / noisy / multi-modal maps):
//   (D') column pass straight from the plane (reflect in y by row index) into a padded float32 plane
//       whose reflect columns are written by the producing thread; (E') row pass with 128-bit windows;
//       only the per-task maxima are kept, tasks that can hold a candidate are recomputed;
//   (G) exact re-evaluation (double, reference table) of the pixels within the prefilter's rigorous
//       error band of its maximum, NumPy tie-break, and of the winner's four neighbours;
//   (H) float32 sub-pixel fit in the reference's operation order, outputs.
//
// The filter passes are written once for all radii (taps zero-padded to a multiple of 8, 8 outputs per
// task, 16-register sliding window): per-radius template instances blew the instruction cache
// (profiles/r01c: stall_no_inst dominated every line).
#pragma once

// Developer-only phase timing (build with -DPP_PHASE_TIMING): thread 0 of every CTA accumulates the
// cycles between phase marks into a device array that tools/decode_phases.py reads back.
#ifdef PP_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PP_MARK(slot)                                                         \
  do {                                                                        \
    if (threadIdx.x == 0) {                                                   \
      const long long now_ = clock64();                                       \
      atomicAdd(&g_phase_cycles[slot], static_cast<unsigned long long>(now_ - mark_)); \
      mark_ = now_;                                                           \
    }                                                                         \
  } while (0)
#else
#define PP_MARK(slot) do { } while (0)
#endif

constexpr int kFThreads = 128;
constexpr int kFWarps = kFThreads / 32;
constexpr int kFRegion = 24;                    // largest output region side of the tile path
constexpr int kFMarg = 12;                      // column margin: >= radius + 1, multiple of 4
constexpr int kFTileStride = 60;                // >= 12 + 24 + 12 + 8, multiple of 4 with odd quarter
constexpr int kFTileRows = kFRegion + 2 * (PP_MAX_OKS_RADIUS + 1);   // 44
constexpr int kFTmpRows = kFRegion + 2 * PP_MAX_OKS_RADIUS;          // 42
constexpr int kFTmpStride = 28;                 // 24 -> multiple of 4 with odd quarter
constexpr int kFTileFloats = kFTileRows * kFTileStride + (kFTmpRows + 8) * kFTmpStride;
constexpr int kFTaps = 24;                      // taps padded to 3 chunks of 8
constexpr int kFMaxChannels = 2048;
constexpr int kFGroup = 25;                     // threads per exact evaluation (5 evaluations at once)             // per-channel scheduling uses the work area as scratch

struct FastGeom {
  unsigned plane_bytes;   // H*W*sizeof(T)
  unsigned work_off;      // byte offset of the float work area (tile + tmp | padded full plane)
  unsigned work_floats;
  unsigned taskmax_off;   // byte offset of the per-task maxima (full path)
  unsigned w2d_off;       // byte offset of the staged d x d table (doubles)
  int full_stride;        // row stride of the padded full plane (floats)
  int W8;                 // W rounded up to 8
  unsigned div_WV, div_W, div_H;   // reciprocals for the index decompositions
};

struct FastShared {
  float red_f[2][kFWarps];
  int red_i[kFWarps];
  __align__(16) float grow[kFTaps];   // row-pass taps, shifted so that the window starts 16-byte aligned
  __align__(16) float gcol[kFTaps];   // column-pass taps
  float nb[4];
  float L;
  int p0;
  int ev_idx[5];      // pixels handed to the grouped exact evaluation
  float ev_val[5];
  int bbox[4];        // min x, max x, min y, max y of S
  int cand[kMaxCand];
  int cand_count;
};

// n / d for small n, d via a precomputed reciprocal (exact for n * d < 2^31).  d == 1 has no 32-bit reciprocal
// (0xFFFFFFFF / 1 + 1 wraps to 0): magic 0 means "divide by one".  Round 1 shipped without that case and the team
// kernel split the row tasks of a band with a single row pair wrongly (q = t, xb = 0): a convolved maximum in the last
// one or two rows of a region whose height is 1 or 2 modulo the band height was missed -- 6 of 408 576 heatmaps of
// C5, found by comparing against the tensor-core kernel on whole batches (tests: *_maximum_in_a_short_last_band).
__host__ __device__ inline unsigned div_magic(unsigned d) { return d <= 1u ? 0u : 0xFFFFFFFFu / d + 1u; }
__device__ __forceinline__ int fast_div(int n, unsigned magic) {
  return magic ? static_cast<int>(__umulhi(static_cast<unsigned>(n), magic)) : n;
}

template <typename T>
__device__ __forceinline__ float plane_value(const T* plane, int idx) {
  return Elem<T>::to_f32(plane[idx]);
}

// clamp(x / temperature, 0, 1) rounded like torch does for the tensor dtype; out of line: the IEEE
// division expands to a long sequence and this is called from one place only.
template <typename T>
__device__ __noinline__ float tail_value(float v, float temperature) {
  return apply_tail<T>(v, true, temperature);
}

// single reflection, valid for -n <= i < 2n (the kernel requires radius < min(H, W))
__device__ __forceinline__ int reflect1(int i, int n) {
  i = i < 0 ? -1 - i : i;
  return i >= n ? 2 * n - 1 - i : i;
}

__device__ __forceinline__ void ld8(const float* p, float* v) {   // 32-byte aligned shared address
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// acc[o] = sum_{j < 8 nch} g[j] * src[o + j], o = 0..7; src 16-byte aligned and contiguous
__device__ __forceinline__ void conv8_contiguous(const float* __restrict__ src, const float* __restrict__ g, int nch,
                                                 float (&acc)[8]) {
  float w[16];
  ld8(src, w);
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.0f;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    ld8(src + 8 * c + 8, w + 8);
    float t[8];
    ld8(g + 8 * c, t);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = fmaf(t[jj], w[o + jj], acc[o]);
#pragma unroll
    for (int o = 0; o < 8; ++o) w[o] = w[o + 8];
  }
}

// same with the inputs fetched one by one through `load(j)` (strided / reflected / converted sources)
template <typename Load>
__device__ __forceinline__ void conv8_gather(Load load, const float* __restrict__ g, int nch, float (&acc)[8]) {
  float w[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = load(j);
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.0f;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) w[8 + j] = load(8 * c + 8 + j);
    float t[8];
    ld8(g + 8 * c, t);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = fmaf(t[jj], w[o + jj], acc[o]);
#pragma unroll
    for (int o = 0; o < 8; ++o) w[o] = w[o + 8];
  }
}

template <typename T>
__device__ __forceinline__ void load2(const T* p, float& a, float& b);
template <>
__device__ __forceinline__ void load2<float>(const float* p, float& a, float& b) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  a = v.x; b = v.y;
}
template <>
__device__ __forceinline__ void load2<__nv_bfloat16>(const __nv_bfloat16* p, float& a, float& b) {
  const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
  a = __uint_as_float(v << 16); b = __uint_as_float(v & 0xffff0000u);
}

// Full-path column pass for two adjacent columns (x even) and 4 rows, taps in chunks of 4:
//   full[y][kFMarg + x] = sum_j tap[j] h[reflect(y - r + j)][x];
// columns within r of an edge are mirrored into the pad by the thread that produces them.
template <typename T, bool kInside, bool kZeroPad = false>
__device__ __forceinline__ void full_col_task(const T* __restrict__ plane, float* __restrict__ work,
                                              const float* __restrict__ g, int nch4, int x, int y0, int r, int H, int W,
                                              int FS) {
  float w0[8], w1[8], a0[4], a1[4];
  const int need = 4 + 2 * r;
  const T* base = plane + x + (kInside ? (y0 - r) * W : 0);
  auto fetch = [&](int j, float& u, float& v) {
    if (j < need) {
      if (kInside) {
        load2<T>(base + j * W, u, v);
      } else if (kZeroPad) {   // zero padding (the DARK blur): rows outside the map contribute nothing
        const int yy = y0 - r + j;
        if (yy >= 0 && yy < H) load2<T>(base + yy * W, u, v);
        else { u = 0.0f; v = 0.0f; }
      } else {
        load2<T>(base + reflect1(y0 - r + j, H) * W, u, v);
      }
    } else {
      u = 0.0f; v = 0.0f;
    }
  };
#pragma unroll
  for (int j = 0; j < 4; ++j) fetch(j, w0[j], w1[j]);
#pragma unroll
  for (int o = 0; o < 4; ++o) { a0[o] = 0.0f; a1[o] = 0.0f; }
#pragma unroll 1
  for (int c = 0; c < nch4; ++c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) fetch(4 * c + 4 + j, w0[4 + j], w1[4 + j]);
    const float4 t4 = *reinterpret_cast<const float4*>(g + 4 * c);
    const float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        a0[o] = fmaf(t[jj], w0[o + jj], a0[o]);
        a1[o] = fmaf(t[jj], w1[o + jj], a1[o]);
      }
#pragma unroll
    for (int o = 0; o < 4; ++o) { w0[o] = w0[o + 4]; w1[o] = w1[o + 4]; }
  }
  const bool left = !kZeroPad && x < r, left1 = !kZeroPad && x + 1 < r;
  const bool right = !kZeroPad && x >= W - r, right1 = !kZeroPad && x + 1 >= W - r;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (y0 + o < H) {
      float* row = work + (y0 + o) * FS + kFMarg;
      *reinterpret_cast<float2*>(row + x) = make_float2(a0[o], a1[o]);
      if (left) row[-1 - x] = a0[o];
      if (left1) row[-2 - x] = a1[o];
      if (right) row[2 * W - 1 - x] = a0[o];
      if (right1) row[2 * W - 2 - x] = a1[o];
    }
  }
}

// exact value of one convolved pixel: double accumulation of the staged d x d table, float32 result
// (what scipy stores, heatmap.py:362-364).  Warp-collective; lanes stride the flattened taps, two
// accumulators shorten the dependency chain.  One out-of-line copy serves every call site.
template <typename T>
struct ExactCtx {
  const T* plane;        // whole heatmap (reflect by index) ...
  const float* tile;     // ... or the gathered tile with its halo materialised
  const double* w2d;
  int H, W, r, d, oy0, ox0;
  bool tile_path;
};

template <typename T>
__device__ __noinline__ float exact_eval(const ExactCtx<T>& c, int y, int x) {
  const int lane = threadIdx.x & 31, d = c.d, n = d * d;
  const int q32 = 32 / d, r32 = 32 - q32 * d;   // 32 = q32 * d + r32
  int ti = lane / d, tj = lane - ti * d;
  double a0 = 0.0, a1 = 0.0;
  if (c.tile_path) {
    const float* base = c.tile + (y - c.oy0 + 1) * kFTileStride + (x - c.ox0 + kFMarg - c.r);
#pragma unroll 1
    for (int i = lane; i < n; i += 64) {
      a0 = fma(c.w2d[i], static_cast<double>(base[ti * kFTileStride + tj]), a0);
      tj += r32; ti += q32;
      if (tj >= d) { tj -= d; ++ti; }
      if (i + 32 < n) a1 = fma(c.w2d[i + 32], static_cast<double>(base[ti * kFTileStride + tj]), a1);
      tj += r32; ti += q32;
      if (tj >= d) { tj -= d; ++ti; }
    }
  } else {
#pragma unroll 1
    for (int i = lane; i < n; i += 64) {
      a0 = fma(c.w2d[i], static_cast<double>(plane_value<T>(c.plane, reflect1(y + ti - c.r, c.H) * c.W + reflect1(x + tj - c.r, c.W))), a0);
      tj += r32; ti += q32;
      if (tj >= d) { tj -= d; ++ti; }
      if (i + 32 < n)
        a1 = fma(c.w2d[i + 32], static_cast<double>(plane_value<T>(c.plane, reflect1(y + ti - c.r, c.H) * c.W + reflect1(x + tj - c.r, c.W))), a1);
      tj += r32; ti += q32;
      if (tj >= d) { tj -= d; ++ti; }
    }
  }
  return static_cast<float>(warp_sum(a0 + a1));
}

// More near-maximal pixels than the candidate list holds (plateaus / heavily quantised maps): every warp
// walks its share of the region (tile path) or plane (full path), recomputes the prefilter value with the
// same operation order as the passes and re-evaluates the qualifying pixels exactly.  Rare; out of line.
struct OverflowCtx {
  const float* src;      // second-pass input: tmp (tile path, taps along rows) or padded plane (full path)
  const float* taps;
  int step, row_stride;  // tap stride / row stride in `src`
  int x_lo, y_lo, x_n, y_n;
  float thr;
};

template <typename T>
__device__ __noinline__ void overflow_scan(const OverflowCtx& oc, const ExactCtx<T>& ectx, float& wv, int& wi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned magic = div_magic(oc.x_n);
  for (int base = warp * 32; base < oc.x_n * oc.y_n; base += kFWarps * 32) {
    const int q = base + lane;
    bool want = false;
    int y = 0, x = 0;
    if (q < oc.x_n * oc.y_n) {
      const int qy = fast_div(q, magic), qx = q - qy * oc.x_n;
      y = oc.y_lo + qy; x = oc.x_lo + qx;
      const float* c = oc.src + (ectx.tile_path ? qy : y) * oc.row_stride + (ectx.tile_path ? qx : x);
      float a = 0.0f;
      for (int j = 0; j < ectx.d; ++j) a = fmaf(oc.taps[j], c[j * oc.step], a);
      want = a >= oc.thr;
    }
    unsigned msk = __ballot_sync(0xffffffffu, want);
    while (msk) {
      const int b = __ffs(msk) - 1;
      msk &= msk - 1;
      const int yy = __shfl_sync(0xffffffffu, y, b), xx = __shfl_sync(0xffffffffu, x, b);
      argmax_combine(wv, wi, exact_eval<T>(ectx, yy, xx), yy * ectx.W + xx);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kFThreads, 6)
decode_expected_fast_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps,
                            float* __restrict__ locs, float* __restrict__ vals, int32_t* __restrict__ argmax,
                            double* __restrict__ keypoints, FastGeom geo, unsigned* __restrict__ work_counter) {
  extern __shared__ __align__(128) unsigned char fsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ FastShared sh;

  const T* plane = reinterpret_cast<const T*>(fsm);
  T* plane_rw = reinterpret_cast<T*>(fsm);
  float* work = reinterpret_cast<float*>(fsm + geo.work_off);
  float* tile = work;
  float* tmp = work + kFTileRows * kFTileStride;
  float* task_max = reinterpret_cast<float*>(fsm + geo.taskmax_off);
  double* w2d = reinterpret_cast<double*>(fsm + geo.w2d_off);
  // partial sums of the grouped exact evaluation: the per-task maxima are dead by then, reuse their space
  double* ev_part = reinterpret_cast<double*>(fsm + geo.taskmax_off);

  constexpr int V = Elem<T>::kVec;
  const int H = p.H, W = p.W, HW = H * W, WV = W / V, FS = geo.full_stride;
  const int64_t N = static_cast<int64_t>(p.B) * p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tail = p.apply_tail != 0;
  const float temp = p.temperature;
  // (row, vector-in-row) of this thread's first vector and the per-iteration step of the scans
  const int step_y = fast_div(kFThreads, geo.div_WV), step_x = kFThreads - step_y * WV;
  const int first_y = fast_div(tid, geo.div_WV), first_x = tid - first_y * WV;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  // stale work-area contents only ever meet zero taps or unused outputs, but they must be finite
  for (int i = tid; i < static_cast<int>(geo.work_floats); i += kFThreads) work[i] = 0.0f;
  __syncthreads();
  // ---- work distribution.  Heatmaps cost between ~4 ns (constant maps) and ~25 ns (flat maps, wide
  // kernels) each and a CTA only sees a handful of them, so a static split leaves the slowest CTA ~1.5x
  // behind the average.  CTAs therefore pull work items one at a time from a global counter.  Items are
  // numbered channel-major (item j -> channel j / B, image j % B): a CTA changes channel at most a few
  // times in its life, so the per-channel tables are staged almost once.  Without a counter (no scratch)
  // the heatmaps are strided statically.
  __shared__ long long next_item, next_hm;
  const bool dynamic = work_counter != nullptr;
  auto item_to_hm = [&](long long j) -> long long {   // thread 0 only (64-bit division)
    if (!dynamic) return j;
    const long long slot = j / p.B, b = j - slot * p.B;
    const long long kk = tab.order ? tab.order[slot] : slot;   // widest kernels first
    return b * p.K + kk;
  };
  auto pull = [&]() -> long long {   // thread 0 only
    return dynamic ? static_cast<long long>(atomicAdd(work_counter, 1u)) : -1;
  };
  if (tid == 0) {
    next_item = dynamic ? pull() : static_cast<long long>(blockIdx.x);
    next_hm = next_item < N ? item_to_hm(next_item) : N;
    if (next_item < N) {
      mbar_expect_tx(&bar, geo.plane_bytes);
      tma_load_1d(fsm, heatmaps + next_hm * HW, geo.plane_bytes, &bar);
    }
  }
  __syncthreads();
  long long item = next_item;
  int64_t hm = next_hm;

  int k_loaded = -1, r = 1, d = 3, padr = 4, nch_row = 1, nch_col = 1;

#ifdef PP_PHASE_TIMING
  long long mark_ = clock64();
#endif
  for (int it = 0; item < N; ++it) {
    const int k = static_cast<int>(hm % p.K);
    if (k != k_loaded) {
      __syncthreads();   // nobody still reads the previous tables
      r = tab.radius[k];
      d = 2 * r + 1;
      padr = (r + 3) & ~3;
      const int shift = padr - r;
      nch_row = (shift + d + 7) >> 3;
      nch_col = (d + 7) >> 3;
      const double* src = tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;
      for (int i = tid; i < d * d; i += kFThreads) w2d[i] = src[i];
      const float* t1 = tab.taps_f32 + k * PP_OKS_TAPS;
      if (tid < kFTaps) {
        sh.gcol[tid] = tid < d ? t1[tid] : 0.0f;
        sh.grow[tid] = (tid >= shift && tid < shift + d) ? t1[tid - shift] : 0.0f;
      }
      k_loaded = k;
    }
    if (tid == 0) {
      sh.bbox[0] = W; sh.bbox[1] = -1; sh.bbox[2] = H; sh.bbox[3] = -1;
      sh.cand_count = 0;
    }
    PP_MARK(0);   // loop top / tables
    mbar_wait(&bar, it & 1);
    PP_MARK(1);   // TMA wait

    // ---- A: one scan of the plane: max and min (the index of a maximum is looked up afterwards by the
    // threads that hold it).  The head tail (head.py:526-532), when requested, is applied first, once,
    // in place, so that every later phase reads plain values.
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
    const float tmax = xmax;   // this thread's own maximum
    xmax = warp_max(xmax);
    xmin = -warp_max(-xmin);
    if (lane == 0) { sh.red_f[0][warp] = xmax; sh.red_f[1][warp] = xmin; }
    if (tid == 0) sh.p0 = 0x7fffffff;
    __syncthreads();
    xmax = fmaxf(fmaxf(sh.red_f[0][0], sh.red_f[0][1]), fmaxf(sh.red_f[0][2], sh.red_f[0][3]));
    xmin = fminf(fminf(sh.red_f[1][0], sh.red_f[1][1]), fminf(sh.red_f[1][2], sh.red_f[1][3]));
    const float vmax = xmax, vmin = xmin;
    PP_MARK(2);   // scan A

    // every flag below is uniform across the CTA
    const bool constant = vmax == vmin;   // e.g. all zeros after the clamp: first index wins, border pixel
    bool tile_path = false;
    int best = 0;
    float best_val = 0.0f, score = vmax;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    bool interior = false;
    int ox0 = 0, oy0 = 0, OW = 0, OH = 0;

    ExactCtx<T> ectx;
    ectx.plane = plane; ectx.tile = tile; ectx.w2d = w2d;
    ectx.H = H; ectx.W = W; ectx.r = r; ectx.d = d; ectx.oy0 = 0; ectx.ox0 = 0;
    ectx.tile_path = false;

    if (!constant) {
      // ---- B: the raw maximum p0 and a lower bound L of the convolved maximum
      if (tmax == xmax) {   // lowest index among the pixels that hold the maximum
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
      const int imax = sh.p0;
      const int py = fast_div(imax, geo.div_W), px = imax - py * W;
      if (warp == 0) {
        // L: a lower bound of R(p0), hence of max R, from the central 5 x 5 taps and "everything else is
        // at least vmin" (the taps are non-negative and sum to 1):  sum_25 w h + (1 - sum_25 w) vmin.
        // For the narrowest kernels (d = 5) this is R(p0) itself.
        const int half = min(2, r), side = 2 * half + 1;
        double sw = 0.0, swh = 0.0;
        if (lane < side * side) {
          const int ti = lane / side, tj = lane - ti * side;
          const double w = w2d[(r - half + ti) * d + (r - half + tj)];
          const float v = plane_value<T>(plane, reflect1(py - half + ti, H) * W + reflect1(px - half + tj, W));
          sw = w; swh = w * static_cast<double>(v);
        }
        sw = warp_sum(sw); swh = warp_sum(swh);
        if (lane == 0) {
          const float e = static_cast<float>(swh + fmax(1.0 - sw, 0.0) * static_cast<double>(xmin));
          sh.L = e - fabsf(e) * 1e-6f - 1e-37f;
        }
      }
      __syncthreads();
      const float L = sh.L;
      PP_MARK(3);   // B: p0 + exact L

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
          if (m >= L) {   // rare on blob-shaped maps
#pragma unroll
            for (int j = 0; j < V; ++j) {
              if (f[j] >= L) {
                bx0 = min(bx0, xv * V + j);
                bx1 = max(bx1, xv * V + j);
              }
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
        if (holders != 0u && __popc(holders) <= 6) {   // blob-shaped maps: a few lanes publish directly
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
      // p0 itself is in S (h(p0) = max >= R(p0)), so the box is never empty
      ox0 = max(sh.bbox[0] - r, 0);
      oy0 = max(sh.bbox[2] - r, 0);
      OW = min(sh.bbox[1] + r, W - 1) - ox0 + 1;
      OH = min(sh.bbox[3] + r, H - 1) - oy0 + 1;
      tile_path = OW <= kFRegion && OH <= kFRegion;
      PP_MARK(4);   // C: bounding box
    }

    if (!constant && tile_path) {
      // ---- D: gather the region (+ radius + 1 halo, reflect-extended, tail applied) into the tile
      const int rows = OH + 2 * (r + 1);
      const int c_lo = kFMarg - (r + 1), c_hi = kFMarg + OW + r;   // inclusive
      // flattened over (row, column) so that every thread has several independent copies in flight
      const int ncols = c_hi - c_lo + 1, total = rows * ncols;
      const unsigned mcols = div_magic(ncols);
#pragma unroll 4
      for (int e = tid; e < total; e += kFThreads) {
        const int ty = fast_div(e, mcols), c = c_lo + (e - ty * ncols);
        const int yy = reflect1(oy0 - (r + 1) + ty, H), xx = reflect1(ox0 - kFMarg + c, W);
        tile[ty * kFTileStride + c] = plane_value<T>(plane, yy * W + xx);
      }
    } else if (!constant) {
      // ---- D': full path, column pass from the plane into the padded float32 plane:
      // full[y][kFMarg + x] = sum_j tap[j] h[reflect(y - r + j)][x]; columns within r of an edge are
      // mirrored into the pad by the thread that produces them.
      const int yblocks = (H + 3) >> 2, W2 = W >> 1, nch4 = (d + 3) >> 2;
      const unsigned mW2 = div_magic(W2);
      for (int t = tid; t < W2 * yblocks; t += kFThreads) {
        const int yb = fast_div(t, mW2), x = (t - yb * W2) * 2;
        const int y0 = yb * 4;
        if (y0 - r >= 0 && y0 + 3 + r < H)
          full_col_task<T, true>(plane, work, sh.gcol, nch4, x, y0, r, H, W, FS);
        else
          full_col_task<T, false>(plane, work, sh.gcol, nch4, x, y0, r, H, W, FS);
      }
    }
    __syncthreads();
    PP_MARK(5);   // D: gather / column pass
    const bool plane_free = constant || tile_path;   // the full path still needs the plane for step G
    auto fetch_next = [&]() {   // thread 0: claim the next work item and start its copy
      const long long j = dynamic ? pull() : item + static_cast<long long>(gridDim.x);
      const long long h = j < N ? item_to_hm(j) : N;
      next_item = j;
      next_hm = h;
      if (j < N) {
        fence_proxy_async();
        mbar_expect_tx(&bar, geo.plane_bytes);
        tma_load_1d(fsm, heatmaps + h * HW, geo.plane_bytes, &bar);
      }
    };
    if (plane_free && tid == 0) fetch_next();   // the next heatmap streams in while this one is finished

    if (!constant) {
      // ---- E/F: separable float32 prefilter
      float pmax = -INFINITY;
      float cv[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) cv[o] = -INFINITY;
      int cx = 0, cyb = 0;
      bool have_cv = false;
      if (tile_path) {
        // rows of the tile: tmp[ty][x] = sum_j tap[j] tile[ty + 1][kFMarg + x - r + j]
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
        // columns: conv(x, y) = sum_j tap[j] tmp[y + j][x]; one task (8 rows) per thread, kept in registers
        const int yblocks = (OH + 7) >> 3;
        if (tid < OW * yblocks) {
          cyb = fast_div(tid, div_magic(OW)); cx = tid - cyb * OW;
          have_cv = true;
          const float* colp = tmp + cx;
          const int y0 = cyb * 8;
          float acc[8];
          conv8_gather([&](int j) -> float { return colp[min(y0 + j, kFTmpRows + 7) * kFTmpStride]; }, sh.gcol, nch_col,
                       acc);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            if (y0 + o < OH) { cv[o] = acc[o]; pmax = fmaxf(pmax, acc[o]); }
          }
        }
      } else {
        // rows of the padded plane, 8 outputs per task; only the per-task maximum is kept
        const int tasks = (geo.W8 >> 3) * H;
        for (int t = tid; t < tasks; t += kFThreads) {
          const int xb = fast_div(t, geo.div_H), y = t - xb * H;
          float acc[8];
          conv8_contiguous(work + y * FS + kFMarg + xb * 8 - padr, sh.grow, nch_row, acc);
          float m = -INFINITY;
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (xb * 8 + o < W) m = fmaxf(m, acc[o]);
          task_max[t] = m;
          pmax = fmaxf(pmax, m);
        }
      }
      pmax = warp_max(pmax);
      if (lane == 0) sh.red_f[0][warp] = pmax;
      __syncthreads();
      PP_MARK(6);   // E/F: prefilter
      pmax = fmaxf(fmaxf(sh.red_f[0][0], sh.red_f[0][1]), fmaxf(sh.red_f[0][2], sh.red_f[0][3]));
      const float amax = fmaxf(fabsf(vmax), fabsf(vmin));
      const float gamma = static_cast<float>(2 * d + 8) * 1.1920929e-7f;   // (2d + 8) * 2^-23
      const float thr = pmax - (2.0f * gamma + 4.0f * 5.9604645e-8f) * amax;

      // candidates: pixels whose prefilter value is within the error band of the maximum
      if (tile_path) {
        if (have_cv) {
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            if (cv[o] >= thr) {   // cv is -inf for rows beyond the region
              const int slot = atomicAdd(&sh.cand_count, 1);
              if (slot < kMaxCand) sh.cand[slot] = (oy0 + cyb * 8 + o) * W + ox0 + cx;
            }
          }
        }
      } else {
        const int tasks = (geo.W8 >> 3) * H;
        for (int t = tid; t < tasks; t += kFThreads) {
          if (!(task_max[t] >= thr)) continue;
          const int xb = fast_div(t, geo.div_H), y = t - xb * H;
          float acc[8];
          conv8_contiguous(work + y * FS + kFMarg + xb * 8 - padr, sh.grow, nch_row, acc);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            if (xb * 8 + o < W && acc[o] >= thr) {
              const int slot = atomicAdd(&sh.cand_count, 1);
              if (slot < kMaxCand) sh.cand[slot] = y * W + xb * 8 + o;
            }
          }
        }
      }
      __syncthreads();
      const int count = sh.cand_count;
      PP_MARK(7);   // candidates

      // ---- G: exact values (double accumulation of the reference's d x d table, float32 result) of the
      // candidates and of the winner's four neighbours.  Up to five pixels are evaluated at once: group g
      // (25 threads) strides the taps of pixel sh.ev_idx[g]; the partial sums meet in shared memory and
      // are added in a fixed order.
      ectx.tile_path = tile_path; ectx.oy0 = oy0; ectx.ox0 = ox0;
      auto run_evals = [&](int n) {
        __syncthreads();   // sh.ev_idx is visible
        const int g = tid / kFGroup, gl = tid - g * kFGroup;
        if (g < n) {
          const int idx = sh.ev_idx[g], y = fast_div(idx, geo.div_W), x = idx - y * W;
          const int nt = d * d, q25 = kFGroup / d, r25 = kFGroup - q25 * d;
          int ti = gl / d, tj = gl - ti * d;
          double a0 = 0.0, a1 = 0.0;
          if (tile_path) {
            const float* base = tile + (y - oy0 + 1) * kFTileStride + (x - ox0 + kFMarg - r);
#pragma unroll 1
            for (int i = gl; i < nt; i += 2 * kFGroup) {
              a0 = fma(w2d[i], static_cast<double>(base[ti * kFTileStride + tj]), a0);
              tj += r25; ti += q25;
              if (tj >= d) { tj -= d; ++ti; }
              if (i + kFGroup < nt) a1 = fma(w2d[i + kFGroup], static_cast<double>(base[ti * kFTileStride + tj]), a1);
              tj += r25; ti += q25;
              if (tj >= d) { tj -= d; ++ti; }
            }
          } else {
#pragma unroll 1
            for (int i = gl; i < nt; i += 2 * kFGroup) {
              a0 = fma(w2d[i], static_cast<double>(plane_value<T>(plane, reflect1(y + ti - r, H) * W + reflect1(x + tj - r, W))), a0);
              tj += r25; ti += q25;
              if (tj >= d) { tj -= d; ++ti; }
              if (i + kFGroup < nt)
                a1 = fma(w2d[i + kFGroup], static_cast<double>(plane_value<T>(plane, reflect1(y + ti - r, H) * W + reflect1(x + tj - r, W))), a1);
              tj += r25; ti += q25;
              if (tj >= d) { tj -= d; ++ti; }
            }
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
        // the usual case: a single pixel can be the maximum; its value and its neighbours in one round
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
        float wv = -INFINITY;
        int wi = 0x7fffffff;
        OverflowCtx oc;
        oc.src = tile_path ? tmp : work + kFMarg - r;
        oc.step = tile_path ? kFTmpStride : 1;
        oc.row_stride = tile_path ? kFTmpStride : FS;
        oc.x_lo = tile_path ? ox0 : 0; oc.y_lo = tile_path ? oy0 : 0;
        oc.x_n = tile_path ? OW : W; oc.y_n = tile_path ? OH : H;
        oc.taps = sh.gcol; oc.thr = thr;
        overflow_scan<T>(oc, ectx, wv, wi);
        if (lane == 0) { sh.red_f[0][warp] = wv; sh.red_i[warp] = wi; }
        __syncthreads();
        best_val = sh.red_f[0][0]; best = sh.red_i[0];
#pragma unroll
        for (int w = 1; w < kFWarps; ++w) argmax_combine(best_val, best, sh.red_f[0][w], sh.red_i[w]);
      }
      PP_MARK(8);   // G: exact candidates
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
      score = tile_path ? tile[(by - oy0 + r + 1) * kFTileStride + bx - ox0 + kFMarg]
                        : plane_value<T>(plane, best);
    }

    PP_MARK(9);   // neighbours
    // ---- H: outputs.  The x and y halves of the sub-pixel fit (heatmap.py:136-165, float32, the
    // reference's operation order) are independent: thread 0 does x, thread 32 does y.
    if (tid == 0 || tid == 32) {
      const int by = fast_div(best, geo.div_W), bx = best - by * W;
      const bool is_y = tid == 32;
      float f = static_cast<float>(is_y ? by : bx);
      if (interior) {
        const float lo = is_y ? nb[2] : nb[0], hi = is_y ? nb[3] : nb[1], c = best_val;   // left/up, right/down
        const float g = __fdiv_rn(__fsub_rn(hi, lo), 2.0f);
        float h = __fsub_rn(__fadd_rn(hi, lo), __fmul_rn(2.0f, c));
        if (h == 0.0f) h = 1e-6f;
        f = __fadd_rn(f, __fdiv_rn(-g, h));
      }
      locs[hm * 2 + (is_y ? 1 : 0)] = f;
      if (keypoints)   // float32 / int -> float64, then * input_size (codec.py:237)
        keypoints[hm * 2 + (is_y ? 1 : 0)] =
            static_cast<double>(f) / static_cast<double>(is_y ? H - 1 : W - 1) * (is_y ? p.input_h : p.input_w);
      if (is_y) {
        vals[hm] = score;
        if (argmax) argmax[hm] = best;
      }
    }
    PP_MARK(10);  // H: outputs
    __syncthreads();   // work area / scratch / plane are reused by the next heatmap
    PP_MARK(11);  // final barrier
    if (!plane_free) {
      if (tid == 0) fetch_next();
      __syncthreads();   // publish next_item
    }
    item = next_item;
    hm = next_hm;
  }
}

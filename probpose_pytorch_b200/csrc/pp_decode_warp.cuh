// Expected-OKS decoder, one WARP (generally: a team of G warps) per heatmap.  Included into pp_decode.cu after
// pp_decode_fast.cuh, same namespace.
//
// Why.  The CTA-per-heatmap kernel (pp_decode_fast.cuh) spends most of its time waiting: ncu capture J shows
// 2.8 barrier-stall cycles per issued instruction, 1.5 no-instruction cycles (a 5.7 k-instruction kernel whose six
// resident CTAs sit in different phases) and 49 % issue utilisation; capping the resident CTAs (tools/
// decode_ctas_sweep.py: 199 / 114 / 89 / 78 / 72 / 69 us at 1..6 CTAs per SM) shows the curve has flattened, so more
// of the same does not help.  Here a heatmap belongs to ONE warp from its bulk copy to its outputs:
//   * no __syncthreads anywhere after the prologue, no cross-warp reduction, no single-thread section that idles
//     three other warps; warp-wide reductions are shuffles / redux;
//   * a heatmap costs its plane + a 4.6 KB band buffer of shared memory (17.9 KB for 64x48 float32), so twelve
//     heatmaps are in flight per SM instead of six;
//   * one code path for every map: the pruned region of the convolution (the neighbourhood of S = {h >= L}, see
//     pp_decode_fast.cuh for the bound) is processed in bands of row pairs -- column pass from the plane into the band
//     buffer (reflect in y through a CTA-wide row-offset table, reflected columns written by the producing lane), row
//     pass from the band buffer with 128-bit windows, both on two-wide FFMA2.  Blob-shaped maps need one small band,
//     flat / noisy maps four full-width ones; the wide kernels are prefiltered with their central taps only;
//   * candidates come from a warp-wide running maximum (one shuffle reduction per round of row tasks): a task
//     reports pixels only when it comes within the error band of it, into a 64-entry list that is filtered with the
//     final threshold; nothing but the band buffer and that list is ever stored;
//   * exact double-precision re-evaluation with the reference's d x d table, read through the read-only cache: the
//     winner and its four neighbours share every tap load (5 DFMA per tap per lane).
// The arithmetic of the prefilter (fmaf chains over zero-padded taps), the error band and the exact evaluation are
// those of pp_decode_fast.cuh; only the order in which partial double sums meet differs (any order is within
// 2^-50 of scipy's sum, far inside a float32 rounding cell).
// G = 2 (two warps share a heatmap through a named barrier) is used for planes so large that fewer than ten fit on an
// SM; at 64x48 it is slower than G = 1.  What was measured on the way, including everything that did not help, is in
// profiles/r01s_summary.md.
#pragma once

// developer-only phase timing (see pp_decode_fast.cuh): lane 0 of every team accumulates cycles between marks
#ifdef PP_PHASE_TIMING
#define PP_WMARK(slot)                                                        \
  do {                                                                        \
    if (tl == 0) {                                                            \
      const long long now_ = clock64();                                       \
      atomicAdd(&g_phase_cycles[slot], static_cast<unsigned long long>(now_ - mark_)); \
      mark_ = now_;                                                           \
    }                                                                         \
  } while (0)
#else
#define PP_WMARK(slot) do { } while (0)
#endif

constexpr int kWTmpPairs = 592;     // band buffer per team, in float2 (row pair) elements: 4.6 KB (8 full-width row pairs of a 64x48 map)
constexpr int kWCand = 64;          // candidate list per warp
constexpr int kWTaps = 24;          // 1-D taps zero-padded to 3 chunks of 8
constexpr int kWRowPad = 16;        // the reflected row-offset table covers rows -16 .. H + 15

struct WarpGeom {
  unsigned plane_bytes;   // H * W * sizeof(T)
  unsigned tmp_off;       // byte offsets inside a warp's shared-memory slot
  unsigned taps_off;      // two tables of kWTaps (t, t) pairs: taps at j, taps at j + 1
  unsigned cand_off;
  unsigned exch_off;      // TeamExchange
  unsigned slot_bytes;    // per-team slot (multiple of 128)
  unsigned rowoff_off;    // CTA-wide table of reflected row offsets, after the last slot
  unsigned full_taps;     // 1: prefilter with every tap (PP_DECODE_FULLTAPS=1, for A/B measurements and tests)
  unsigned div_WV, div_W;
};

// ---- two-wide float32 arithmetic (PTX fma.rn.f32x2 -> SASS FFMA2: two ordinary IEEE fused multiply-adds) ----
// Both filter passes are written on pairs: the column pass on (column x, column x + 1), which is how a 64-bit
// shared-memory load delivers the plane, the row pass on (row 2q, row 2q + 1), which is how the band buffer is laid
// out.  The warp is latency / issue bound, not FMA-pipe bound (capture P: FMA pipe 19 % busy), so halving the
// multiply-add instruction count is what counts.
typedef unsigned long long wf2;
__device__ __forceinline__ wf2 wf2_make(float lo, float hi) {
  wf2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void wf2_split(wf2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ wf2 wf2_fma(wf2 a, wf2 b, wf2 c) {
  wf2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <typename T>
__device__ __forceinline__ wf2 load_pair(const T* p);   // (p[0], p[1]) as float32 pair, p 2-element aligned
template <>
__device__ __forceinline__ wf2 load_pair<float>(const float* p) {
  return *reinterpret_cast<const wf2*>(p);
}
template <>
__device__ __forceinline__ wf2 load_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
  return wf2_make(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// column pass of one task: two adjacent columns (x even), four consecutive rows, taps in chunks of four
//   acc[o] = (sum_j tap[j] h[reflect(y0 + o - r + j)][x], same for x + 1)
// Row offsets come from the CTA's table rowoff[i] = reflect(i - kWRowPad) * W, which also covers the overhang of the
// zero-padded chunks (those rows meet zero taps).  A second, table-free body for tasks whose rows all lie inside the
// map was measured and dropped: it saves one shared load per row but the larger kernel runs slower (160 us against
// 154 us at B = 1024) -- code size matters more than instruction count here.
template <typename T>
__device__ __forceinline__ void warp_col_task(const T* __restrict__ plane, const wf2* __restrict__ g2, int nch4, int x,
                                              int y0, int r, const int* __restrict__ rowoff, wf2 (&acc)[4]) {
  wf2 w[8];
  const T* base = plane + x;
  const int* ro = rowoff + (y0 - r + kWRowPad);
  auto fetch = [&](int j) -> wf2 { return load_pair<T>(base + ro[j]); };
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = fetch(j);
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = 0ull;
  // Measured and rejected: requesting the next chunk's rows ahead of the multiply-adds (68 us against 64.5 us -- the
  // extra register moves cost more than the shared-memory latency they hide), and full unrolling per chunk count
  // through a switch (no window shifting, no loop: 67.7 us against 59.6 us on the mixed C2 inputs -- the five + three
  // unrolled bodies no longer fit the instruction cache once warps with different radii share an SM).
#pragma unroll 1
  for (int c = 0; c < nch4; ++c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) w[4 + j] = fetch(4 * c + 4 + j);
    const ulonglong2 ta = *reinterpret_cast<const ulonglong2*>(g2 + 4 * c);
    const ulonglong2 tb = *reinterpret_cast<const ulonglong2*>(g2 + 4 * c + 2);
    const wf2 t[4] = {ta.x, ta.y, tb.x, tb.y};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = wf2_fma(t[jj], w[o + jj], acc[o]);
#pragma unroll
    for (int o = 0; o < 4; ++o) w[o] = w[o + 4];
  }
}

// row pass of one task: one row pair, eight consecutive outputs from a 128-bit aligned window of the band buffer
//   acc[o] = sum_j tap[j] src[o + j]   (every element a (row 2q, row 2q + 1) pair)
__device__ __forceinline__ void warp_row_task(const wf2* __restrict__ src, const wf2* __restrict__ g2, int nch8,
                                              wf2 (&acc)[8]) {
  wf2 w[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(src)[i];
    w[2 * i] = v.x; w[2 * i + 1] = v.y;
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0ull;
#pragma unroll 1
  for (int c = 0; c < nch8; ++c) {
    wf2 t[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const ulonglong2 v = reinterpret_cast<const ulonglong2*>(src + 8 * c + 8)[i];
      w[8 + 2 * i] = v.x; w[9 + 2 * i] = v.y;
      const ulonglong2 u = reinterpret_cast<const ulonglong2*>(g2 + 8 * c)[i];
      t[2 * i] = u.x; t[2 * i + 1] = u.y;
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = wf2_fma(t[jj], w[o + jj], acc[o]);
#pragma unroll
    for (int o = 0; o < 8; ++o) w[o] = w[o + 8];
  }
}

// ---- a TEAM of G warps owns a heatmap (G = 1: a single warp, no barrier at all; G = 2, 4: named barrier of 32 G
// threads).  Every loop below strides the team; reductions are warp shuffles followed by a G-slot exchange.
template <int G>
__device__ __forceinline__ void team_sync(int team) {
  if (G == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(32 * G) : "memory");
}

constexpr int kWMaxTeam = 2;   // warps per team: 1 or 2 (4 was measured and is slower everywhere)
struct TeamExchange {      // per-team scratch for the cross-warp steps (256 bytes)
  double d[2][5][kWMaxTeam];   // partial sums of the exact evaluations, double buffered
  float f[4][kWMaxTeam];       // max, min, prefilter max
  int i[8][kWMaxTeam];         // p0, bounding box, next item / heatmap
};

template <typename T>
struct TeamExact {
  const T* plane;
  const double* w2d;   // the reference's d x d table of this channel (global memory, read-only path)
  TeamExchange* ex;
  int H, W, r, d;
  int tl, tw, team;    // lane within the team, warp within the team, team within the CTA
  int calls;           // parity selects the exchange buffer
};

// exact value of one convolved pixel: double accumulation of the d x d table, float32 result (what scipy stores,
// heatmap.py:362-364).  Team-collective, every thread returns the value.
template <typename T, int G>
__device__ __noinline__ float team_exact1(TeamExact<T>& c, int y, int x) {
  constexpr int S = 32 * G;
  const int d = c.d, n = d * d;
  const int qs = S / d, rs = S - qs * d;
  int ti = c.tl / d, tj = c.tl - ti * d;
  double a = 0.0;
#pragma unroll 1
  for (int i = c.tl; i < n; i += S) {
    const float v = plane_value<T>(c.plane, reflect1(y + ti - c.r, c.H) * c.W + reflect1(x + tj - c.r, c.W));
    a = fma(__ldg(c.w2d + i), static_cast<double>(v), a);
    tj += rs; ti += qs;
    if (tj >= d) { tj -= d; ++ti; }
  }
  a = warp_sum(a);
  if (G == 1) return static_cast<float>(a);
  const int buf = c.calls++ & 1;
  if ((c.tl & 31) == 0) c.ex->d[buf][0][c.tw] = a;
  team_sync<G>(c.team);
  double sum = c.ex->d[buf][0][0];
#pragma unroll
  for (int g = 1; g < G; ++g) sum += c.ex->d[buf][0][g];
  return static_cast<float>(sum);
}

// exact values of an interior pixel and of its left / right / upper / lower neighbours; the five windows share
// every tap.  out[0] = centre, out[1..4] = left, right, up, down.
template <typename T, int G>
__device__ __noinline__ void team_exact5(TeamExact<T>& c, int y, int x, float (&out)[5]) {
  constexpr int S = 32 * G;
  const int d = c.d, n = d * d;
  const int qs = S / d, rs = S - qs * d;
  int ti = c.tl / d, tj = c.tl - ti * d;
  double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int i = c.tl; i < n; i += S) {
    const int yy = y + ti - c.r, xx = x + tj - c.r;
    const int r0 = reflect1(yy - 1, c.H) * c.W, r1 = reflect1(yy, c.H) * c.W, r2 = reflect1(yy + 1, c.H) * c.W;
    const int c0 = reflect1(xx - 1, c.W), c1 = reflect1(xx, c.W), c2 = reflect1(xx + 1, c.W);
    const double w = __ldg(c.w2d + i);
    a[0] = fma(w, static_cast<double>(plane_value<T>(c.plane, r1 + c1)), a[0]);
    a[1] = fma(w, static_cast<double>(plane_value<T>(c.plane, r1 + c0)), a[1]);
    a[2] = fma(w, static_cast<double>(plane_value<T>(c.plane, r1 + c2)), a[2]);
    a[3] = fma(w, static_cast<double>(plane_value<T>(c.plane, r0 + c1)), a[3]);
    a[4] = fma(w, static_cast<double>(plane_value<T>(c.plane, r2 + c1)), a[4]);
    tj += rs; ti += qs;
    if (tj >= d) { tj -= d; ++ti; }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) a[q] = warp_sum(a[q]);
  if (G == 1) {
#pragma unroll
    for (int q = 0; q < 5; ++q) out[q] = static_cast<float>(a[q]);
    return;
  }
  const int buf = c.calls++ & 1;
  if ((c.tl & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 5; ++q) c.ex->d[buf][q][c.tw] = a[q];
  }
  team_sync<G>(c.team);
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    double sum = c.ex->d[buf][q][0];
#pragma unroll
    for (int g = 1; g < G; ++g) sum += c.ex->d[buf][q][g];
    out[q] = static_cast<float>(sum);
  }
}

// G warps per heatmap, TPC teams per CTA
template <typename T, int G, int TPC>
__global__ void __launch_bounds__(32 * G * TPC, 6)
decode_expected_warp_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps,
                            float* __restrict__ locs, float* __restrict__ vals, int32_t* __restrict__ argmax,
                            double* __restrict__ keypoints, WarpGeom geo, unsigned* __restrict__ work_counter,
                            const int* __restrict__ list, const unsigned* __restrict__ list_count) {
  extern __shared__ __align__(128) unsigned char wsm[];
  __shared__ __align__(8) uint64_t bars[TPC];

  constexpr int V = Elem<T>::kVec;
  constexpr int S = 32 * G;                                    // threads per team
  const int lane = threadIdx.x & 31;
  const int team = threadIdx.x / S, tl = threadIdx.x - team * S, tw = tl >> 5;
  unsigned char* slot = wsm + static_cast<size_t>(team) * geo.slot_bytes;
  const T* plane = reinterpret_cast<const T*>(slot);
  T* plane_rw = reinterpret_cast<T*>(slot);
  wf2* tmp = reinterpret_cast<wf2*>(slot + geo.tmp_off);       // band buffer: [row pair][column] of (row 2q, row 2q+1)
  wf2* taps = reinterpret_cast<wf2*>(slot + geo.taps_off);     // taps[s * kWTaps + j] = (g[j - s], g[j - s]), s = 0, 1
  int* cand = reinterpret_cast<int*>(slot + geo.cand_off);     // cand[0 .. kWCand) = pixels, cand[kWCand] = count,
  float* cand_val = reinterpret_cast<float*>(cand + kWCand + 4);   // their prefilter values
  TeamExchange* ex = reinterpret_cast<TeamExchange*>(slot + geo.exch_off);
  uint64_t* bar = &bars[team];

  const int H = p.H, W = p.W, HW = H * W, WV = W / V, NV = HW / V;
  // `list` (with its device-side length): decode only the listed heatmaps -- the ones the tensor-core kernel
  // (pp_decode_mma.cuh) handed on; the launcher guarantees B * K < 2^31
  const int N = list ? static_cast<int>(*list_count) : p.B * p.K;
  const bool tail = p.apply_tail != 0;
  const float temp = p.temperature;
  const bool pp_no_truncation = geo.full_taps != 0;
  const int step_y = fast_div(S, geo.div_WV), step_x = S - step_y * WV;
  const int first_y = fast_div(tl, geo.div_WV), first_x = tl - first_y * WV;

  if (tl == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  for (int i = tl; i < kWTmpPairs; i += S) tmp[i] = 0ull;   // stale band contents must be finite
  int* rowoff = reinterpret_cast<int*>(wsm + geo.rowoff_off);
  for (int i = threadIdx.x; i < H + 2 * kWRowPad; i += blockDim.x) rowoff[i] = reflect1(i - kWRowPad, H) * W;
  __syncthreads();

  // ---- work distribution: teams pull items one at a time from a global counter, channel-major (item j ->
  // channel order[j / B], image j % B; widest kernels first), one item ahead so that the next plane is already on
  // its way into L2 while this one is decoded.  Without a counter the heatmaps are strided statically.
  const bool dynamic = work_counter != nullptr;
  const int gteam = blockIdx.x * TPC + team, nteams = gridDim.x * TPC;
  auto item_to_hm = [&](int j) -> int {
    if (list) return list[j];
    if (!dynamic) return j;
    const int slot_k = j / p.B, b = j - slot_k * p.B;
    const int kk = tab.order ? tab.order[slot_k] : slot_k;
    return b * p.K + kk;
  };
  if (tl == 0) {
    const int j = dynamic ? static_cast<int>(min(atomicAdd(work_counter, 1u), static_cast<unsigned>(N))) : min(gteam, N);
    const int h = j < N ? item_to_hm(j) : N;
    ex->i[6][0] = j;
    ex->i[7][0] = h;
    if (j < N) {
      mbar_expect_tx(bar, geo.plane_bytes);
      tma_load_1d(slot, heatmaps + static_cast<size_t>(h) * HW, geo.plane_bytes, bar);
    }
  }
  team_sync<G>(team);
  int cur_item = ex->i[6][0], cur_hm = ex->i[7][0];
  team_sync<G>(team);

  int k_loaded = -1, r = 1, d = 3, rf = 1, df = 3, nch4 = 1;
  float tail2d = 0.0f;
  // Measured and rejected (tools/decode_phases.py shows 2.6 k cycles at the top of every heatmap for the dependent
  // radius -> taps loads of a new channel, and 4.1 k in the exact evaluation's dependent table loads): prefetching
  // the next channel's taps into registers a heatmap ahead and staging the d x d table in the idle band buffer
  // shorten those phases but not the kernel (59.7 us / 158.6 us at B = 256 / 1024 against 58.2 / 151.3) -- with twelve
  // warps per SM the waits are already covered by other warps, and the extra code costs issue slots.
  const double* w2dk = tab.kernel2d;

#ifdef PP_PHASE_TIMING
  long long mark_ = clock64();
#endif
  for (int it = 0; cur_item < N; ++it) {
    const int hm = cur_hm;
    // claim the item after this one now, look at the answer later (publish_next, after the first phases): the
    // atomic's round trip then overlaps the scans instead of stalling the warp (capture R: 7 % of the samples)
    unsigned pulled_raw = 0u;   // not looked at (not even clamped) before publish_next
    if (tl == 0) pulled_raw = dynamic ? atomicAdd(work_counter, 1u) : static_cast<unsigned>(cur_item) + nteams;
    auto publish_next = [&]() {   // thread 0: next item -> exchange slots (read at the end of the iteration) + L2 prefetch
      if (tl == 0) {
        const int pulled = static_cast<int>(min(pulled_raw, static_cast<unsigned>(N)));
        const int h = pulled < N ? item_to_hm(pulled) : N;
        ex->i[6][0] = pulled;
        ex->i[7][0] = h;
        if (pulled < N) {
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(heatmaps + static_cast<size_t>(h) * HW),
                       "r"(geo.plane_bytes)
                       : "memory");
        }
      }
    };
    const int k = hm % p.K;
    if (k != k_loaded) {   // uniform across the team; the previous heatmap ended with a team barrier
      r = tab.radius[k];
      d = 2 * r + 1;
      // The prefilter keeps the central 2 rf + 1 taps only.  For the wide kernels (radius = ceil(3 s) >= 6) the taps
      // beyond rf = 5 / 6 carry less than 5e-4 of the 2-D mass; dropping them saves a chunk of multiply-adds in
      // each pass.  What they would have added lies in [tail2d * min h, tail2d * max h] at every pixel, so it
      // changes the difference between two pixels by at most tail2d * (max h - min h): the candidate band is
      // widened by exactly that (tail2d is measured from the taps, with 1 % + 1e-6 of slack).  The lower bound L
      // and the exact evaluation always use the full d x d table.
      rf = r >= 8 ? 6 : r >= 6 ? 5 : r;
      if (pp_no_truncation) rf = r;
      df = 2 * rf + 1;
      nch4 = (df + 3) >> 2;
      w2dk = tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;
      const float* t1 = tab.taps_f32 + k * PP_OKS_TAPS + (r - rf);
      {
        // mass of the kept taps (every warp for itself, fixed order -> identical in all of them)
        double s1 = (lane < df) ? static_cast<double>(t1[lane]) : 0.0;
        s1 = warp_sum(s1);
        tail2d = rf == r ? 0.0f : static_cast<float>(fmax(1.0 - s1 * s1, 0.0) * 1.01 + 1e-6);
      }
      if (tl < kWTaps) {
        const float g0 = tl < df ? t1[tl] : 0.0f;
        const float g1 = (tl >= 1 && tl <= df) ? t1[tl - 1] : 0.0f;
        taps[tl] = wf2_make(g0, g0);
        taps[kWTaps + tl] = wf2_make(g1, g1);
      }
      k_loaded = k;
    }
    PP_WMARK(0);
    mbar_wait(bar, it & 1);
    PP_WMARK(1);

    // ---- A: head tail in place (optional), then max / min
    if (tail) {
      for (int i = tl; i < NV; i += S) {
        float f[V];
        uint4* vec = reinterpret_cast<uint4*>(plane_rw + i * V);
        unpack(*vec, f, T());
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = tail_value<T>(f[j], temp);
        *vec = pack(f, T());
      }
      team_sync<G>(team);
    }
    float xmax = -INFINITY, xmin = INFINITY;
    int ivec = 0;   // first 128-bit vector of this thread that holds its maximum: phase B looks at that one only
#pragma unroll 4
    for (int i = tl; i < NV; i += S) {
      float f[V];
      unpack(*reinterpret_cast<const uint4*>(plane + i * V), f, T());
      float m = f[0], mn = f[0];
#pragma unroll
      for (int j = 1; j < V; ++j) { m = fmaxf(m, f[j]); mn = fminf(mn, f[j]); }
      if (m > xmax) { xmax = m; ivec = i; }
      xmin = fminf(xmin, mn);
    }
    const float tmax = xmax;
    float vmax = warp_max(xmax), vmin = -warp_max(-xmin);
    if (G > 1) {
      if (lane == 0) { ex->f[0][tw] = vmax; ex->f[1][tw] = vmin; }
      team_sync<G>(team);   // also publishes the tap tables
#pragma unroll
      for (int g = 0; g < G; ++g) { vmax = fmaxf(vmax, ex->f[0][g]); vmin = fminf(vmin, ex->f[1][g]); }
    } else {
      __syncwarp();
    }

    publish_next();   // the scan above has covered the atomic's round trip
    PP_WMARK(2);

    int best = 0;
    float best_val = 0.0f, score = vmax;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    bool interior = false;

    if (vmax != vmin) {   // constant maps (e.g. all zero after the clamp): first index wins, border pixel
      // ---- B: the raw maximum p0 (lowest index) and a lower bound L of the convolved maximum
      int idx = 0x7fffffff;
      if (tmax == vmax) {
        float f[V];
        unpack(*reinterpret_cast<const uint4*>(plane + ivec * V), f, T());
#pragma unroll
        for (int j = V - 1; j >= 0; --j) idx = (f[j] == vmax) ? ivec * V + j : idx;
      }
      int imax = __reduce_min_sync(0xffffffffu, idx);
      if (G > 1) {
        if (lane == 0) ex->i[0][tw] = imax;
        team_sync<G>(team);
#pragma unroll
        for (int g = 0; g < G; ++g) imax = min(imax, ex->i[0][g]);
      }
      const int py = fast_div(imax, geo.div_W), px = imax - py * W;
      float L;
      {
        // sum over the central 5 x 5 taps + "everything else is at least vmin" (taps >= 0, sum 1); every warp of
        // the team evaluates it for itself
        const int half = min(2, r), side = 2 * half + 1;
        double sw = 0.0, swh = 0.0;
        if (lane < side * side) {
          const int ti = lane / side, tj = lane - ti * side;
          const double w = __ldg(w2dk + (r - half + ti) * d + (r - half + tj));
          const float v = plane_value<T>(plane, reflect1(py - half + ti, H) * W + reflect1(px - half + tj, W));
          sw = w; swh = w * static_cast<double>(v);
        }
        sw = warp_sum(sw); swh = warp_sum(swh);
        const float e = static_cast<float>(swh + fmax(1.0 - sw, 0.0) * static_cast<double>(vmin));
        L = e - fabsf(e) * 1e-6f - 1e-37f;
      }

      PP_WMARK(3);
      // ---- C: bounding box of S = {h >= L}; p0 is in S, so the box is never empty
      int bx0 = W, bx1 = -1, by0 = H, by1 = -1;
      {
        int y = first_y, xv = first_x;
#pragma unroll 4
        for (int i = tl; i < NV; i += S) {
          float f[V];
          unpack(*reinterpret_cast<const uint4*>(plane + i * V), f, T());
          float m = f[0];
#pragma unroll
          for (int j = 1; j < V; ++j) m = fmaxf(m, f[j]);
          if (m >= L) {
            if (xv * V < bx0 || xv * V + V - 1 > bx1) {   // flat maps: the box soon covers every column
#pragma unroll
              for (int j = 0; j < V; ++j) {
                if (f[j] >= L) { bx0 = min(bx0, xv * V + j); bx1 = max(bx1, xv * V + j); }
              }
            }
            by0 = min(by0, y); by1 = max(by1, y);
          }
          xv += step_x; y += step_y;
          if (xv >= WV) { xv -= WV; ++y; }
        }
      }
      bx0 = __reduce_min_sync(0xffffffffu, bx0); bx1 = __reduce_max_sync(0xffffffffu, bx1);
      by0 = __reduce_min_sync(0xffffffffu, by0); by1 = __reduce_max_sync(0xffffffffu, by1);
      if (G > 1) {
        if (lane == 0) { ex->i[1][tw] = bx0; ex->i[2][tw] = bx1; ex->i[3][tw] = by0; ex->i[4][tw] = by1; }
        if (tl == 0) cand[kWCand] = 0;
        team_sync<G>(team);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          bx0 = min(bx0, ex->i[1][g]); bx1 = max(bx1, ex->i[2][g]);
          by0 = min(by0, ex->i[3][g]); by1 = max(by1, ex->i[4][g]);
        }
      } else if (tl == 0) {
        cand[kWCand] = 0;
      }
      const int ox0 = max(bx0 - r, 0), oy0 = max(by0 - r, 0);
      const int ox1 = min(bx1 + r, W - 1), oy1 = min(by1 + r, H - 1);
      const int OW = ox1 - ox0 + 1, OH = oy1 - oy0 + 1;

      PP_WMARK(4);
      // ---- D/E: separable float32 prefilter over the region, in bands of row pairs.
      // Band buffer: tmp[q][c] = column-pass values of rows (ya + 2q, ya + 2q + 1) at source column x = c - cbase
      // (reflected columns included); cbase is even so that a column task stores its 2 x 2 values with one 128-bit
      // store, which shifts the row-pass taps by s = 0 / 1 column (second tap table).  Output block t of a row pair
      // reads tmp[q][8 t .. 8 t + 8 nch8 + 7].
      const int sh = (rf - ox0) & 1;
      const int cbase = rf - ox0 + sh;                      // tmp column of source column x: x + cbase
      const int nch8 = (df + sh + 7) >> 3;
      const wf2* rtaps = taps + sh * kWTaps;
      const int nxb = (OW + 7) >> 3;
      int TS = 8 * nxb + 8 * nch8;                          // row-pair stride in pairs: 2 * odd -> conflict-free
      if (((TS >> 1) & 1) == 0) TS += 2;
      const int BQ = min((OH + 3) >> 2 << 1, (kWTmpPairs / TS) & ~1);   // row pairs per band (even)
      const int BR = 2 * BQ;
      const int xs0 = max(ox0 - rf, 0) & ~1;                // first source pair (even column)
      const int xs1 = min(ox1 + rf, W - 1);
      const int npair = ((xs1 - xs0) >> 1) + 1;
      const unsigned mpair = div_magic(npair);
      const int cmax = OW + 2 * rf - 1 + sh;                // last tmp column that meets a non-zero tap

      const float amax = fmaxf(fabsf(vmax), fabsf(vmin));
      const float gamma = static_cast<float>(2 * df + 8) * 1.1920929e-7f;   // (2 df + 8) * 2^-23
      const float band = (2.0f * gamma + 4.0f * 5.9604645e-8f) * amax + tail2d * (vmax - vmin);

      // Candidates.  Only a pixel whose prefilter value lies within `band` of the prefilter maximum can be the exact
      // maximum, and the final maximum is at least the warp's running one (gm, refreshed with one shuffle reduction
      // per round of tasks): a task reports a pixel only when it comes within `band` of gm, which happens a handful
      // of times per heatmap (new running maxima and near ties).  Reported (pixel, value) pairs go to a short list
      // that is filtered with the final threshold afterwards.  Pass 1 (only when the list overflowed: plateaus,
      // heavily quantised maps) repeats the sweep with the final threshold.
      float gm = -INFINITY, thr = 0.0f;
      int count = 0;

#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const bool collect = pass == 1;
#pragma unroll 1
        for (int ya = oy0; ya <= oy1; ya += BR) {
          const int rows = min(BR, oy1 + 1 - ya), nyb = (rows + 3) >> 2;
          team_sync<G>(team);   // the previous band's row pass is done with tmp
          for (int t = tl; t < npair * nyb; t += S) {
            const int yb = fast_div(t, mpair), x = xs0 + 2 * (t - yb * npair);
            const int y0 = ya + 4 * yb;
            wf2 acc[4];
            warp_col_task<T>(plane, taps, nch4, x, y0, rf, rowoff, acc);
            float a0[4], a1[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) wf2_split(acc[o], a0[o], a1[o]);
            const int c = x + cbase;                       // even, >= 0
            const int cl = -1 - x + cbase;                 // tmp column of the left mirror of x (x + 1: cl - 1)
            const int cr = 2 * W - 1 - x + cbase;          // tmp column of the right mirror of x (x + 1: cr - 1)
            wf2* row = tmp + 2 * yb * TS;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              *reinterpret_cast<float4*>(row + c) = make_float4(a0[2 * q], a0[2 * q + 1], a1[2 * q], a1[2 * q + 1]);
              if (cl >= 0 || cr - 1 <= cmax) {   // columns within the radius of a map edge also fill their mirror
                if (cl >= 0) row[cl] = wf2_make(a0[2 * q], a0[2 * q + 1]);
                if (cl >= 1) row[cl - 1] = wf2_make(a1[2 * q], a1[2 * q + 1]);
                if (cr <= cmax) row[cr] = wf2_make(a0[2 * q], a0[2 * q + 1]);
                if (cr - 1 <= cmax) row[cr - 1] = wf2_make(a1[2 * q], a1[2 * q + 1]);
              }
              row += TS;
            }
          }
          team_sync<G>(team);
          PP_WMARK(5);   // column passes (+ the row passes of all bands but the last)
          const int nq = (rows + 1) >> 1, ntask = nq * nxb;
          const unsigned mq = div_magic(nq);
          for (int t0 = 0; t0 < ntask; t0 += S) {   // whole warps: every lane joins the reduction of the round
            const int t = t0 + tl;
            float v[16];   // v[o] = row ya + 2q, v[8 + o] = row ya + 2q + 1
            float tm = -INFINITY;
            int pix = 0;
            if (t < ntask) {
              const int xb = fast_div(t, mq), q = t - xb * nq;
              wf2 acc[8];
              warp_row_task(tmp + q * TS + 8 * xb, rtaps, nch8, acc);
#pragma unroll
              for (int o = 0; o < 8; ++o) wf2_split(acc[o], v[o], v[8 + o]);
              const int nvalid = min(8, OW - 8 * xb);
              if (nvalid < 8 || 2 * q + 1 >= rows) {   // partial block / odd last row: mask what lies outside
                const bool row1 = 2 * q + 1 < rows;
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                  if (o >= nvalid) v[o] = -INFINITY;
                  if (o >= nvalid || !row1) v[8 + o] = -INFINITY;
                }
              }
              tm = fmaxf(v[0], v[8]);
#pragma unroll
              for (int o = 1; o < 8; ++o) tm = fmaxf(tm, fmaxf(v[o], v[8 + o]));
              pix = (ya + 2 * q) * W + ox0 + 8 * xb;
            }
            float lo = thr;
            if (!collect) {
              gm = fmaxf(gm, warp_max(tm));
              lo = gm - band;
            }
            if (tm >= lo) {   // rare
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                if (v[e] >= lo) {
                  const int s = atomicAdd(&cand[kWCand], 1);
                  if (s < kWCand) { cand[s] = pix + (e & 7) + (e >> 3) * W; cand_val[s] = v[e]; }
                }
              }
            }
          }
        }
        PP_WMARK(6);
        if (collect) break;
        // prefilter maximum -> threshold; keep the listed pixels that are still inside the band
        float pm = gm;
        if (G > 1) {
          if (lane == 0) ex->f[2][tw] = pm;
          team_sync<G>(team);
#pragma unroll
          for (int g = 0; g < G; ++g) pm = fmaxf(pm, ex->f[2][g]);
        } else {
          __syncwarp();
        }
        thr = pm - band;
        const int raw = cand[kWCand];
        if (raw <= kWCand) {
          // every thread takes the entries tl, tl + S, ... into registers, then the survivors are written back
          int keep_id[kWCand / 32];
          bool keep[kWCand / 32];
#pragma unroll
          for (int u = 0; u < kWCand / 32; ++u) {
            const int e = tl + u * S;
            keep[u] = e < raw && cand_val[min(e, kWCand - 1)] >= thr;
            keep_id[u] = cand[min(e, kWCand - 1)];
          }
          team_sync<G>(team);
          if (tl == 0) cand[kWCand] = 0;
          team_sync<G>(team);
#pragma unroll
          for (int u = 0; u < kWCand / 32; ++u)
            if (keep[u]) cand[atomicAdd(&cand[kWCand], 1)] = keep_id[u];
          break;
        }
        team_sync<G>(team);   // everybody has read the count
        if (tl == 0) cand[kWCand] = 0;   // published by the first barrier of the band loop
      }
      team_sync<G>(team);
      count = cand[kWCand];

      PP_WMARK(7);
      // ---- G: exact values of the candidates and of the winner's four neighbours
      TeamExact<T> te;
      te.plane = plane; te.w2d = w2dk; te.ex = ex; te.H = H; te.W = W; te.r = r; te.d = d;
      te.tl = tl; te.tw = tw; te.team = team; te.calls = 0;
      if (count == 1) {
        best = cand[0];
      } else if (count <= kWCand) {
        best_val = -INFINITY; best = 0x7fffffff;
        for (int q = 0; q < count; ++q) {
          const int ci = cand[q], cy = fast_div(ci, geo.div_W);
          argmax_combine(best_val, best, team_exact1<T, G>(te, cy, ci - cy * W), ci);
        }
      } else {
        // more near-maximal pixels than the list holds (plateaus): every pixel of the region, exactly
        best_val = -INFINITY; best = 0x7fffffff;
        for (int y = oy0; y <= oy1; ++y)
          for (int x = ox0; x <= ox1; ++x) argmax_combine(best_val, best, team_exact1<T, G>(te, y, x), y * W + x);
      }
      const int by = fast_div(best, geo.div_W), bx = best - by * W;
      interior = bx > 0 && bx < W - 1 && by > 0 && by < H - 1;
      if (interior) {
        float ev[5];
        team_exact5<T, G>(te, by, bx, ev);
        best_val = ev[0];
        nb[0] = ev[1]; nb[1] = ev[2]; nb[2] = ev[3]; nb[3] = ev[4];
      }
      score = plane_value<T>(plane, best);
      PP_WMARK(8);
    }

    // ---- H: outputs.  The x and y halves of the sub-pixel fit (heatmap.py:136-165, float32, the reference's
    // operation order) are independent: thread 0 does x, thread 1 does y.
    if (tl < 2) {
      const int by = fast_div(best, geo.div_W), bx = best - by * W;
      const bool is_y = tl == 1;
      float f = static_cast<float>(is_y ? by : bx);
      if (interior) {
        const float lo = is_y ? nb[2] : nb[0], hi = is_y ? nb[3] : nb[1], c = best_val;   // left/up, right/down
        const float g = __fdiv_rn(__fsub_rn(hi, lo), 2.0f);
        float h = __fsub_rn(__fadd_rn(hi, lo), __fmul_rn(2.0f, c));
        if (h == 0.0f) h = 1e-6f;
        f = __fadd_rn(f, __fdiv_rn(-g, h));
      }
      locs[static_cast<size_t>(hm) * 2 + (is_y ? 1 : 0)] = f;
      if (keypoints)   // float32 / int -> float64, then * input_size (codec.py:237)
        keypoints[static_cast<size_t>(hm) * 2 + (is_y ? 1 : 0)] =
            static_cast<double>(f) / static_cast<double>(is_y ? H - 1 : W - 1) * (is_y ? p.input_h : p.input_w);
      if (is_y) {
        vals[hm] = score;
        if (argmax) argmax[hm] = best;
      }
    }

    PP_WMARK(10);
    // ---- next heatmap: everybody is done with the plane, thread 0 starts the copy (an L2 hit by now)
    team_sync<G>(team);
    cur_item = ex->i[6][0];
    cur_hm = ex->i[7][0];
    if (tl == 0 && cur_item < N) {
      fence_proxy_async();
      mbar_expect_tx(bar, geo.plane_bytes);
      tma_load_1d(slot, heatmaps + static_cast<size_t>(cur_hm) * HW, geo.plane_bytes, bar);
    }
    team_sync<G>(team);   // ex->i[6..7] are rewritten at the top of the next iteration
    PP_WMARK(11);
  }
}

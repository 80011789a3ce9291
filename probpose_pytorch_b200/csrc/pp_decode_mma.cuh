// Expected-OKS decoder with a tensor-core prefilter: one warp per heatmap, the separable OKS convolution of the WHOLE
// map evaluated as two small banded-Toeplitz GEMMs on mma.sync.m16n8k16 (f16 x f16 -> f32, SASS HMMA.16816.F32).
// Included into pp_decode.cu after pp_decode_warp.cuh, same namespace.
//
// Why.  The pruned float32 prefilter of pp_decode_warp.cuh costs 4.4 k warp instructions per 64x48 heatmap on blob-shaped
// maps and about twice that on flat / noise-only maps (bounding box of {h >= L}, region set-up, banded column and row
// passes, candidate tracking), runs at 41 % issue utilisation with twelve warps per SM and has a long tail at small
// batches (profiles/r01s_summary.md).  The convolution with reflect boundaries is a product with two constant banded
// matrices per channel,
//        R = Ty . H . Tx^T,      Tx[x'][x] = sum of the 1-D taps j with reflect(x' + j - r) == x   (same for Ty),
// i.e. 0.3 MFLOP per 64x48 map: 116 HMMA instructions of one warp instead of ~1800 two-wide FFMAs plus their shared
// memory traffic -- and the same 116 for every map, so there is no bounding box, no region logic and no tail.  The
// tensor cores only PROPOSE candidates; every decision is made on exact values:
//   * h is shifted and scaled into [0, 1) (h~ = (h - min h) * 2^k: the convolution commutes with both, the taps sum to
//     one in every folded row) and rounded to float16, as are the folded taps and the intermediate; accumulation is
//     float32.  With u = 2^-11: |Z(p) - R(p) 2^k - const| <= E = (4 u + slack) for every pixel, rigorous, see kMmaErr;
//   * a pixel can only be the exact argmax if Z(p) >= max Z - 2 E - (float32 rounding of R): those pixels -- one to
//     three on blob-shaped maps, a handful on noise -- are re-evaluated in float64 with the reference's own d x d table
//     (NumPy tie-break), as are the winner's four neighbours for the float32 sub-pixel fit in the reference's operation
//     order.  That part is pp_decode_warp.cuh's;
//   * maps that defeat the proposal step (more than kWCand near-maximal pixels: exact plateaus wider than the kernel,
//     heavily quantised maps; or no float32 dynamic range at all) are appended to a list in global memory and decoded
//     by decode_expected_warp_kernel in a second, usually empty, launch.
// Data flow of one heatmap (64x48 float32): TMA bulk copy of the plane into the warp's 12 KB slot -> one 128-bit scan
// (min / max) -> 48 64-bit loads, scale + convert into the 48 B-fragment registers of H^T (the warp now holds the whole
// map in float16) -> per block of 16 output columns: GEMM 1 (Y^T = Tx . H^T, <= 3 x 8 HMMA, A-fragments of Tx from a
// per-channel table in L1 / L2), accumulators repacked in registers into the A-fragments of GEMM 2 (the C layout of two
// adjacent n-blocks IS the A layout of one k-block), GEMM 2 (Z^T = Y^T . Ty^T, B-fragments of Ty from the table),
// per-lane block maximum; after all blocks the warp-wide maximum fixes the threshold and the one or two blocks that
// reach it are recomputed to list their candidates.  No shared-memory writes besides the candidate list.
#pragma once

#include <cuda_fp16.h>

// Rigorous bound of |Z - (exact scaled value) - constant|, in units of the scaled range (h~ in [0, 1)):
//   h~ -> f16: u h~ + eta;  folded taps -> f16: u t + eta;  GEMM 1 in f32: gamma;  Y -> f16: u Y + eta;  GEMM 2: gamma
//   => (4 u + 2 gamma) + (W + H + 4) eta + |g (x) g - k2d| (1.3e-7),  u = 2^-11 = 4.88e-4, eta = 2^-25 (f16 subnormal
//   half-spacing), gamma <= 64 * 2^-22 (truncating float32 accumulation of at most 96 products).  4 u = 1.953e-3.
constexpr float kMmaErr = 2.2e-3f;

template <int HH, int WW>
struct MmaShape {
  static constexpr int H = HH, W = WW;
  static constexpr int MB = (W + 15) / 16;   // blocks of 16 output columns x' (M of both GEMMs)
  static constexpr int KB = MB;              // blocks of 16 input columns x (K of GEMM 1)
  static constexpr int NB = H / 8;           // blocks of 8 rows (N of both GEMMs)
  static constexpr int NBP = NB / 2;
  static constexpr int KB2 = H / 16;         // blocks of 16 input rows y (K of GEMM 2)
  static constexpr int kT1 = MB * KB * 32;   // uint4 (8 halves) per channel: A-fragments of Tx, [mb][kb][lane]
  static constexpr int kT2 = KB2 * NBP * 32; // uint4 per channel: B-fragments of Ty for n-block pairs, [kb2][nbp][lane]
  static_assert(H % 16 == 0 && W % 8 == 0, "shape not tiled by the fragments");
};

// ---- operand tables (built once per codec and shape by pp_oks_mma_table_build) ----
__device__ __forceinline__ float folded_tap(const float* __restrict__ g, int r, int n, int o, int i, bool zero_pad) {
  // entry [o][i] of the Toeplitz matrix of taps g (radius r) on an axis of length n: taps that reach beyond the axis are
  // folded back by the 'reflect' boundary (scipy.ndimage, heatmap.py:361-362) or dropped (zero padding, codec.py:303-310)
  if (zero_pad) return (i - o >= -r && i - o <= r) ? g[i - o + r] : 0.0f;
  float s = 0.0f;
  for (int j = 0; j <= 2 * r; ++j)
    if (reflect1(o + j - r, n) == i) s += g[j];
  return s;
}

template <int H, int W>
__global__ void __launch_bounds__(256)
build_mma_tables_kernel(const float* __restrict__ taps, const int* __restrict__ radius, int U, __half* __restrict__ out,
                        int taps_stride, int fixed_radius, bool zero_pad) {
  using S = MmaShape<H, W>;
  constexpr int per = (S::kT1 + S::kT2) * 8;   // halves per channel
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < static_cast<long long>(U) * per;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx / per);
    int e = static_cast<int>(idx - static_cast<long long>(u) * per);
    const float* g = taps + u * taps_stride;
    const int r = radius ? radius[u] : fixed_radius;
    float v = 0.0f;
    if (e < S::kT1 * 8) {
      // A fragment of Tx, tile (mb, kb): reg q -> rows g / g+8 (q & 1), k-columns 2t.. / 2t+8.. (q >> 1)
      const int hh = e & 1, q = (e >> 1) & 3, lane = (e >> 3) & 31, tile = e >> 8;
      const int mb = tile / S::KB, kb = tile - mb * S::KB;
      const int gg = lane >> 2, t = lane & 3;
      const int xo = 16 * mb + gg + 8 * (q & 1), xi = 16 * kb + 2 * t + hh + 8 * (q >> 1);
      if (xo < W && xi < W) v = folded_tap(g, r, W, xo, xi, zero_pad);
    } else {
      // B fragments of Ty^T for the n-blocks (2 nbp, 2 nbp + 1), tile (kb2, nbp): reg q -> n-block (q >> 1),
      // k-rows 2t.. / 2t+8.. (q & 1); element [k = y][n = y'] = Ty[y'][y]
      e -= S::kT1 * 8;
      const int hh = e & 1, q = (e >> 1) & 3, lane = (e >> 3) & 31, tile = e >> 8;
      const int kb2 = tile / S::NBP, nbp = tile - kb2 * S::NBP;
      const int gg = lane >> 2, t = lane & 3;
      const int yi = 16 * kb2 + 2 * t + hh + 8 * (q & 1), yo = 8 * (2 * nbp + (q >> 1)) + gg;
      v = folded_tap(g, r, H, yo, yi, zero_pad);
    }
    out[idx] = __float2half_rn(v);
  }
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// (p[0], p[1]) of the staged plane as floats; p is 2-element aligned
template <typename T>
__device__ __forceinline__ float2 plane_pair(const T* p);
template <>
__device__ __forceinline__ float2 plane_pair<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
template <>
__device__ __forceinline__ float2 plane_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

struct MmaGeom {
  unsigned plane_bytes;   // H * W * sizeof(T)
  unsigned cand_off;      // byte offset of the candidate list inside a warp's slot
  unsigned slot_bytes;    // per-warp slot (multiple of 128)
  unsigned div_W;
  unsigned static_split;  // 1: items are split statically over the warps (PP_DECODE_STATIC=1) instead of pulled from a counter
};

// scratch layout (unsigned words): [0] work counter of this kernel, [1] work counter of the hand-over launch,
// [2] number of handed-over heatmaps, [3] unused, [4 ..] their indices
constexpr int kMmaScratchHead = 4;

// ---- exact values (float64 accumulation of the reference's d x d table, float32 result: what scipy stores,
// heatmap.py:362-364), one warp, G = 1 versions of team_exact1 / team_exact5 of pp_decode_warp.cuh.  A lane owns the
// taps lane, lane + 32, ...; they are processed four at a time with the four table entries requested up front, so that
// their L1 / L2 latency is paid once per group instead of once per tap (capture r02c: the dependent table load of every
// loop iteration was the largest stall of the kernel after the work-queue atomic).  One body for every radius and for
// windows inside / across the map border: the fully unrolled per-radius variants of a first version ran into the
// instruction cache instead (capture r02h: 29 % of the stall samples were instruction fetches).
// Exact float64 values R(p) of up to NC candidate pixels at once (cand[0 .. n_c), flat indices; NC = 2 or 4): every tap
// weight is loaded once and used NC times, and the NC accumulation chains overlap.  One candidate per call cost ~1.4 us
// each whatever the radius -- the call's fixed latency (cold table load, fp64 shuffle reduction), not its arithmetic --
// which was the long tail of a small batch (tools/decode_timeline.py).  Slots beyond n_c repeat candidate 0 (NC = 2 for
// the common two-candidate map keeps that waste to one slot).  The taps are summed in the same lane / order whatever
// the grouping (per-lane partial sums, then warp_sum), so the values do not depend on it.
template <typename T, int NC>
__device__ __noinline__ void mma_exact_n(const T* __restrict__ plane, const double* __restrict__ w2d, int H, int W, int r,
                                         const int* __restrict__ cand, int n_c, unsigned div_W, int lane, float (&out)[NC]) {
  const int d = 2 * r + 1, n = d * d;
  const int qs = 32 / d, rs = 32 - qs * d;
  int cy[NC], cx[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int ci = cand[c < n_c ? c : 0];
    cy[c] = fast_div(ci, div_W);
    cx[c] = ci - cy[c] * W - r;
    cy[c] -= r;
  }
  int ti = lane / d, tj = lane - ti * d;
  double a[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) a[c] = 0.0;
  constexpr int U = NC == 2 ? 4 : 2;   // tap weights in flight per lane
#pragma unroll 1
  for (int i0 = lane; i0 < n; i0 += 32 * U) {
    double w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = (i0 + 32 * u < n) ? __ldg(w2d + i0 + 32 * u) : 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + 32 * u < n) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float v = plane_value<T>(plane, reflect1(cy[c] + ti, H) * W + reflect1(cx[c] + tj, W));
          a[c] = fma(w[u], static_cast<double>(v), a[c]);
        }
      }
      tj += rs; ti += qs;
      if (tj >= d) { tj -= d; ++ti; }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) out[c] = static_cast<float>(warp_sum(a[c]));
}

// an interior pixel and its left / right / upper / lower neighbours (out[0..4]); the five windows share every tap.
// inside: all five windows lie inside the map (no reflection) -- the common case, plain offsets.
template <typename T>
__device__ __noinline__ void mma_exact5(const T* __restrict__ plane, const double* __restrict__ w2d, int H, int W, int r,
                                        int y, int x, int lane, float (&out)[5]) {
  const int d = 2 * r + 1, n = d * d;
  const int qs = 32 / d, rs = 32 - qs * d;
  const bool inside = x - r >= 1 && x + r < W - 1 && y - r >= 1 && y + r < H - 1;
  int ti = lane / d, tj = lane - ti * d;
  double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int i0 = lane; i0 < n; i0 += 128) {
    double w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = (i0 + 32 * u < n) ? __ldg(w2d + i0 + 32 * u) : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + 32 * u < n) {
        const int yy = y + ti - r, xx = x + tj - r;
        int i_c = yy * W + xx, i_l = i_c - 1, i_r = i_c + 1, i_u = i_c - W, i_d = i_c + W;
        if (!inside) {
          const int r0 = reflect1(yy - 1, H) * W, r1 = reflect1(yy, H) * W, r2 = reflect1(yy + 1, H) * W;
          const int c0 = reflect1(xx - 1, W), c1 = reflect1(xx, W), c2 = reflect1(xx + 1, W);
          i_c = r1 + c1; i_l = r1 + c0; i_r = r1 + c2; i_u = r0 + c1; i_d = r2 + c1;
        }
        a[0] = fma(w[u], static_cast<double>(plane_value<T>(plane, i_c)), a[0]);
        a[1] = fma(w[u], static_cast<double>(plane_value<T>(plane, i_l)), a[1]);
        a[2] = fma(w[u], static_cast<double>(plane_value<T>(plane, i_r)), a[2]);
        a[3] = fma(w[u], static_cast<double>(plane_value<T>(plane, i_u)), a[3]);
        a[4] = fma(w[u], static_cast<double>(plane_value<T>(plane, i_d)), a[4]);
      }
      tj += rs; ti += qs;
      if (tj >= d) { tj -= d; ++ti; }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) out[q] = static_cast<float>(warp_sum(a[q]));
}

// ---- argmax + DARK-UDP decoder on the same skeleton (kDark): zero-padded ksize x ksize blur of single pixels in the
// operation order of decode_dark_fast_kernel's tile path, which is cv2's separable filter (rows first): float32 fmaf
// chains over the taps in ascending order.  Two pixels per call: half-warp h evaluates pix[h], its lane dy owns row dy
// of the window (ksize <= 16); the column chain runs over shuffled row values.  Returns (blur(pixA), blur(pixB)).
template <typename T>
__device__ __noinline__ float2 dark_blur2(const T* __restrict__ plane, const float* __restrict__ taps, int ksize, int H, int W,
                                          int pixA, int pixB, int lane) {
  const int r = ksize >> 1, dy = lane & 15, pix = (lane & 16) ? pixB : pixA;
  const int y = pix / W, x = pix - y * W, yy = y + dy - r;
  float row = 0.0f;
  if (dy < ksize && yy >= 0 && yy < H) {
    const T* line = plane + yy * W;
#pragma unroll 1
    for (int j = 0; j < ksize; ++j) {
      const int xx = x + j - r;
      const float v = (xx >= 0 && xx < W) ? Elem<T>::to_f32(line[xx]) : 0.0f;
      row = fmaf(taps[j], v, row);
    }
  }
  float acc = 0.0f;
#pragma unroll 1
  for (int d = 0; d < ksize; ++d) acc = fmaf(taps[d], __shfl_sync(0xffffffffu, row, (lane & 16) + d), acc);
  return make_float2(__shfl_sync(0xffffffffu, acc, 0), __shfl_sync(0xffffffffu, acc, 16));
}

struct MmaDarkArgs {   // what the kDark instance needs besides the expected-OKS kernel's arguments
  const float* blur_taps;   // (ksize) float32
  int ksize;
  float* peaks;             // out (N, 2) or null
  unsigned long long* dbg_times;   // kDebug only: 8 words per heatmap (tools/decode_timeline.py), or null
};

constexpr int kMmaMaxK = 256;    // channels whose work-queue order / radius / table index are staged in shared memory

// kDark: the argmax + DARK-UDP decoder (ArgMaxProbMap.decode, codec.py:515-543) on the same skeleton: the proposal step
// runs on the zero-padded 11 x 11 Gaussian blur instead of the OKS kernel (one operand table for every channel), its
// candidates give the blurred maximum (exact float32 re-evaluation), seven more exact values around the raw peak feed
// the refinement of pp_dark_fast.cuh.  locs = refined coordinates, vals = raw maxima (scores).
template <typename T, int H, int W, int WPC, int MINB, bool kDebug, bool kDark>
__global__ void __launch_bounds__(32 * WPC, MINB)
decode_expected_mma_kernel(pp_decode_params p, pp_oks_table tab, const T* __restrict__ heatmaps,
                           float* __restrict__ locs, float* __restrict__ vals, int32_t* __restrict__ argmax,
                           double* __restrict__ keypoints, MmaGeom geo, unsigned* __restrict__ scratch,
                           float* __restrict__ dbg_prefilter, MmaDarkArgs dk) {
  using S = MmaShape<H, W>;
  extern __shared__ __align__(128) unsigned char msm[];
  __shared__ __align__(8) uint64_t bars[WPC];
  __shared__ int ch_hm[kMmaMaxK];                 // channel of queue position q (the table's `order`)
  __shared__ unsigned short ch_tab[kMmaMaxK];     // operand table of channel k
  __shared__ unsigned char ch_rad[kMmaMaxK];      // radius of channel k
  __shared__ float blur_taps[kDark ? 16 : 1];     // kDark: the blur's 1-D taps

  constexpr int V = Elem<T>::kVec;
  constexpr int HW = H * W, NV = HW / V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gg = lane >> 2, t4 = lane & 3;
  unsigned char* slot = msm + static_cast<size_t>(warp) * geo.slot_bytes;
  const T* plane = reinterpret_cast<const T*>(slot);
  T* plane_rw = reinterpret_cast<T*>(slot);
  int* cand = reinterpret_cast<int*>(slot + geo.cand_off);          // cand[0 .. kWCand) pixels, cand[kWCand] count
  uint64_t* bar = &bars[warp];

  const int N = p.B * p.K;   // the launcher guarantees N < 2^31
  const bool tail = p.apply_tail != 0;
  const float temp = p.temperature;
  unsigned* retry_count = scratch + 2;
  int* retry_list = reinterpret_cast<int*>(scratch + kMmaScratchHead);

  // per-channel constants: staged once per CTA, so that nothing on a heatmap's path waits for a dependent global load
  if (kDark) {
    if (threadIdx.x < 16) blur_taps[kDark ? threadIdx.x : 0] = static_cast<int>(threadIdx.x) < dk.ksize ? dk.blur_taps[threadIdx.x] : 0.0f;
  } else {
    for (int i = threadIdx.x; i < p.K; i += blockDim.x) {
      ch_hm[i] = tab.order ? tab.order[i] : i;
      ch_tab[i] = static_cast<unsigned short>(tab.mma_index[i]);
      ch_rad[i] = static_cast<unsigned char>(tab.radius[i]);
    }
  }
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  // Work distribution.  Items are channel-major (item j -> channel order[j / B], image j % B; consecutive warps take
  // consecutive items), so the warps of an SM mostly share one channel's operand tables in L1.  A warp's first two
  // items are fixed (warp, warp + warps: no atomic on the way in -- 2 x 2368 same-address atomics at kernel start cost
  // ~15 us, measured); from the third on it pulls them from a global counter, TWO heatmaps ahead: the item after the
  // next one is claimed at the top of an iteration and looked at at its end, so the atomic's round trip never stalls
  // the warp (consumed half an iteration later it cost 17 % of the samples, capture r02c), while the tail of a large
  // batch still balances (B = 1024: 146 us against 173 us with a purely static split).  geo.static_split: static only.
  const int gwarp = blockIdx.x * WPC + warp, nwarps = gridDim.x * WPC;
  const bool dynamic = geo.static_split == 0 && N > 2 * nwarps;
  unsigned* work_counter = scratch;
  auto item_to_hm = [&](int j) -> int {
    if (kDark) return j;   // one operand table for every channel: natural order
    const int slot_k = j / p.B, b = j - slot_k * p.B;
    return b * p.K + ch_hm[slot_k];
  };
  int cur_item = min(gwarp, N), next_item = min(gwarp + nwarps, N);
  if (lane == 0 && cur_item < N) {
    mbar_expect_tx(bar, geo.plane_bytes);
    tma_load_1d(slot, heatmaps + static_cast<size_t>(item_to_hm(cur_item)) * HW, geo.plane_bytes, bar);
  }

  for (int it = 0; cur_item < N; ++it) {
    const int hm = item_to_hm(cur_item);
    unsigned pulled_raw = 0u;
    if (dynamic && lane == 0) pulled_raw = atomicAdd(work_counter, 1u);   // the item after the next one
    const int next_hm = next_item < N ? item_to_hm(next_item) : 0;

    const int k = kDark ? 0 : hm % p.K;
    const int r = kDark ? dk.ksize >> 1 : ch_rad[k];
    const uint4* t1 = reinterpret_cast<const uint4*>(tab.mma_tables) + (kDark ? 0 : static_cast<size_t>(ch_tab[k]) * (S::kT1 + S::kT2));
    const uint4* t2 = t1 + S::kT1;
    const double* w2dk = kDark ? nullptr : tab.kernel2d + static_cast<size_t>(k) * PP_OKS_TAPS * PP_OKS_TAPS;

    // kDebug: per-heatmap timeline -- [0] global ns at the start, [1..6] SM clock at start / plane arrived / scan done /
    // sweep done / re-evaluation done / outputs written, [7] smid << 32 | candidates << 8 | iteration
    unsigned long long* tl = (kDebug && dk.dbg_times) ? dk.dbg_times + static_cast<size_t>(hm) * 8 : nullptr;
    int tl_count = 0;
    if (kDebug && tl && lane == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      tl[0] = ns;
      tl[1] = clock64();
    }

    mbar_wait(bar, it & 1);
    if (kDebug && tl && lane == 0) tl[2] = clock64();
    // the next plane is on its way into L2 while this one is decoded -- asked for AFTER this one has arrived: at kernel
    // start every warp's first request would otherwise compete with its second (timeline: first plane after 1.5 us
    // instead of 4 us)
    if (lane == 0 && next_item < N)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(heatmaps + static_cast<size_t>(next_hm) * HW),
                   "r"(geo.plane_bytes)
                   : "memory");

    // ---- A: head tail in place (optional), then min / max
    if (tail) {
      for (int i = lane; i < NV; i += 32) {
        float f[V];
        uint4* vec = reinterpret_cast<uint4*>(plane_rw + i * V);
        unpack(*vec, f, T());
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = tail_value<T>(f[j], temp);
        *vec = pack(f, T());
      }
      __syncwarp();
    }
    // ---- A': min / max: a 128-bit scan of the plane (kDark: and the first vector that holds the lane's maximum); the
    // B-fragment pairs are loaded (64-bit) and converted after the scale is known.  Measured and rejected: loading the
    // pairs once and taking min / max from the same registers -- the 96 live registers spill (68 bytes per thread,
    // reloaded through L2): 45.7 us against 44.1 us at C2, 286 us against 259 us at C5.
    float vmax = -INFINITY, vmin = INFINITY;
    int ivec = 0;
#pragma unroll 4
    for (int i = lane; i < NV; i += 32) {
      float f[V];
      unpack(*reinterpret_cast<const uint4*>(plane + i * V), f, T());
      float m = fmaxf(f[0], f[1]);
      vmin = fminf(vmin, fminf(f[0], f[1]));
#pragma unroll
      for (int j = 2; j < V; j += 2) {
        m = fmaxf(m, fmaxf(f[j], f[j + 1]));
        vmin = fminf(vmin, fminf(f[j], f[j + 1]));
      }
      if (kDark) {
        if (m > vmax) { vmax = m; ivec = i; }
      } else {
        vmax = fmaxf(vmax, m);
      }
    }
    const float lane_max = vmax;
    vmax = warp_max(vmax);
    vmin = -warp_max(-vmin);
    int p0 = 0;   // kDark: the raw maximum, lowest flat index (NumPy argmax, heatmap.py:35-41)
    if (kDark) {
      int idx = 0x7fffffff;
      if (lane_max == vmax) {
        float f[V];
        unpack(*reinterpret_cast<const uint4*>(plane + ivec * V), f, T());
#pragma unroll
        for (int j = V - 1; j >= 0; --j) idx = (f[j] == vmax) ? ivec * V + j : idx;
      }
      p0 = __reduce_min_sync(0xffffffffu, idx);
    }

    if (kDebug && tl && lane == 0) tl[3] = clock64();
    int best = 0;
    float best_val = 0.0f, score = vmax;
    float nb[4] = {0.f, 0.f, 0.f, 0.f};
    bool interior = false, handed_over = false;

    float dark_x = -1.0f, dark_y = -1.0f;   // kDark: refined coordinates; empty channels (max <= 0) keep the -1 sentinel
    if (kDark ? vmax > 0.0f : vmax != vmin) {   // constant maps (e.g. all zero after the clamp): first index wins, border pixel
      // ---- B: scale.  range 2^k in [0.5, 1); maps without float32 dynamic range go to the general kernel.  kDark: the
      // zero-padded blur does not commute with an offset (the tap mass inside the map varies along the border), so the
      // map is only scaled, by its largest magnitude
      const float amax = fmaxf(fabsf(vmax), fabsf(vmin));
      const float range = kDark ? amax : vmax - vmin;
      const int ebits = static_cast<int>((__float_as_uint(range) >> 23) & 0xffu);
      const float sc = __uint_as_float(static_cast<unsigned>(253 - ebits) << 23);
      // candidate band in scaled units: twice the proposal error + the float32 rounding of the two exact values
      const float band = 2.0f * kMmaErr + 4.0f * 5.9604645e-8f * amax * sc;
      handed_over = ebits < 30 || ebits > 250 || !(band < 0.25f);
      if (!handed_over) {
        const float off = kDark ? 0.0f : -vmin * sc;
        const wf2 sc2 = wf2_make(sc, sc), off2 = wf2_make(off, off);
        // ---- C: the whole map as float16 B-fragments of H^T: hf[kb][nb] = rows 8 nb + g, columns 16 kb + 2 t (+ 8)
        uint32_t hf[S::KB][S::NB][2];
#pragma unroll
        for (int kb = 0; kb < S::KB; ++kb)
#pragma unroll
          for (int nbk = 0; nbk < S::NB; ++nbk)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int x = 16 * kb + 8 * q;              // + 2 t4
              if (x + 8 <= W || x + 2 * t4 < W) {         // columns beyond W (W = 8 mod 16) meet zero taps; keep them finite
                const float2 v = plane_pair<T>(plane + (8 * nbk + gg) * W + x + 2 * t4);
                float lo, hi;   // (v.x, v.y) * sc + off as one two-wide FFMA2
                wf2_split(wf2_fma(wf2_make(v.x, v.y), sc2, off2), lo, hi);
                hf[kb][nbk][q] = pack_h2(lo, hi);
              } else {
                hf[kb][nbk][q] = 0u;
              }
            }

        if (lane == 0) cand[kWCand] = 0;
        __syncwarp();
        // Two passes over the blocks of 16 output columns.  Pass 0 computes every block and keeps only the lane's
        // maximum per block; the warp-wide maximum then fixes the candidate threshold, and pass 1 recomputes just the
        // blocks that reach it (one, rarely two) and lists their pixels above the threshold.  (Listing against a
        // running maximum in a single pass does not work here: the band is ~0.5 % of the map's range, so the flat
        // background of the blocks visited before the blob would be listed wholesale.)
        float gm = -INFINITY, thr = INFINITY;
        float mblk[S::MB];
        unsigned need = 0u;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
          for (int mb = 0; mb < S::MB; ++mb) {
            if (pass == 1 && !((need >> mb) & 1u)) continue;
            // GEMM 1: Y^T[16 mb + ..][y] = sum_x Tx[x'][x] h~[y][x]; only |mb - kb| <= 1 tiles are non-zero (radius <= 9)
            float acc1[S::NB][4];
#pragma unroll
            for (int nbk = 0; nbk < S::NB; ++nbk)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc1[nbk][c] = 0.0f;
#pragma unroll
            for (int kb = 0; kb < S::KB; ++kb) {
              if (kb < mb - 1 || kb > mb + 1) continue;
              const uint4 a4 = __ldg(t1 + (mb * S::KB + kb) * 32 + lane);
              const uint32_t a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int nbk = 0; nbk < S::NB; ++nbk) mma16816(acc1[nbk], a, hf[kb][nbk][0], hf[kb][nbk][1]);
            }
            // the accumulators of n-blocks (2 j, 2 j + 1) are the A fragment of k-block j of GEMM 2
            uint32_t ya[S::KB2][4];
#pragma unroll
            for (int j = 0; j < S::KB2; ++j) {
              ya[j][0] = pack_h2(acc1[2 * j][0], acc1[2 * j][1]);
              ya[j][1] = pack_h2(acc1[2 * j][2], acc1[2 * j][3]);
              ya[j][2] = pack_h2(acc1[2 * j + 1][0], acc1[2 * j + 1][1]);
              ya[j][3] = pack_h2(acc1[2 * j + 1][2], acc1[2 * j + 1][3]);
            }
            // GEMM 2: Z^T[x'][y'] = sum_y Y^T[x'][y] Ty[y'][y]; tiles with |kb2 - nbp| <= 1 only
            float z[S::NB][4];
#pragma unroll
            for (int nbk = 0; nbk < S::NB; ++nbk)
#pragma unroll
              for (int c = 0; c < 4; ++c) z[nbk][c] = 0.0f;
#pragma unroll
            for (int nbp = 0; nbp < S::NBP; ++nbp)
#pragma unroll
              for (int kb2 = 0; kb2 < S::KB2; ++kb2) {
                if (kb2 < nbp - 1 || kb2 > nbp + 1) continue;
                const uint4 b4 = __ldg(t2 + (kb2 * S::NBP + nbp) * 32 + lane);
                mma16816(z[2 * nbp], ya[kb2], b4.x, b4.y);
                mma16816(z[2 * nbp + 1], ya[kb2], b4.z, b4.w);
              }
            // z[nb][c]: column x' = 16 mb + g + 8 (c >> 1), row y' = 8 nb + 2 t + (c & 1)
            const bool hi_ok = 16 * mb + 8 + 8 <= W;   // the rows g + 8 of the last block lie beyond W when W = 8 mod 16
            if (pass == 0) {
              float m = -INFINITY;
#pragma unroll
              for (int nbk = 0; nbk < S::NB; ++nbk) {
                m = fmaxf(m, fmaxf(z[nbk][0], z[nbk][1]));
                if (hi_ok) m = fmaxf(m, fmaxf(z[nbk][2], z[nbk][3]));
              }
              mblk[mb] = m;
              gm = fmaxf(gm, m);
              if (kDebug && dbg_prefilter) {
                const float inv = 1.0f / sc;
#pragma unroll
                for (int nbk = 0; nbk < S::NB; ++nbk)
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    if (c < 2 || hi_ok)
                      dbg_prefilter[static_cast<size_t>(hm) * HW + (8 * nbk + 2 * t4 + (c & 1)) * W + 16 * mb + gg + 8 * (c >> 1)] =
                          z[nbk][c] * inv + vmin;
              }
            } else if (mblk[mb] >= thr) {
              // bit (4 nb + c) of the lane's masks (32 bits each: n-blocks 0..7 and 8..): value (nb, c) is a candidate.
              // The listing loop below is ONE small body: one compare-and-push site per value was 1.2 k instructions of
              // rarely executed code in the middle of the hot path (instruction-fetch stalls, capture r02c).
              unsigned mask[(S::NB + 7) / 8];
#pragma unroll
              for (int w = 0; w < (S::NB + 7) / 8; ++w) mask[w] = 0u;
#pragma unroll
              for (int nbk = 0; nbk < S::NB; ++nbk)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if ((c < 2 || hi_ok) && z[nbk][c] >= thr) mask[nbk >> 3] |= 1u << (4 * (nbk & 7) + c);
#pragma unroll
              for (int w = 0; w < (S::NB + 7) / 8; ++w) {
                unsigned m = mask[w];
                while (m) {
                  const int e = __ffs(m) - 1;
                  m &= m - 1u;
                  const int nbk = 8 * w + (e >> 2), c = e & 3;
                  const int s = atomicAdd(&cand[kWCand], 1);
                  if (s < kWCand) cand[s] = (8 * nbk + 2 * t4 + (c & 1)) * W + 16 * mb + gg + 8 * (c >> 1);
                }
              }
            }
          }
          if (pass == 0) {
            gm = warp_max(gm);
            thr = gm - band;
#pragma unroll
            for (int mb = 0; mb < S::MB; ++mb) need |= (__any_sync(0xffffffffu, mblk[mb] >= thr) ? 1u : 0u) << mb;
          }
        }
        __syncwarp();
        const int count = cand[kWCand];
        handed_over = count > kWCand;
        if (kDebug && tl && lane == 0) { tl[4] = clock64(); tl_count = count; }
        if (kDark && !handed_over) {
          // ---- G (kDark): the blurred maximum = the largest exact blur among the candidates; the blur at the raw peak
          // and its six DARK neighbours, edge-clamped (codec.py:346-359).  Pixel list: candidates, then the stencil.
          const int py = fast_div(p0, geo.div_W), px = p0 - py * W;
          if (lane < 7) {
            const int dxs = (0x2424 >> (2 * lane)) & 3, dys = (0x2640 >> (2 * lane)) & 3;   // 0, +1, -1 (as 2) per stencil point
            const int yy = min(max(py + (dys == 2 ? -1 : dys), 0), H - 1), xx = min(max(px + (dxs == 2 ? -1 : dxs), 0), W - 1);
            cand[kWCand + 4 + lane] = yy * W + xx;   // the stencil pixels live behind the candidate list
          }
          __syncwarp();
          float bmax = -INFINITY, sten = 0.0f;   // lane q < 7 keeps stencil value q
          const int n_pix = count + 7;
          auto pixel = [&](int q) { return q < count ? cand[q] : cand[kWCand + 4 + min(q - count, 6)]; };
          for (int q = 0; q < n_pix; q += 2) {
            const float2 v = dark_blur2<T>(plane, blur_taps, dk.ksize, H, W, pixel(q), pixel(q + 1), lane);
            if (q < count) bmax = fmaxf(bmax, v.x); else if (lane == q - count) sten = v.x;
            if (q + 1 < count) bmax = fmaxf(bmax, v.y); else if (q + 1 < n_pix && lane == q + 1 - count) sten = v.y;
          }
          const float ratio = __fdiv_rn(vmax, __fadd_rn(bmax, 1e-12f));   // codec.py:312
          const float lg = dark_log(sten, ratio);
          float b[7];
#pragma unroll
          for (int q = 0; q < 7; ++q) b[q] = __shfl_sync(0xffffffffu, lg, q);
          if (lane == 0) dark_shift(b, px, py, dark_x, dark_y);
        }
        if (!kDark && !handed_over) {
          // ---- G: exact values of the candidates and of the winner's four neighbours
          if (count == 1) {
            best = cand[0];
          } else {
            best_val = -INFINITY; best = 0x7fffffff;
            if (count == 2) {
              float ev[2];
              mma_exact_n<T, 2>(plane, w2dk, H, W, r, cand, 2, geo.div_W, lane, ev);
              argmax_combine(best_val, best, ev[0], cand[0]);
              argmax_combine(best_val, best, ev[1], cand[1]);
            } else {
              for (int q = 0; q < count; q += 4) {
                float ev[4];
                mma_exact_n<T, 4>(plane, w2dk, H, W, r, cand + q, min(4, count - q), geo.div_W, lane, ev);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (q + c < count) argmax_combine(best_val, best, ev[c], cand[q + c]);
              }
            }
          }
          const int by = fast_div(best, geo.div_W), bx = best - by * W;
          interior = bx > 0 && bx < W - 1 && by > 0 && by < H - 1;
          if (interior) {
            float ev[5];
            mma_exact5<T>(plane, w2dk, H, W, r, by, bx, lane, ev);
            best_val = ev[0];
            nb[0] = ev[1]; nb[1] = ev[2]; nb[2] = ev[3]; nb[3] = ev[4];
          }
          score = plane_value<T>(plane, best);
        }
      }
    }

    if (kDebug && tl && lane == 0) tl[5] = clock64();
    // ---- H: outputs (thread 0: x, thread 1: y; heatmap.py:136-165 in float32, the reference's operation order) or
    // hand-over to the general kernel
    if (handed_over) {
      if (lane == 0) retry_list[atomicAdd(retry_count, 1u)] = hm;
    } else if (kDark) {
      if (lane == 0) {
        const bool empty = !(vmax > 0.0f);   // locs[vals <= 0] = -1 (heatmap.py:46): the sentinel is kept
        const int py = fast_div(p0, geo.div_W), px = p0 - py * W;
        if (dk.peaks) {
          dk.peaks[static_cast<size_t>(hm) * 2] = empty ? -1.0f : static_cast<float>(px);
          dk.peaks[static_cast<size_t>(hm) * 2 + 1] = empty ? -1.0f : static_cast<float>(py);
        }
        vals[hm] = vmax;
        locs[static_cast<size_t>(hm) * 2] = dark_x;
        locs[static_cast<size_t>(hm) * 2 + 1] = dark_y;
        if (keypoints) {   // codec.py:541
          keypoints[static_cast<size_t>(hm) * 2] = static_cast<double>(dark_x) / static_cast<double>(W - 1) * p.input_w;
          keypoints[static_cast<size_t>(hm) * 2 + 1] = static_cast<double>(dark_y) / static_cast<double>(H - 1) * p.input_h;
        }
      }
    } else if (lane < 2) {
      const int by = fast_div(best, geo.div_W), bx = best - by * W;
      const bool is_y = lane == 1;
      float f = static_cast<float>(is_y ? by : bx);
      if (interior) {
        const float lo = is_y ? nb[2] : nb[0], hi = is_y ? nb[3] : nb[1], c = best_val;   // left/up, right/down
        const float g = __fmul_rn(__fsub_rn(hi, lo), 0.5f);   // == / 2 (heatmap.py:139,146): halving is exact in binary floating point
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

    if (kDebug && tl && lane == 0) {
      unsigned smid;
      asm("mov.u32 %0, %%smid;" : "=r"(smid));
      tl[6] = clock64();
      tl[7] = (static_cast<unsigned long long>(smid) << 32) | (static_cast<unsigned long long>(tl_count) << 8) | static_cast<unsigned>(it & 255);
    }
    // ---- next heatmap: the warp is done with the plane, lane 0 starts the copy (an L2 hit by now)
    __syncwarp();
    if (lane == 0 && next_item < N) {
      fence_proxy_async();
      mbar_expect_tx(bar, geo.plane_bytes);
      tma_load_1d(slot, heatmaps + static_cast<size_t>(next_hm) * HW, geo.plane_bytes, bar);
    }
    cur_item = next_item;
    next_item = dynamic ? static_cast<int>(min(__shfl_sync(0xffffffffu, pulled_raw, 0) + 2u * nwarps, static_cast<unsigned>(N)))
                        : min(next_item + nwarps, N);
  }
}

// Fast path of the OKS heatmap loss, float32, two heatmaps per thread with packed arithmetic.
//
// Same algorithm, staging and register-rolling row pipeline as pp_loss_fast.cuh (read that header first).
// That kernel is instruction-issue bound (capture J: 70 thread instructions per pixel against ~66 at the HBM
// roofline, FMA pipe 44 % busy).  sm_100a has two-wide float32 instructions (PTX add/sub/mul/fma.rn.f32x2 ->
// SASS FADD2 / FMUL2 / FFMA2, each a pair of ordinary IEEE operations), so here a thread runs the SAME strip
// of TWO consecutive heatmaps in lock step: every value of the pipeline is a pair (heatmap A, heatmap B), the
// geometry (strip, rows, border predicates) is shared, and the arithmetic instruction count per pixel halves.
// The two halves of a pair come from different shared-memory planes, so loads are scalar (two 32-bit loads
// fill one pair); the shared-memory wavefront count per pixel is unchanged.
#pragma once

#include "pp_common.cuh"
#include "pp_loss_fast.cuh"

namespace pp_loss_pair {

using namespace pp;
using pp_loss_fast::FastArgs;

typedef unsigned long long f2;   // (lo, hi) = (heatmap A, heatmap B)

__device__ __forceinline__ f2 mk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void un(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

struct RowState2 {
  f2 hd[3][6], hs[3][6];   // row factors of the Sobel pair, rows r-2, r-1, r
  f2 dP[3][4], sQ[3][4];   // row factors of the adjoint stencil, rows r-3, r-2, r-1
};

struct Coef2 {
  f2 c2;             // 2 lw w_s u m   (P = c2 gx, Q = c2 gy)
  f2 k_a, k_b, k_g;  // direct gradient = k_a + k_b t + k_g (o - t)
  f2 a_o, a_t, two;
  bool has_mse;
};

struct Sums2 {
  f2 se, so, sm;
  float tmin, tmax;
};

// One row step of the pipeline for a pair of heatmaps; PH = step index mod 3 selects the rotating slots.
template <bool kFwd, bool kGrad, int PH>
__device__ __forceinline__ void row_step2(RowState2& st, Sums2& sums, const Coef2& cf, int q, int y0, int y1, int H, int W,
                                          int x0, bool left_ok, bool right_ok, const float* __restrict__ pa,
                                          const float* __restrict__ pb, const float* __restrict__ ta,
                                          const float* __restrict__ tb, float* __restrict__ ga, float* __restrict__ gb,
                                          bool store_b) {
  constexpr int cur = PH, p1 = (PH + 2) % 3, p2 = (PH + 1) % 3;
  const int r = y0 - 2 + q;

  // 1. row r of `output` -> row factors of the Sobel pair at columns x0-1 .. x0+4
  f2 av[8];
  if (r >= 0 && r < H) {
    const float* ra = pa + r * W + x0 - 2;
    const float* rb = pb + r * W + x0 - 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = mk(ra[j], rb[j]);
    if (!left_ok) { av[0] = 0ull; av[1] = 0ull; }
    if (!right_ok) { av[6] = 0ull; av[7] = 0ull; }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = 0ull;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    st.hd[cur][j] = sub2(av[j], av[j + 2]);
    st.hs[cur][j] = fma2(cf.two, av[j + 1], add2(av[j], av[j + 2]));
  }

  // 2. gx, gy at row r-1, columns x0-1 .. x0+4
  const int rm = r - 1;
  f2 gx[6], gy[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    gx[j] = fma2(cf.two, st.hd[p1][j], add2(st.hd[p2][j], st.hd[cur][j]));
    gy[j] = sub2(st.hs[p2][j], st.hs[cur][j]);
  }
  if (kFwd && rm >= y0 && rm < y1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sums.se = add2(sums.se, fma2(gx[1 + i], gx[1 + i], mul2(gy[1 + i], gy[1 + i])));
  }
  if (kGrad) {
    const f2 cr = (rm >= 0 && rm < H) ? cf.c2 : 0ull;
    f2 P[6], Q[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { P[j] = mul2(cr, gx[j]); Q[j] = mul2(cr, gy[j]); }
    if (!left_ok) { P[0] = 0ull; Q[0] = 0ull; }     // column -1 is outside the map
    if (!right_ok) { P[5] = 0ull; Q[5] = 0ull; }    // column W is outside the map
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      st.dP[p1][i] = sub2(P[i], P[i + 2]);
      st.sQ[p1][i] = fma2(cf.two, Q[i + 1], add2(Q[i], Q[i + 2]));
    }
  }

  // 3. emit row r-2
  const int ro = r - 2;
  if (ro >= y0 && ro < y1) {
    const int off = ro * W + x0;
    float g_a[4], g_b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float t_a = ta[off + i], t_b = tb[off + i];
      const f2 o = mk(pa[off + i], pb[off + i]), t = mk(t_a, t_b);
      if (kFwd) {
        // oks = a_o o + a_t t - o t = o (a_o - t) + a_t t   ("minus": exactly out * (1 - tgt), loss.py:93)
        sums.so = add2(sums.so, fma2(o, sub2(cf.a_o, t), mul2(cf.a_t, t)));
        if (cf.has_mse) { const f2 d = sub2(o, t); sums.sm = fma2(d, d, sums.sm); }
        sums.tmin = fminf(sums.tmin, fminf(t_a, t_b));
        sums.tmax = fmaxf(sums.tmax, fmaxf(t_a, t_b));
      }
      if (kGrad) {
        f2 direct = fma2(cf.k_b, t, cf.k_a);
        if (cf.has_mse) direct = fma2(cf.k_g, sub2(o, t), direct);
        const f2 sx = fma2(cf.two, st.dP[p2][i], add2(st.dP[cur][i], st.dP[p1][i]));   // rows r-3, r-1, r-2
        const f2 sy = sub2(st.sQ[cur][i], st.sQ[p1][i]);                              // rows r-3, r-1
        un(sub2(sub2(direct, sx), sy), g_a[i], g_b[i]);
      }
    }
    if (kGrad) {
      pp_loss_fast::store_grad4<float>(ga + off, g_a);
      if (store_b) pp_loss_fast::store_grad4<float>(gb + off, g_b);
    }
  }
}

// a.G = heatmaps per unit (even); thread group g of a CTA owns heatmaps (2g, 2g + 1) of the unit.
template <bool kFwd, bool kGrad>
__global__ void __launch_bounds__(256)
oks_loss_pair_kernel(FastArgs a) {
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ double red[8];
  __shared__ int red_flag;

  const int tid = threadIdx.x;
  const int H = a.H, W = a.W;
  const long long HW = static_cast<long long>(H) * W;
  const int per = a.strips * a.segs;           // threads per pair of heatmaps
  const int g = tid / per, local = tid - g * per;
  const int sx = local % a.strips, sy = local / a.strips;
  const int x0 = sx * 4, y0 = sy * a.T, y1 = min(y0 + a.T, H);
  const bool left_ok = sx > 0, right_ok = sx < a.strips - 1;
  const float* out = static_cast<const float*>(a.output);
  const float* tgt_all = static_cast<const float*>(a.target);
  float* grad_all = static_cast<float*>(a.grad);
  auto stage_of = [&](int s) { return reinterpret_cast<const float*>(stage_mem + static_cast<size_t>(s) * a.stage_bytes + 16); };
  auto tgt_stage_of = [&](int s) { return reinterpret_cast<const float*>(stage_mem + static_cast<size_t>(s) * a.stage_bytes + a.tgt_off); };
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
    mbar_expect_tx(&bars[s], 2 * bytes);
    tma_load_1d(const_cast<float*>(stage_of(s)), out + un * a.G * HW, bytes, &bars[s]);
    tma_load_1d(const_cast<float*>(tgt_stage_of(s)), tgt_all + un * a.G * HW, bytes, &bars[s]);
  };
  if (tid == 0 && unit < units) fetch_unit(unit, 0);

  double acc = 0.0;
  Sums2 sums{0ull, 0ull, 0ull, INFINITY, -INFINITY};
  const int nsteps = (y1 - y0) + 4;

  for (int it = 0; unit < units; unit += gridDim.x, ++it) {
    const int s = (a.stages == 2) ? (it & 1) : 0;
    const long long nxt = unit + gridDim.x;
    if (a.stages == 2 && tid == 0 && nxt < units) fetch_unit(nxt, s ^ 1);
    const long long hm_a = unit * a.G + 2 * g;
    const bool active = 2 * g < a.G && hm_a < a.N && sy < a.segs;
    const bool has_b = active && hm_a + 1 < a.N;
    const float m_a = (active && a.kp_weights) ? a.kp_weights[hm_a] : 1.0f;
    const float m_b = (has_b && a.kp_weights) ? a.kp_weights[hm_a + 1] : (has_b ? 1.0f : 0.0f);
    Coef2 cf;
    cf.c2 = mk(2.0f * a.lw * a.w_s * u * m_a, 2.0f * a.lw * a.w_s * u * m_b);
    cf.k_a = mk(a.lw * u * m_a * a.w_o * a.d_a, a.lw * u * m_b * a.w_o * a.d_a);
    cf.k_b = mk(a.lw * u * m_a * a.w_o * a.d_b, a.lw * u * m_b * a.w_o * a.d_b);
    cf.k_g = mk(2.0f * a.lw * u * m_a * a.w_g, 2.0f * a.lw * u * m_b * a.w_g);
    cf.a_o = mk(a.a_o, a.a_o);
    cf.a_t = mk(a.a_t, a.a_t);
    cf.two = mk(2.0f, 2.0f);
    cf.has_mse = a.w_g != 0.0f;

    RowState2 st;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { st.hd[i][j] = 0ull; st.hs[i][j] = 0ull; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { st.dP[i][j] = 0ull; st.sQ[i][j] = 0ull; }
    }
    sums.se = sums.so = sums.sm = 0ull;

    mbar_wait(&bars[s], (a.stages == 2) ? ((it >> 1) & 1) : (it & 1));
    if (active) {
      // an odd tail has no heatmap B: its lanes recompute heatmap A (finite values), contribute nothing, store nothing
      const size_t ia = static_cast<size_t>(2 * g) * HW, ib = has_b ? ia + HW : ia;
      const float* pa = stage_of(s) + ia;
      const float* pb = stage_of(s) + ib;
      const float* ta = tgt_stage_of(s) + ia;
      const float* tb = tgt_stage_of(s) + ib;
      float* ga = kGrad ? grad_all + hm_a * HW : nullptr;
      float* gb = kGrad ? ga + HW : nullptr;
      for (int q = 0; q < nsteps; q += 3) {
        row_step2<kFwd, kGrad, 0>(st, sums, cf, q, y0, y1, H, W, x0, left_ok, right_ok, pa, pb, ta, tb, ga, gb, has_b);
        if (q + 1 < nsteps)
          row_step2<kFwd, kGrad, 1>(st, sums, cf, q + 1, y0, y1, H, W, x0, left_ok, right_ok, pa, pb, ta, tb, ga, gb, has_b);
        if (q + 2 < nsteps)
          row_step2<kFwd, kGrad, 2>(st, sums, cf, q + 2, y0, y1, H, W, x0, left_ok, right_ok, pa, pb, ta, tb, ga, gb, has_b);
      }
      if (kFwd) {
        // per-pixel loss = (w_s e + w_o oks + w_g mse) m lw (loss.py:122-127, 143), summed per strip
        float se_a, se_b, so_a, so_b, sm_a, sm_b;
        un(sums.se, se_a, se_b);
        un(sums.so, so_a, so_b);
        un(sums.sm, sm_a, sm_b);
        acc += static_cast<double>((a.w_s * se_a + a.w_o * so_a + a.w_g * sm_a) * (m_a * a.lw));
        if (has_b) acc += static_cast<double>((a.w_s * se_b + a.w_o * so_b + a.w_g * sm_b) * (m_b * a.lw));
      }
    }
    __syncthreads();  // every thread is done with stage s before it is refilled
    if (a.stages == 1 && tid == 0 && nxt < units) fetch_unit(nxt, 0);
  }

  if (kFwd) {
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    if (a.range_flag && (sums.tmin < 0.0f || sums.tmax > 1.0f)) red_flag = 1;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
      a.partials[blockIdx.x] = t;
      if (a.range_flag && red_flag) atomicOr(a.range_flag, 1);
    }
  }
}

}  // namespace pp_loss_pair

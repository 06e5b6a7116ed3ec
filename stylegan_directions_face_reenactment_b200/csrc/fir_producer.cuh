// FIR pass of an upsampling layer folded into the CONSUMER: producer warps of the following 3x3 convolution
// (modconv_halo_sm100.cu, fused mode) build the halo tile of its A operand straight from the fp32 parity planes the
// scatter GEMM left in HBM — same math as up_finish_kernel (Blur / upfirdn2d 4x4 FIR, libs/gan/StyleGAN2/model.py:72-88,
// 256-257; demodulation, NoiseInjection, FusedLeakyReLU :282-287,331-337; next style multiply + hi/lo split) — and write
// it into shared memory in the UMMA operand layout [plane][chunk][row][pixel][8 ch].  The activation of the up layer is
// then never written to or read from HBM (2 x 4 B per element) and the HBM-bound pass hides behind the consumer's MMAs.
#pragma once
#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct FP4 {
  float2 a, b;
};
__device__ __forceinline__ FP4 fp4_load(const float* p, bool ok) {
  if (!ok) return FP4{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return FP4{make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ FP4 fp4_mul(float w, const FP4& x) {
  const float2 ww = make_float2(w, w);
  return FP4{__fmul2_rn(ww, x.a), __fmul2_rn(ww, x.b)};
}
__device__ __forceinline__ FP4 fp4_fma(float w, const FP4& x, const FP4& acc) {
  const float2 ww = make_float2(w, w);
  return FP4{__ffma2_rn(ww, x.a, acc.a), __ffma2_rn(ww, x.b, acc.b)};
}
__device__ __forceinline__ FP4 fp4_shfl(const FP4& x, int src_lane) {
  FP4 r;
  r.a.x = __shfl_sync(0xffffffffu, x.a.x, src_lane);
  r.a.y = __shfl_sync(0xffffffffu, x.a.y, src_lane);
  r.b.x = __shfl_sync(0xffffffffu, x.b.x, src_lane);
  r.b.y = __shfl_sync(0xffffffffu, x.b.y, src_lane);
  return r;
}

// One warp fills, for ONE 8-channel chunk, the halo-tile rows produced by plane rows m_begin .. m_begin + iters - 1 (plane
// row m yields output rows 2m and 2m+1) of the tile whose halo origin is output pixel (Y0, X0), both odd.
//   lanes 0-15 / 16-31: channels 0-3 / 4-7 of the chunk; lane l & 15 = plane column n_first - 1 + l (first and last used
//   lane are halo-only: they feed their neighbours' horizontal taps through shuffles).
// The parity planes are STAGED IN SHARED MEMORY by TMA (reading them straight from global memory is latency bound: 8
// producer warps cannot keep enough loads in flight; measured 2x slower).  A plane stage holds 16 channels (4 groups of 4) of the tile's
// plane window: [plane (oe,ee,eo,oo)][group][row: 12 = plane rows m_first-1 .. m_first+10][col: PC][4 floats], out-of-range
// rows / columns zero-filled by TMA.  The warp handles groups g0, g0+1 (= one 8-channel chunk) and `iters` plane rows from
// stage row r_begin (plane row m_begin).
// bf16 hi/lo split of two values with paired conversions (as up_finish_kernel)
__device__ __forceinline__ void fir_split_pair(float v0, float v1, int fmt, uint32_t& hi, uint32_t& lo) {
  if (fmt == kFmtBF16) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - h0, v1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    split2(v0, v1, fmt, hi, lo);
  }
}

template <int kHW, int kHH, int FMT, int PC>
__device__ __forceinline__ void fir_produce_chunk_smem(const FusedFirParams& f, const float* sc, int b, int chunk, int Y0, int X0,
                                                       const uint8_t* stage, int g0, int m_begin, int r_begin, int iters,
                                                       uint8_t* dst, uint32_t lo_off, int lane) {
  const int half = lane >> 4, l = lane & 15;
  const int Ho = 2 * f.Hin, Wo = 2 * f.Win;
  const int n_first = (X0 - 1) >> 1;
  const int n = n_first - 1 + l;
  constexpr int kCols = kHW / 2 + 1;
  static_assert(PC >= kCols + 2 && PC <= 16, "plane window columns");
  const bool lane_out = l >= 1 && l <= kCols;
  constexpr uint32_t kRowB = PC * 16, kGroupB = 12 * kRowB, kPlaneB = 4 * kGroupB;
  const uint8_t* s_oe = stage + static_cast<uint32_t>(g0 + half) * kGroupB + static_cast<uint32_t>(min(l, PC - 1)) * 16;
  auto ld = [&](int pl, int r) {
    const float4 v = *reinterpret_cast<const float4*>(s_oe + pl * kPlaneB + static_cast<uint32_t>(r) * kRowB);
    return FP4{make_float2(v.x, v.y), make_float2(v.z, v.w)};
  };

  const float gx0 = sc[4], gx1 = sc[5], gx2 = sc[6], gx3 = sc[7];
  const float s_oe_ = f.plane_scale[0], s_ee = f.plane_scale[1], s_eo = f.plane_scale[2], s_oo = f.plane_scale[3];
  const float e0o = sc[0] * s_oe_, e1e = sc[1] * s_ee, e2o = sc[2] * s_oe_, e3e = sc[3] * s_ee;
  const float e0e = sc[0] * s_ee, e1o = sc[1] * s_oe_, e2e = sc[2] * s_ee, e3o = sc[3] * s_oe_;
  const float o0o = sc[0] * s_oo, o1e = sc[1] * s_eo, o2o = sc[2] * s_oo, o3e = sc[3] * s_eo;
  const float o0e = sc[0] * s_eo, o1o = sc[1] * s_oo, o2e = sc[2] * s_eo, o3o = sc[3] * s_oo;

  const int c0 = chunk * 8 + half * 4;
  float2 d01, d23, b01, b23, g01, g23;
  {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(f.demod + static_cast<size_t>(b) * f.C + c0));
    d01 = make_float2(d0.x, d0.y); d23 = make_float2(d0.z, d0.w);
    float bi[4], g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bi[e] = f.bias ? __ldg(f.bias + c0 + e) : 0.f;
      g[e] = (f.s2 ? __ldg(f.s2 + static_cast<size_t>(b) * f.C + c0 + e) : f.act_gain) * f.out_scale;
    }
    b01 = make_float2(bi[0], bi[1]); b23 = make_float2(bi[2], bi[3]);
    g01 = make_float2(g[0], g[1]); g23 = make_float2(g[2], g[3]);
  }
  const float nw = f.noise ? __ldg(f.noise_w) : 0.f;
  const float slope = f.act ? 0.2f : 1.f;
  const float2 slope2 = make_float2(slope, slope);
  const float* nz_base = f.noise ? f.noise + static_cast<size_t>(b) * f.noise_bstride : nullptr;
  const int X = 2 * n + half;
  const bool x_in_tile = lane_out && X >= X0 && X < X0 + kHW;
  const bool x_in_image = X >= 0 && X < Wo;
  const int Xn = 2 * n;
  const bool nzx_ok = nz_base && n >= 0 && Xn + 1 < Wo;
  auto load_noise = [&](int m, float2* out) {
    out[0] = out[1] = make_float2(0.f, 0.f);
    if (nzx_ok) {
      const int Y = 2 * m;
      if (Y >= 0 && Y < Ho) out[0] = __ldg(reinterpret_cast<const float2*>(nz_base + static_cast<size_t>(Y) * Wo + Xn));
      if (Y + 1 >= 0 && Y + 1 < Ho) out[1] = __ldg(reinterpret_cast<const float2*>(nz_base + static_cast<size_t>(Y + 1) * Wo + Xn));
    }
  };

  int m = m_begin, r = r_begin;
  FP4 oe_m1 = ld(0, r - 1), oo_m1 = ld(3, r - 1);
  FP4 ee_0 = ld(1, r), eo_0 = ld(2, r), oe_0 = ld(0, r), oo_0 = ld(3, r);
  float2 nz_n[2];
  load_noise(m, nz_n);
#pragma unroll 1
  for (int it = 0; it < iters; ++it, ++m, ++r) {
    const FP4 ee_1 = ld(1, r + 1), eo_1 = ld(2, r + 1), oe_1 = ld(0, r + 1), oo_1 = ld(3, r + 1);
    const float2 nz_c[2] = {nz_n[0], nz_n[1]};
    if (it + 1 < iters) load_noise(m + 1, nz_n);
    FP4 ve[2], vo[2];
    ve[0] = fp4_fma(e3e, ee_1, fp4_fma(e2o, oe_0, fp4_fma(e1e, ee_0, fp4_mul(e0o, oe_m1))));
    ve[1] = fp4_fma(e3o, oe_1, fp4_fma(e2e, ee_1, fp4_fma(e1o, oe_0, fp4_mul(e0e, ee_0))));
    vo[0] = fp4_fma(o3e, eo_1, fp4_fma(o2o, oo_0, fp4_fma(o1e, eo_0, fp4_mul(o0o, oo_m1))));
    vo[1] = fp4_fma(o3o, oo_1, fp4_fma(o2e, eo_1, fp4_fma(o1o, oo_0, fp4_mul(o0e, eo_0))));
    oe_m1 = oe_0; oo_m1 = oo_0;
    ee_0 = ee_1; eo_0 = eo_1; oe_0 = oe_1; oo_0 = oo_1;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const int Y = 2 * m + py;
      const int rr = Y - Y0;
      if (rr < 0 || rr >= kHH) continue;                      // warp-uniform: first / last plane row feed one tile row only
      const FP4 vo_l = fp4_shfl(vo[py], lane - 1);
      const FP4 ve_r = fp4_shfl(ve[py], lane + 1);
      const FP4 vo_r = fp4_shfl(vo[py], lane + 1);
      const FP4 z0 = fp4_fma(gx3, ve_r, fp4_fma(gx2, vo[py], fp4_fma(gx1, ve[py], fp4_mul(gx0, vo_l))));
      const FP4 z1 = fp4_fma(gx3, vo_r, fp4_fma(gx2, ve_r, fp4_fma(gx1, vo[py], fp4_mul(gx0, ve[py]))));
      const float2 nn = nz_c[py];
      const float2 n0 = make_float2(nw * nn.x, nw * nn.x), n1 = make_float2(nw * nn.y, nw * nn.y);
      const float2 nz0 = __fadd2_rn(b01, n0), nz1 = __fadd2_rn(b01, n1), nz2 = __fadd2_rn(b23, n0), nz3 = __fadd2_rn(b23, n1);
      float2 t0a = __ffma2_rn(z0.a, d01, nz0), t0b = __ffma2_rn(z0.b, d23, nz2);
      float2 t1a = __ffma2_rn(z1.a, d01, nz1), t1b = __ffma2_rn(z1.b, d23, nz3);
      {
        const float2 s0a = __fmul2_rn(t0a, slope2), s0b = __fmul2_rn(t0b, slope2);
        const float2 s1a = __fmul2_rn(t1a, slope2), s1b = __fmul2_rn(t1b, slope2);
        t0a = make_float2(fmaxf(t0a.x, s0a.x), fmaxf(t0a.y, s0a.y)); t0b = make_float2(fmaxf(t0b.x, s0b.x), fmaxf(t0b.y, s0b.y));
        t1a = make_float2(fmaxf(t1a.x, s1a.x), fmaxf(t1a.y, s1a.y)); t1b = make_float2(fmaxf(t1b.x, s1b.x), fmaxf(t1b.y, s1b.y));
      }
      const float2 u0a = __fmul2_rn(t0a, g01), u0b = __fmul2_rn(t0b, g23), u1a = __fmul2_rn(t1a, g01), u1b = __fmul2_rn(t1b, g23);
      uint32_t hi[2][2], lo[2][2];
      fir_split_pair(u0a.x, u0a.y, FMT, hi[0][0], lo[0][0]);
      fir_split_pair(u0b.x, u0b.y, FMT, hi[0][1], lo[0][1]);
      fir_split_pair(u1a.x, u1a.y, FMT, hi[1][0], lo[1][0]);
      fir_split_pair(u1b.x, u1b.y, FMT, hi[1][1], lo[1][1]);
      uint32_t rh[2], rl[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        rh[e] = __shfl_xor_sync(0xffffffffu, half ? hi[0][e] : hi[1][e], 16);
        rl[e] = __shfl_xor_sync(0xffffffffu, half ? lo[0][e] : lo[1][e], 16);
      }
      if (x_in_tile) {
        uint4 h4 = half ? make_uint4(rh[0], rh[1], hi[1][0], hi[1][1]) : make_uint4(hi[0][0], hi[0][1], rh[0], rh[1]);
        uint4 l4 = half ? make_uint4(rl[0], rl[1], lo[1][0], lo[1][1]) : make_uint4(lo[0][0], lo[0][1], rl[0], rl[1]);
        if (!(x_in_image && Y >= 0 && Y < Ho)) h4 = l4 = make_uint4(0u, 0u, 0u, 0u);     // the convolution's zero padding
        uint8_t* d = dst + (static_cast<uint32_t>(rr) * kHW + static_cast<uint32_t>(X - X0)) * 16;
        *reinterpret_cast<uint4*>(d) = h4;
        *reinterpret_cast<uint4*>(d + lo_off) = l4;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shuffle-free mapping (second version): the warp handles ONE 4-channel group and five plane rows; lane = (producing plane
// column c, row sub-group): kCols x RS lanes (10 x 3 or 6 x 5 of 32) each walk 1-2 plane rows.  The FIR runs horizontally
// first — every lane reads its three plane columns straight from the staged planes, so there are no halo lanes and no
// warp-collective operations — then vertically over a register window of the row results.
//   He_r(px0) = gx3 ee[r][n+1] + gx2 eo[r][n] + gx1 ee[r][n] + gx0 eo[r][n-1]      (even plane rows; px0: X = 2n)
//   He_r(px1) = gx3 eo[r][n+1] + gx2 ee[r][n+1] + gx1 eo[r][n] + gx0 ee[r][n]      (px1: X = 2n+1)
//   Ho_r likewise from oe / oo;  z(2m) = gy3 He[m+1] + gy2 Ho[m] + gy1 He[m] + gy0 Ho[m-1],
//                                z(2m+1) = gy3 Ho[m+1] + gy2 He[m+1] + gy1 Ho[m] + gy0 He[m]
// dst8: shared-memory address of [row 0][pixel 0] of the chunk this group belongs to, plus 8 for the chunk's upper half.
struct FH {          // row result of one row parity: two pixels x 4 channels
  FP4 p0, p1;
};

template <int kHW, int kHH, int FMT, int PC>
__device__ __forceinline__ void fir_produce_group_smem(const FusedFirParams& f, const float* sc, int b, int group, int Y0, int X0,
                                                       const uint8_t* stage, int g_in_stage, int m_h, int r_h, uint8_t* dst8,
                                                       uint32_t lo_off, int lane) {
  constexpr int kCols = kHW / 2 + 1;
  constexpr int RS = 32 / kCols;
  static_assert(PC >= kCols + 2, "plane window columns");
  const int c = lane % kCols, rsub = lane / kCols;
  if (rsub >= RS) return;                                     // spare lanes: nothing below is warp-collective
  constexpr int kBase = 5 / RS, kRem = 5 % RS;
  const int row_off = rsub * kBase + min(rsub, kRem);
  const int cnt = kBase + (rsub < kRem ? 1 : 0);
  const int Ho = 2 * f.Hin, Wo = 2 * f.Win;
  const int n = ((X0 - 1) >> 1) + c;                          // plane column; stage column = c + 1
  constexpr uint32_t kRowB = PC * 16, kGroupB = 12 * kRowB, kPlaneB = 4 * kGroupB;
  const uint8_t* s0 = stage + static_cast<uint32_t>(g_in_stage) * kGroupB + static_cast<uint32_t>(c) * 16;   // column n-1
  auto ld = [&](int pl, int r, int dc) {                      // plane pl, stage row r, column n - 1 + dc
    const float4 v = *reinterpret_cast<const float4*>(s0 + pl * kPlaneB + static_cast<uint32_t>(r) * kRowB + dc * 16);
    return FP4{make_float2(v.x, v.y), make_float2(v.z, v.w)};
  };
  const float gx0 = sc[4], gx1 = sc[5], gx2 = sc[6], gx3 = sc[7];
  const float gy0 = sc[0], gy1 = sc[1], gy2 = sc[2], gy3 = sc[3];
  const float s_oe = f.plane_scale[0], s_ee = f.plane_scale[1], s_eo = f.plane_scale[2], s_oo = f.plane_scale[3];
  // row result of an even (planes ee = 1, eo = 2) or odd (oe = 0, oo = 3) plane row r
  // horizontal taps with the plane scales folded in: [row parity][even-column plane | odd-column plane][tap]
  const float ce[2][4] = {{gx0 * s_ee, gx1 * s_ee, gx2 * s_ee, gx3 * s_ee}, {gx0 * s_oe, gx1 * s_oe, gx2 * s_oe, gx3 * s_oe}};
  const float co[2][4] = {{gx0 * s_eo, gx1 * s_eo, gx2 * s_eo, gx3 * s_eo}, {gx0 * s_oo, gx1 * s_oo, gx2 * s_oo, gx3 * s_oo}};
  auto hrow = [&](bool odd, int r) {
    const int pe = odd ? 0 : 1, po = odd ? 3 : 2;            // even-column / odd-column plane of this row parity
    const float* e = ce[odd ? 1 : 0];
    const float* o = co[odd ? 1 : 0];
    const FP4 e_n = ld(pe, r, 1), e_r = ld(pe, r, 2), o_l = ld(po, r, 0), o_n = ld(po, r, 1), o_r = ld(po, r, 2);
    FH h;
    h.p0 = fp4_fma(e[3], e_r, fp4_fma(o[2], o_n, fp4_fma(e[1], e_n, fp4_mul(o[0], o_l))));
    h.p1 = fp4_fma(o[3], o_r, fp4_fma(e[2], e_r, fp4_fma(o[1], o_n, fp4_mul(e[0], e_n))));
    return h;
  };

  const int c0 = group * 4;
  float2 d01, d23, b01, b23, g01, g23;
  {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(f.demod + static_cast<size_t>(b) * f.C + c0));
    d01 = make_float2(d0.x, d0.y); d23 = make_float2(d0.z, d0.w);
    float bi[4], g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bi[e] = f.bias ? __ldg(f.bias + c0 + e) : 0.f;
      g[e] = (f.s2 ? __ldg(f.s2 + static_cast<size_t>(b) * f.C + c0 + e) : f.act_gain) * f.out_scale;
    }
    b01 = make_float2(bi[0], bi[1]); b23 = make_float2(bi[2], bi[3]);
    g01 = make_float2(g[0], g[1]); g23 = make_float2(g[2], g[3]);
  }
  const float nw = f.noise ? __ldg(f.noise_w) : 0.f;
  const float slope = f.act ? 0.2f : 1.f;
  const float2 slope2 = make_float2(slope, slope);
  const float* nz_base = f.noise ? f.noise + static_cast<size_t>(b) * f.noise_bstride : nullptr;
  const int Xa = 2 * n;                                       // the lane's pixels: Xa (px0), Xa + 1 (px1)
  const int pa = Xa - X0;                                     // tile column of px0; px1 at pa + 1
  const bool ok0 = pa >= 0 && pa < kHW, ok1 = pa + 1 >= 0 && pa + 1 < kHW;
  const bool img0 = Xa >= 0 && Xa < Wo, img1 = Xa + 1 >= 0 && Xa + 1 < Wo;
  const bool nzx_ok = nz_base && Xa >= 0 && Xa + 1 < Wo;

  // epilogue of one pixel (4 channels): t = z * demod + noise + bias; leaky relu; * style; hi/lo split -> 8-byte stores
  auto emit = [&](const FP4& z, float nzv, int rr, int pcol, bool in_img) {
    const float2 nn = make_float2(nw * nzv, nw * nzv);
    float2 ta = __ffma2_rn(z.a, d01, __fadd2_rn(b01, nn)), tb = __ffma2_rn(z.b, d23, __fadd2_rn(b23, nn));
    const float2 sa = __fmul2_rn(ta, slope2), sb = __fmul2_rn(tb, slope2);
    ta = make_float2(fmaxf(ta.x, sa.x), fmaxf(ta.y, sa.y));
    tb = make_float2(fmaxf(tb.x, sb.x), fmaxf(tb.y, sb.y));
    const float2 ua = __fmul2_rn(ta, g01), ub = __fmul2_rn(tb, g23);
    uint32_t h0, l0, h1, l1;
    fir_split_pair(ua.x, ua.y, FMT, h0, l0);
    fir_split_pair(ub.x, ub.y, FMT, h1, l1);
    if (!in_img) h0 = l0 = h1 = l1 = 0u;                      // the convolution's zero padding
    uint8_t* d = dst8 + (static_cast<uint32_t>(rr) * kHW + static_cast<uint32_t>(pcol)) * 16;
    *reinterpret_cast<uint2*>(d) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(d + lo_off) = make_uint2(l0, l1);
  };

  int m = m_h + row_off, r = r_h + row_off;
  // the noise of all (at most four) output rows of this lane is requested before any arithmetic: an L2 round trip per
  // row in the loop below was the top stall of the first version (ncu: long scoreboard)
  float2 nzv[2 * (kBase + (kRem ? 1 : 0))];
#pragma unroll
  for (int q = 0; q < 2 * (kBase + (kRem ? 1 : 0)); ++q) {
    const int Y = 2 * m + q;
    nzv[q] = make_float2(0.f, 0.f);
    if (nzx_ok && q < 2 * cnt && Y >= 0 && Y < Ho) nzv[q] = __ldg(reinterpret_cast<const float2*>(nz_base + static_cast<size_t>(Y) * Wo + Xa));
  }
  FH ho_prev = hrow(true, r - 1), he_cur = hrow(false, r), ho_cur = hrow(true, r);
#pragma unroll
  for (int it = 0; it < kBase + (kRem ? 1 : 0); ++it, ++m, ++r) {
    if (it >= cnt) break;
    const FH he_next = hrow(false, r + 1), ho_next = hrow(true, r + 1);
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const int Y = 2 * m + py, rr = Y - Y0;
      if (rr < 0 || rr >= kHH) continue;
      const FH& a3 = py ? ho_next : he_next;                  // gy3 ..
      const FH& a2 = py ? he_next : ho_cur;
      const FH& a1 = py ? ho_cur : he_cur;
      const FH& a0 = py ? he_cur : ho_prev;
      const bool yin = Y >= 0 && Y < Ho;
      const float2 nz = nzv[2 * it + py];
      if (ok0) {
        const FP4 z = fp4_fma(gy3, a3.p0, fp4_fma(gy2, a2.p0, fp4_fma(gy1, a1.p0, fp4_mul(gy0, a0.p0))));
        emit(z, nz.x, rr, pa, yin && img0);
      }
      if (ok1) {
        const FP4 z = fp4_fma(gy3, a3.p1, fp4_fma(gy2, a2.p1, fp4_fma(gy1, a1.p1, fp4_mul(gy0, a0.p1))));
        emit(z, nz.y, rr, pa + 1, yin && img1);
      }
    }
    ho_prev = ho_cur; he_cur = he_next; ho_cur = ho_next;
  }
}

}  // namespace sgr

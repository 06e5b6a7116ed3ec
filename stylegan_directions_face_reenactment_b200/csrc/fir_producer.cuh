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
__device__ __forceinline__ FP4 fp4_mul(float w, const FP4& x) {
  const float2 ww = make_float2(w, w);
  return FP4{__fmul2_rn(ww, x.a), __fmul2_rn(ww, x.b)};
}
__device__ __forceinline__ FP4 fp4_fma(float w, const FP4& x, const FP4& acc) {
  const float2 ww = make_float2(w, w);
  return FP4{__ffma2_rn(ww, x.a, acc.a), __ffma2_rn(ww, x.b, acc.b)};
}

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

// The parity planes are STAGED IN SHARED MEMORY by TMA (reading them straight from global memory is latency bound: eight
// producer warps cannot keep enough loads in flight).  A plane stage holds 16 channels (4 groups of 4) of the tile's plane
// window: [plane (oe,ee,eo,oo)][group][row: 12 = plane rows m_first-1 .. m_first+10][col: PC][4 floats], out-of-range rows /
// columns zero-filled by TMA.
//
// Shuffle-free mapping (the first version used up_finish_kernel's: 16 columns per half-warp with halo lanes and shuffles
// for the horizontal taps — 62 % of the lanes productive, 1.6x the instructions): the warp handles ONE 4-channel group and
// five plane rows; lane = (producing plane
// column c, row sub-group): kCols x RS lanes (10 x 3 or 6 x 5 of 32) each walk 1-2 of the task's `nrows` plane rows.  The FIR runs horizontally
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

template <int kHW, int kHH, int FMT, int PC, int NR_MAX>
__device__ __forceinline__ void fir_produce_group_smem(const FusedFirParams& f, const float* sc, int b, int group, int Y0, int X0,
                                                       const uint8_t* stage, int g_in_stage, int m_h, int r_h, int nrows,
                                                       uint8_t* dst8, uint32_t lo_off, int lane) {
  constexpr int kCols = kHW / 2 + 1;
  constexpr int RS = 32 / kCols;
  static_assert(PC >= kCols + 2, "plane window columns");
  const int c = lane % kCols, rsub = lane / kCols;
  if (rsub >= RS) return;                                     // spare lanes: nothing below is warp-collective
  constexpr int kMaxCnt = (NR_MAX + RS - 1) / RS;             // plane rows per lane: the task's nrows (<= NR_MAX) over RS sub-groups
  const int base = nrows / RS, rem = nrows - base * RS;
  const int row_off = rsub * base + min(rsub, rem);
  const int cnt = base + (rsub < rem ? 1 : 0);
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
  float2 nzv[2 * kMaxCnt];
#pragma unroll
  for (int q = 0; q < 2 * kMaxCnt; ++q) {
    const int Y = 2 * m + q;
    nzv[q] = make_float2(0.f, 0.f);
    if (nzx_ok && q < 2 * cnt && Y >= 0 && Y < Ho) nzv[q] = __ldg(reinterpret_cast<const float2*>(nz_base + static_cast<size_t>(Y) * Wo + Xa));
  }
  FH ho_prev = hrow(true, r - 1), he_cur = hrow(false, r), ho_cur = hrow(true, r);
#pragma unroll
  for (int it = 0; it < kMaxCnt; ++it, ++m, ++r) {
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

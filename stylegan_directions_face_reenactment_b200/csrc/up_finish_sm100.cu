// Second half of an upsampling StyledConv in scatter mode (HBM-bound): the 4x4 FIR of Blur/upfirdn2d
// (libs/gan/StyleGAN2/model.py:72-88,256-257; op/upfirdn2d_kernel.cu:52-137) applied to the four parity planes the
// tcgen05 kernel left in HBM, fused with demodulation, NoiseInjection, FusedLeakyReLU (model.py:282-287,331-337) and the
// next layer's style multiply + hi/lo operand split.
//
//   T[u][v], u = 2I+pu, v = 2J+pv  (the (2H+1)^2 output of conv_transpose2d; plane (pu,pv), index (I,J) on (H+1)x(W+1))
//   z[Y][X] = sum_{a,b} fir[3-a][3-b] * T[Y+a-1][X+b-1]          upfirdn2d(pad=(1,1)): true convolution, zero outside
//   t = z * demod[b,o] + noise_w * noise[Y,X] + bias[o];  t = max(t, 0.2 t)
//   out_f32 = t * act_gain;   out_c8 = split(t * s2[b,o])
// Planes above index H/W are stored by the GEMM as exact zeros (their inputs are TMA out-of-bounds zeros), so only the
// low side (u, v < 0) needs a predicate.
//
#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

struct UpFinishParams {
  int B, C, H, W;              // input resolution; output is 2H x 2W; planes are (H+1) x (W+1)
  const float* t;              // [B][4 (oe,ee,eo,oo)][C/4][H+1][W+1][4]
  const float* fir;            // [4][4] blur.kernel
  const float* demod;          // [B,C]
  float plane_scale[4];        // undoes operand scales (+ accumulate-truncation compensation) per plane
  const float* bias;           // [C] or NULL
  const float* noise;          // [2H,2W] (+ batch stride) or NULL
  long long noise_bstride;
  const float* noise_w;
  const float* s2;             // [B,C] or NULL
  int act;
  float act_gain;
  int out_fmt;
  float out_scale;
  __nv_bfloat16* out_c8;       // [2][B][C/8][2H][2W][8] or NULL
  float* out_f32;              // [B,C,2H,2W] or NULL
  int single_out;              // single-pass consumers: the lo plane is not written
  int strips;                  // column strips of 14 output-producing columns
  int rows;                    // input rows walked by one warp
  int nslices;                 // split-K scatter GEMM: `t` holds this many copies of the plane tensor (one per K slice), added
  long long slice_stride;      // in slice order on load; floats between copies
};

// Work decomposition (instruction-issue bound otherwise: the generic 16-tap form costs ~37 instructions per output):
//   * the FIR is applied separably, fir = gy (x) gx (every blur kernel made by make_kernel from 1-D taps is rank 1,
//     model.py:19-27; the host mirror falls back to the polyphase packing for anything else): a vertical 4-tap pass over
//     the parity planes of one column, then a horizontal 4-tap pass over the column results of the lane neighbours;
//   * a thread owns column n and 4 channels and walks down `rows` input rows with a 3-row register window, so every
//     plane element is loaded once per thread column (one coalesced 16 B load per lane and plane per row);
//   * lanes l and l+16 hold the two 4-channel halves of column n0-1+l: a warp covers 16 columns of which the inner 14
//     produce output, the two outer ones only feed their neighbours' horizontal taps through warp shuffles;
//   * packed fp32x2 FMAs (FFMA2) halve the arithmetic issue slots.
struct F4 {
  float2 a, b;
};
__device__ __forceinline__ F4 f4_zero() { return F4{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}; }
__device__ __forceinline__ F4 f4_load(const float* p, bool ok) {
  if (!ok) return f4_zero();
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return F4{make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ F4 f4_mul(float w, const F4& x) {
  const float2 ww = make_float2(w, w);
  return F4{__fmul2_rn(ww, x.a), __fmul2_rn(ww, x.b)};
}
__device__ __forceinline__ F4 f4_fma(float w, const F4& x, const F4& acc) {
  const float2 ww = make_float2(w, w);
  return F4{__ffma2_rn(ww, x.a, acc.a), __ffma2_rn(ww, x.b, acc.b)};
}
__device__ __forceinline__ F4 f4_shfl(const F4& x, int src_lane) {
  F4 r;
  r.a.x = __shfl_sync(0xffffffffu, x.a.x, src_lane);
  r.a.y = __shfl_sync(0xffffffffu, x.a.y, src_lane);
  r.b.x = __shfl_sync(0xffffffffu, x.b.x, src_lane);
  r.b.y = __shfl_sync(0xffffffffu, x.b.y, src_lane);
  return r;
}

// bf16 hi/lo split of two values with paired conversions: hi = cvt.rn.bf16x2(v), lo = cvt.rn.bf16x2(v - float(hi))
__device__ __forceinline__ void split_pair_bf16(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - h0, v1 - h1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ROWS: input rows walked by one warp (H % ROWS == 0, so the loop is uniform and the shuffles need no re-convergence code)
template <int ROWS, int FMT, bool F32OUT, int NSL = 1>      // NSL: plane copies of a split-K scatter GEMM (1, 2, 4, 8)
__global__ void __launch_bounds__(128, NSL > 1 ? 2 : 4) up_finish_kernel(const UpFinishParams p) {
  // sc[0..3] = gy (vertical taps, flipped), sc[4..7] = gx (horizontal taps, flipped, normalised by the tap sum)
  __shared__ float sc[8];
  pdl_launch_dependents();
  pdl_wait();                                        // the parity planes come from the scatter GEMM right before
  if (threadIdx.x < 4) {
    const int a = threadIdx.x;
    float rs = 0.f, cs = 0.f, tot = 0.f;
    for (int i = 0; i < 4; ++i) {
      rs += __ldg(p.fir + (3 - a) * 4 + i);
      cs += __ldg(p.fir + i * 4 + (3 - a));
      for (int j = 0; j < 4; ++j) tot += __ldg(p.fir + i * 4 + j);
    }
    sc[a] = rs;
    sc[4 + a] = cs / tot;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l = lane & 15;              // lanes 0-15: channels 0-3 of the chunk, 16-31: channels 4-7
  const int strip = blockIdx.x % p.strips;
  const int rgroup = (blockIdx.x / p.strips) * (blockDim.x >> 5) + warp;      // exact grid: always < H / ROWS
  const int m0 = rgroup * ROWS;
  const int n = strip * 14 - 1 + l;
  const bool col_ok = n >= 0 && n <= p.W;                  // plane columns 0..W exist
  const bool out_ok = l >= 1 && l <= 14 && n < p.W;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int Hp = p.H + 1, Wp = p.W + 1;
  const int chunks = p.C >> 3;
  const size_t plane_stride = static_cast<size_t>(p.C >> 2) * Hp * Wp * 4;
  const size_t row_stride = static_cast<size_t>(Wp) * 4;
  // planes in [oe, ee, eo, oo] order; 4-channel group (chunk*2 + half); pointers sit on row m0
  const float* t_oe = p.t + ((static_cast<size_t>(b) * 4 * (p.C >> 2)) + chunk * 2 + half) * Hp * Wp * 4 +
                      static_cast<size_t>(col_ok ? n : 0) * 4 + static_cast<size_t>(m0) * row_stride;
  const float* t_ee = t_oe + plane_stride;
  const float* t_eo = t_ee + plane_stride;
  const float* t_oo = t_eo + plane_stride;

  // vertical coefficients with the per-plane scale folded in
  const float gx0 = sc[4], gx1 = sc[5], gx2 = sc[6], gx3 = sc[7];
  const float s_oe = p.plane_scale[0], s_ee = p.plane_scale[1], s_eo = p.plane_scale[2], s_oo = p.plane_scale[3];
  const float e0o = sc[0] * s_oe, e1e = sc[1] * s_ee, e2o = sc[2] * s_oe, e3e = sc[3] * s_ee;   // even column, py = 0
  const float e0e = sc[0] * s_ee, e1o = sc[1] * s_oe, e2e = sc[2] * s_ee, e3o = sc[3] * s_oe;   // even column, py = 1
  const float o0o = sc[0] * s_oo, o1e = sc[1] * s_eo, o2o = sc[2] * s_oo, o3e = sc[3] * s_eo;   // odd column, py = 0
  const float o0e = sc[0] * s_eo, o1o = sc[1] * s_oo, o2e = sc[2] * s_eo, o3o = sc[3] * s_oo;   // odd column, py = 1

  const int c0 = chunk * 8 + half * 4;
  const int Ho = 2 * p.H, Wo = 2 * p.W;
  float2 d01, d23, b01, b23, g01, g23;
  {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(p.demod + static_cast<size_t>(b) * p.C + c0));
    d01 = make_float2(d0.x, d0.y); d23 = make_float2(d0.z, d0.w);
    float bi[4], g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bi[e] = p.bias ? __ldg(p.bias + c0 + e) : 0.f;
      g[e] = (p.s2 ? __ldg(p.s2 + static_cast<size_t>(b) * p.C + c0 + e) : p.act_gain) * p.out_scale;
    }
    b01 = make_float2(bi[0], bi[1]); b23 = make_float2(bi[2], bi[3]);
    g01 = make_float2(g[0], g[1]); g23 = make_float2(g[2], g[3]);
  }
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  const float slope = p.act ? 0.2f : 1.f;
  const float2 slope2 = make_float2(slope, slope);
  const size_t out_plane = static_cast<size_t>(p.B) * p.C * Ho * Wo;       // elements per hi/lo plane
  // output pointers on row 2*m0: C8 pixel (2n + half) [lane `half` stores output pixel px == half], f32 row pair
  __nv_bfloat16* o_c8 = p.out_c8 ? p.out_c8 + (((static_cast<size_t>(b) * chunks + chunk) * Ho + 2 * m0) * Wo + 2 * (out_ok ? n : 0) + half) * 8
                                 : nullptr;
  float* o_f32 = F32OUT ? p.out_f32 + ((static_cast<size_t>(b) * p.C + c0) * Ho + 2 * m0) * Wo + 2 * (out_ok ? n : 0) : nullptr;
  const float* nz_ptr = p.noise ? p.noise + static_cast<size_t>(b) * p.noise_bstride + static_cast<size_t>(2 * m0) * Wo + 2 * (out_ok ? n : 0)
                                : nullptr;

  // split-K scatter GEMM (NSL > 1; small batches / small grids): every plane element is the sum of NSL partial planes, added in
  // slice order; all copies are requested before the first add, so a load costs one memory latency, not one per slice
  auto ld = [&](const float* q, bool ok) -> F4 {
    F4 r = f4_load(q, ok);
    if (NSL > 1) {
      F4 t[NSL > 1 ? NSL - 1 : 1];
#pragma unroll
      for (int sl = 1; sl < NSL; ++sl) t[sl - 1] = f4_load(q + sl * p.slice_stride, ok);
#pragma unroll
      for (int sl = 1; sl < NSL; ++sl) {
        r.a = __fadd2_rn(r.a, t[sl - 1].a);
        r.b = __fadd2_rn(r.b, t[sl - 1].b);
      }
    }
    return r;
  };
  // register window: odd planes at rows m-1, m; even planes at row m
  F4 oe_m1 = ld(t_oe - row_stride, col_ok && m0 > 0);
  F4 oo_m1 = ld(t_oo - row_stride, col_ok && m0 > 0);
  F4 ee_0 = ld(t_ee, col_ok);
  F4 eo_0 = ld(t_eo, col_ok);
  F4 oe_0 = ld(t_oe, col_ok);
  F4 oo_0 = ld(t_oo, col_ok);

  // software pipeline: the plane rows (and noise) of the NEXT iteration are requested before this iteration's math, so
  // the HBM latency overlaps ~400 instructions of work instead of stalling the first FMA (61 % long-scoreboard stalls
  // without it)
  t_oe += row_stride; t_ee += row_stride; t_eo += row_stride; t_oo += row_stride;        // row m0+1 <= H always exists
  F4 ee_n = ld(t_ee, col_ok), eo_n = ld(t_eo, col_ok), oe_n = ld(t_oe, col_ok), oo_n = ld(t_oo, col_ok);
  float2 nz_n[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  if (nz_ptr) {
    nz_n[0] = __ldg(reinterpret_cast<const float2*>(nz_ptr));
    nz_n[1] = __ldg(reinterpret_cast<const float2*>(nz_ptr + Wo));
  }
#pragma unroll 1
  for (int i = 0; i < ROWS; ++i) {
    const F4 ee_1 = ee_n, eo_1 = eo_n, oe_1 = oe_n, oo_1 = oo_n;
    const float2 nz_c[2] = {nz_n[0], nz_n[1]};
    if (i + 1 < ROWS) {
      t_oe += row_stride; t_ee += row_stride; t_eo += row_stride; t_oo += row_stride;    // row m+2 <= H inside the strip
      ee_n = ld(t_ee, col_ok); eo_n = ld(t_eo, col_ok); oe_n = ld(t_oe, col_ok); oo_n = ld(t_oo, col_ok);
      if (nz_ptr) {
        nz_ptr += 2 * Wo;
        nz_n[0] = __ldg(reinterpret_cast<const float2*>(nz_ptr));
        nz_n[1] = __ldg(reinterpret_cast<const float2*>(nz_ptr + Wo));
      }
    }
    // vertical pass: T rows 2m-1 .. 2m+3 of the even (2n) and odd (2n+1) column
    F4 ve[2], vo[2];
    ve[0] = f4_fma(e3e, ee_1, f4_fma(e2o, oe_0, f4_fma(e1e, ee_0, f4_mul(e0o, oe_m1))));
    ve[1] = f4_fma(e3o, oe_1, f4_fma(e2e, ee_1, f4_fma(e1o, oe_0, f4_mul(e0e, ee_0))));
    vo[0] = f4_fma(o3e, eo_1, f4_fma(o2o, oo_0, f4_fma(o1e, eo_0, f4_mul(o0o, oo_m1))));
    vo[1] = f4_fma(o3o, oo_1, f4_fma(o2e, eo_1, f4_fma(o1o, oo_0, f4_mul(o0e, eo_0))));
    oe_m1 = oe_0; oo_m1 = oo_0;
    ee_0 = ee_1; eo_0 = eo_1; oe_0 = oe_1; oo_0 = oo_1;

#pragma unroll
    for (int py = 0; py < 2; ++py) {
      // horizontal pass: columns 2n-1 .. 2n+2 (lane - 1 = column n-1, lane + 1 = column n+1; edge lanes are halo only)
      const F4 vo_l = f4_shfl(vo[py], lane - 1);
      const F4 ve_r = f4_shfl(ve[py], lane + 1);
      const F4 vo_r = f4_shfl(vo[py], lane + 1);
      const F4 z0 = f4_fma(gx3, ve_r, f4_fma(gx2, vo[py], f4_fma(gx1, ve[py], f4_mul(gx0, vo_l))));
      const F4 z1 = f4_fma(gx3, vo_r, f4_fma(gx2, ve_r, f4_fma(gx1, vo[py], f4_mul(gx0, ve[py]))));
      float2 nz0 = b01, nz1 = b01, nz2 = b23, nz3 = b23;          // (px0: ch01, px1: ch01, px0: ch23, px1: ch23)
      if (nz_ptr) {
        const float2 nn = nz_c[py];
        const float2 n0 = make_float2(nw * nn.x, nw * nn.x), n1 = make_float2(nw * nn.y, nw * nn.y);
        nz0 = __fadd2_rn(b01, n0); nz1 = __fadd2_rn(b01, n1); nz2 = __fadd2_rn(b23, n0); nz3 = __fadd2_rn(b23, n1);
      }
      // t = z * demod + noise + bias ; leaky relu
      float2 t0a = __ffma2_rn(z0.a, d01, nz0), t0b = __ffma2_rn(z0.b, d23, nz2);      // px 0: channels 01, 23
      float2 t1a = __ffma2_rn(z1.a, d01, nz1), t1b = __ffma2_rn(z1.b, d23, nz3);      // px 1
      {
        const float2 s0a = __fmul2_rn(t0a, slope2), s0b = __fmul2_rn(t0b, slope2);
        const float2 s1a = __fmul2_rn(t1a, slope2), s1b = __fmul2_rn(t1b, slope2);
        t0a = make_float2(fmaxf(t0a.x, s0a.x), fmaxf(t0a.y, s0a.y)); t0b = make_float2(fmaxf(t0b.x, s0b.x), fmaxf(t0b.y, s0b.y));
        t1a = make_float2(fmaxf(t1a.x, s1a.x), fmaxf(t1a.y, s1a.y)); t1b = make_float2(fmaxf(t1b.x, s1b.x), fmaxf(t1b.y, s1b.y));
      }
      if (F32OUT) {
        if (out_ok) {
          const size_t cs = static_cast<size_t>(Ho) * Wo;
          const float ag = p.act_gain;
          *reinterpret_cast<float2*>(o_f32) = make_float2(t0a.x * ag, t1a.x * ag);
          *reinterpret_cast<float2*>(o_f32 + cs) = make_float2(t0a.y * ag, t1a.y * ag);
          *reinterpret_cast<float2*>(o_f32 + 2 * cs) = make_float2(t0b.x * ag, t1b.x * ag);
          *reinterpret_cast<float2*>(o_f32 + 3 * cs) = make_float2(t0b.y * ag, t1b.y * ag);
        }
        o_f32 += Wo;
      }
      if (o_c8) {
        const float2 u0a = __fmul2_rn(t0a, g01), u0b = __fmul2_rn(t0b, g23), u1a = __fmul2_rn(t1a, g01), u1b = __fmul2_rn(t1b, g23);
        uint32_t hi[2][2], lo[2][2];
        if (FMT == kFmtBF16) {
          split_pair_bf16(u0a.x, u0a.y, hi[0][0], lo[0][0]);
          split_pair_bf16(u0b.x, u0b.y, hi[0][1], lo[0][1]);
          split_pair_bf16(u1a.x, u1a.y, hi[1][0], lo[1][0]);
          split_pair_bf16(u1b.x, u1b.y, hi[1][1], lo[1][1]);
        } else {
          split2(u0a.x, u0a.y, FMT, hi[0][0], lo[0][0]);
          split2(u0b.x, u0b.y, FMT, hi[0][1], lo[0][1]);
          split2(u1a.x, u1a.y, FMT, hi[1][0], lo[1][0]);
          split2(u1b.x, u1b.y, FMT, hi[1][1], lo[1][1]);
        }
        // lane `half` keeps output pixel px == half and receives the partner lane's 4 channels of that pixel
        uint32_t rh[2], rl[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          rh[e] = __shfl_xor_sync(0xffffffffu, half ? hi[0][e] : hi[1][e], 16);
          rl[e] = __shfl_xor_sync(0xffffffffu, half ? lo[0][e] : lo[1][e], 16);
        }
        if (out_ok) {
          const uint4 h4 = half ? make_uint4(rh[0], rh[1], hi[1][0], hi[1][1]) : make_uint4(hi[0][0], hi[0][1], rh[0], rh[1]);
          const uint4 l4 = half ? make_uint4(rl[0], rl[1], lo[1][0], lo[1][1]) : make_uint4(lo[0][0], lo[0][1], rl[0], rl[1]);
          *reinterpret_cast<uint4*>(o_c8) = h4;
          if (!p.single_out) *reinterpret_cast<uint4*>(o_c8 + out_plane) = l4;
        }
        o_c8 += static_cast<size_t>(Wo) * 8;
      }
    }
  }
}

template <int ROWS, int NSL>
static void up_finish_dispatch_nsl(const UpFinishParams& p, dim3 grid, int threads, cudaStream_t st) {
  const bool f32 = p.out_f32 != nullptr;
  if (p.out_fmt == kFmtBF16) {
    if (f32) launch_pdl(up_finish_kernel<ROWS, kFmtBF16, true, NSL>, grid, dim3(threads), 0, st, p);
    else launch_pdl(up_finish_kernel<ROWS, kFmtBF16, false, NSL>, grid, dim3(threads), 0, st, p);
  } else {
    if (f32) launch_pdl(up_finish_kernel<ROWS, kFmtFP16, true, NSL>, grid, dim3(threads), 0, st, p);
    else launch_pdl(up_finish_kernel<ROWS, kFmtFP16, false, NSL>, grid, dim3(threads), 0, st, p);
  }
}

template <int ROWS>
static void up_finish_dispatch(const UpFinishParams& p, dim3 grid, int threads, cudaStream_t st) {
  switch (p.nslices) {
    case 8: up_finish_dispatch_nsl<ROWS, 8>(p, grid, threads, st); break;
    case 4: up_finish_dispatch_nsl<ROWS, 4>(p, grid, threads, st); break;
    case 2: up_finish_dispatch_nsl<ROWS, 2>(p, grid, threads, st); break;
    default: up_finish_dispatch_nsl<ROWS, 1>(p, grid, threads, st); break;
  }
}

int up_finish_launch(const sgr_conv_args* a, float acc_scale, float comp_per_tap, cudaStream_t st, const float* planes,
                     int nslices) {
  UpFinishParams p;
  p.B = a->batch; p.C = a->cout; p.H = a->h_in; p.W = a->w_in;
  p.t = planes ? planes : a->t_scratch;
  if (nslices != 1 && nslices != 2 && nslices != 4 && nslices != 8) {
    set_error("up_finish: %d plane copies (1, 2, 4 or 8)", nslices);
    return 1;
  }
  p.nslices = nslices < 1 ? 1 : nslices;
  p.slice_stride = static_cast<long long>(a->batch) * 4 * a->cout * (a->h_in + 1) * (a->w_in + 1);
  p.fir = a->fir;
  p.demod = a->demod;
  const int taps[4] = {2, 4, 2, 1};                  // MMAs chains accumulated per plane: oe, ee, eo, oo
  for (int i = 0; i < 4; ++i) p.plane_scale[i] = acc_scale * (1.f + comp_per_tap * taps[i]);
  p.bias = a->bias;
  p.noise = a->noise;
  p.noise_bstride = a->noise_batch_stride;
  p.noise_w = a->noise_weight;
  p.s2 = a->s2;
  p.act = a->act;
  p.act_gain = a->act_gain;
  p.out_fmt = a->out_format;
  p.out_scale = act_scale(a->out_format);
  p.out_c8 = static_cast<__nv_bfloat16*>(a->out_c8);
  p.out_f32 = a->out_f32;
  p.single_out = a->single_pass ? 1 : 0;
  p.strips = (a->w_in + 13) / 14;
  // rows per warp: long strips amortise the two preloaded window rows, short ones keep small layers parallel
  int rows = a->h_in >= 64 ? 16 : (a->h_in >= 32 ? 8 : (a->h_in >= 8 ? 4 : 2));
  while (a->h_in % rows != 0) rows >>= 1;            // any height: fall back to shorter strips (1 always divides)
  // small batches: a warp walks its rows one after the other (a chain of `rows` + 2 memory latencies), so when the whole
  // launch has fewer warps than the machine wants (~8 per SM) shorter strips finish sooner (batch 1, 64 -> 128: 16 -> 8 us)
  {
    const int sms = num_sms() > 0 ? num_sms() : 148;
    const long long per_row_group = static_cast<long long>(p.strips) * (a->cout / 8) * a->batch;
    while (rows > 1 && per_row_group * (a->h_in / rows) < 8LL * sms) rows >>= 1;
  }
  p.rows = rows;
  const int rgroups = a->h_in / rows;
  int warps = 4;
  while (rgroups % warps != 0) warps >>= 1;
  dim3 grid(p.strips * (rgroups / warps), a->cout / 8, a->batch);
  switch (rows) {
    case 16: up_finish_dispatch<16>(p, grid, warps * 32, st); break;
    case 8: up_finish_dispatch<8>(p, grid, warps * 32, st); break;
    case 4: up_finish_dispatch<4>(p, grid, warps * 32, st); break;
    case 2: up_finish_dispatch<2>(p, grid, warps * 32, st); break;
    case 1: up_finish_dispatch<1>(p, grid, warps * 32, st); break;
    default: set_error("up_finish: unsupported input height %d", a->h_in); return 1;
  }
  count_launch();
  return check_launch("up_finish_kernel") ? 0 : 1;
}

}  // namespace sgr

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
  const float* t;              // [B][4 (oe,ee,eo,oo)][C/8][H+1][W+1][8]
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
  int strips;                  // column strips of 14 output-producing columns
  int rows;                    // input rows walked by one warp
};

// Work decomposition (instruction-issue bound otherwise: the generic 16-tap form costs ~37 instructions per output):
//   * the FIR is applied separably, fir = gy (x) gx (every blur kernel made by make_kernel from 1-D taps is rank 1,
//     model.py:19-27; the host mirror falls back to the polyphase packing for anything else): a vertical 4-tap pass over
//     the parity planes of one column, then a horizontal 4-tap pass over the column results of the lane neighbours;
//   * a thread owns column n and 4 channels and walks down `rows` input rows with a 3-row register window, so every
//     plane element is loaded once per thread column (one coalesced 16 B load per lane and plane per row);
//   * lanes (2l, 2l+1) hold the two 4-channel halves of column n0-1+l: a warp covers 16 columns of which the inner 14
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

__global__ void __launch_bounds__(128, 6) up_finish_kernel(const UpFinishParams p) {
  // sc[0..3] = gy (vertical taps, flipped), sc[4..7] = gx (horizontal taps, flipped, normalised by the tap sum)
  __shared__ float sc[8];
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
  const int half = lane & 1, l = lane >> 1;
  const int strip = blockIdx.x % p.strips;
  const int rgroup = (blockIdx.x / p.strips) * (blockDim.x >> 5) + warp;
  const int m0 = rgroup * p.rows;
  if (m0 >= p.H) return;                                   // whole warp
  const int m1 = min(m0 + p.rows, p.H);
  const int n = strip * 14 - 1 + l;
  const bool col_ok = n >= 0 && n <= p.W;                  // plane columns 0..W exist
  const bool out_ok = l >= 1 && l <= 14 && n < p.W;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int Hp = p.H + 1, Wp = p.W + 1;
  const int chunks = p.C >> 3;
  const size_t plane_stride = static_cast<size_t>(chunks) * Hp * Wp * 8;
  // planes in [oe, ee, eo, oo] order
  const float* t_oe = p.t + (static_cast<size_t>(b) * 4 * chunks + chunk) * Hp * Wp * 8 + half * 4 + static_cast<size_t>(col_ok ? n : 0) * 8;
  const float* t_ee = t_oe + plane_stride;
  const float* t_eo = t_ee + plane_stride;
  const float* t_oo = t_eo + plane_stride;
  const size_t row_stride = static_cast<size_t>(Wp) * 8;

  // vertical coefficients with the per-plane scale folded in
  const float gy0 = sc[0], gy1 = sc[1], gy2 = sc[2], gy3 = sc[3];
  const float gx0 = sc[4], gx1 = sc[5], gx2 = sc[6], gx3 = sc[7];
  const float s_oe = p.plane_scale[0], s_ee = p.plane_scale[1], s_eo = p.plane_scale[2], s_oo = p.plane_scale[3];

  const int c0 = chunk * 8 + half * 4;
  const int Ho = 2 * p.H, Wo = 2 * p.W;
  float d[4], bi[4], g[4];
  {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(p.demod + static_cast<size_t>(b) * p.C + c0));
    d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bi[e] = p.bias ? __ldg(p.bias + c0 + e) : 0.f;
      g[e] = (p.s2 ? __ldg(p.s2 + static_cast<size_t>(b) * p.C + c0 + e) : p.act_gain) * p.out_scale;
    }
  }
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  const size_t out_plane = static_cast<size_t>(p.B) * p.C * Ho * Wo;       // elements per hi/lo plane

  // register window: odd planes at rows m-1, m; even planes at row m
  F4 oe_m1 = f4_load(t_oe + static_cast<size_t>(m0 - 1) * row_stride, col_ok && m0 > 0);
  F4 oo_m1 = f4_load(t_oo + static_cast<size_t>(m0 - 1) * row_stride, col_ok && m0 > 0);
  F4 ee_0 = f4_load(t_ee + static_cast<size_t>(m0) * row_stride, col_ok);
  F4 eo_0 = f4_load(t_eo + static_cast<size_t>(m0) * row_stride, col_ok);
  F4 oe_0 = f4_load(t_oe + static_cast<size_t>(m0) * row_stride, col_ok);
  F4 oo_0 = f4_load(t_oo + static_cast<size_t>(m0) * row_stride, col_ok);

  for (int m = m0; m < m1; ++m) {
    const size_t r1 = static_cast<size_t>(m + 1) * row_stride;             // row m+1 <= H always exists
    const F4 ee_1 = f4_load(t_ee + r1, col_ok);
    const F4 eo_1 = f4_load(t_eo + r1, col_ok);
    const F4 oe_1 = f4_load(t_oe + r1, col_ok);
    const F4 oo_1 = f4_load(t_oo + r1, col_ok);
    // vertical pass: T rows 2m-1 .. 2m+3 of the even (2n) and odd (2n+1) column
    F4 ve[2], vo[2];
    ve[0] = f4_fma(gy3 * s_ee, ee_1, f4_fma(gy2 * s_oe, oe_0, f4_fma(gy1 * s_ee, ee_0, f4_mul(gy0 * s_oe, oe_m1))));
    ve[1] = f4_fma(gy3 * s_oe, oe_1, f4_fma(gy2 * s_ee, ee_1, f4_fma(gy1 * s_oe, oe_0, f4_mul(gy0 * s_ee, ee_0))));
    vo[0] = f4_fma(gy3 * s_eo, eo_1, f4_fma(gy2 * s_oo, oo_0, f4_fma(gy1 * s_eo, eo_0, f4_mul(gy0 * s_oo, oo_m1))));
    vo[1] = f4_fma(gy3 * s_oo, oo_1, f4_fma(gy2 * s_eo, eo_1, f4_fma(gy1 * s_oo, oo_0, f4_mul(gy0 * s_eo, eo_0))));
    oe_m1 = oe_0; oo_m1 = oo_0;
    ee_0 = ee_1; eo_0 = eo_1; oe_0 = oe_1; oo_0 = oo_1;

#pragma unroll
    for (int py = 0; py < 2; ++py) {
      // horizontal pass: columns 2n-1 .. 2n+2 (lane - 2 = column n-1, lane + 2 = column n+1; edge lanes are halo only)
      const F4 vo_l = f4_shfl(vo[py], lane - 2);
      const F4 ve_r = f4_shfl(ve[py], lane + 2);
      const F4 vo_r = f4_shfl(vo[py], lane + 2);
      const F4 z0 = f4_fma(gx3, ve_r, f4_fma(gx2, vo[py], f4_fma(gx1, ve[py], f4_mul(gx0, vo_l))));
      const F4 z1 = f4_fma(gx3, vo_r, f4_fma(gx2, ve_r, f4_fma(gx1, vo[py], f4_mul(gx0, ve[py]))));
      const float z[2][4] = {{z0.a.x, z0.a.y, z0.b.x, z0.b.y}, {z1.a.x, z1.a.y, z1.b.x, z1.b.y}};
      const int oy = 2 * m + py;
      float nz[2] = {0.f, 0.f};
      if (p.noise && out_ok) {
        const float2 nn = __ldg(reinterpret_cast<const float2*>(p.noise + static_cast<size_t>(b) * p.noise_bstride +
                                                                 static_cast<size_t>(oy) * Wo + 2 * n));
        nz[0] = nw * nn.x;
        nz[1] = nw * nn.y;
      }
      float t[2][4];
#pragma unroll
      for (int px = 0; px < 2; ++px)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = fmaf(z[px][e], d[e], nz[px] + bi[e]);
          if (p.act) v = fmaxf(v, 0.2f * v);
          t[px][e] = v;
        }
      if (p.out_f32 && out_ok) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2* dst = reinterpret_cast<float2*>(p.out_f32 + ((static_cast<size_t>(b) * p.C + c0 + e) * Ho + oy) * Wo + 2 * n);
          *dst = make_float2(t[0][e] * p.act_gain, t[1][e] * p.act_gain);
        }
      }
      if (p.out_c8) {
        uint32_t hi[2][2], lo[2][2];
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            split2(t[px][2 * e] * g[2 * e], t[px][2 * e + 1] * g[2 * e + 1], p.out_fmt, hi[px][e], lo[px][e]);
        // lane `half` keeps output pixel px == half and receives the partner lane's 4 channels of that pixel
        uint32_t rh[2], rl[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          rh[e] = __shfl_xor_sync(0xffffffffu, half ? hi[0][e] : hi[1][e], 1);
          rl[e] = __shfl_xor_sync(0xffffffffu, half ? lo[0][e] : lo[1][e], 1);
        }
        if (out_ok) {
          const size_t off = (((static_cast<size_t>(b) * chunks + chunk) * Ho + oy) * Wo + 2 * n + half) * 8;
          const uint4 h4 = half ? make_uint4(rh[0], rh[1], hi[1][0], hi[1][1]) : make_uint4(hi[0][0], hi[0][1], rh[0], rh[1]);
          const uint4 l4 = half ? make_uint4(rl[0], rl[1], lo[1][0], lo[1][1]) : make_uint4(lo[0][0], lo[0][1], rl[0], rl[1]);
          *reinterpret_cast<uint4*>(p.out_c8 + off) = h4;
          *reinterpret_cast<uint4*>(p.out_c8 + out_plane + off) = l4;
        }
      }
    }
  }
}

int up_finish_launch(const sgr_conv_args* a, float acc_scale, float comp_per_tap, cudaStream_t st) {
  UpFinishParams p;
  p.B = a->batch; p.C = a->cout; p.H = a->h_in; p.W = a->w_in;
  p.t = a->t_scratch;
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
  p.strips = (a->w_in + 13) / 14;
  p.rows = a->h_in >= 64 ? 16 : (a->h_in >= 16 ? 8 : a->h_in);
  const int rgroups = (a->h_in + p.rows - 1) / p.rows;
  const int warps = rgroups < 4 ? rgroups : 4;
  dim3 grid(p.strips * ((rgroups + warps - 1) / warps), a->cout / 8, a->batch);
  up_finish_kernel<<<grid, warps * 32, 0, st>>>(p);
  count_launch();
  return check_launch("up_finish_kernel") ? 0 : 1;
}

}  // namespace sgr

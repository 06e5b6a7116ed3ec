// Fused epilogue math shared by the tcgen05 convolution kernels (modconv_sm100.cu, modconv_halo_sm100.cu):
// demodulation, NoiseInjection, FusedLeakyReLU (model.py:282-287,331-337; op/fused_bias_act_kernel.cu:18-49), the next
// layer's style multiply + hi/lo operand split, and the partial sums of the following ToRGB (model.py:350-354).
#pragma once
#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

// v: 32 consecutive accumulator columns n0 .. n0+31 of pixel (b, y, x) of the GEMM grid.
__device__ __forceinline__ void epilogue_32cols(const ConvKernelParams& p, float (&v)[32], int n0, int b, int y, int x,
                                                float nw, size_t plane_stride, float& rgb0, float& rgb1, float& rgb2) {
        const int phase = p.up ? n0 / p.cout : 0;
    const int o0 = n0 & (p.cout - 1);
    const int oy = p.up ? 2 * y + (phase >> 1) : y;
    const int ox = p.up ? 2 * x + (phase & 1) : x;
    const float nz = p.noise ? nw * __ldg(p.noise + static_cast<size_t>(b) * p.noise_bstride + oy * p.Wout + ox) : 0.f;
    const float* dptr = p.demod ? p.demod + static_cast<size_t>(b) * p.cout + o0 : nullptr;
    const float* bptr = p.bias ? p.bias + o0 : nullptr;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 d4 = dptr ? __ldg(reinterpret_cast<const float4*>(dptr) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
      d4.x *= p.acc_scale; d4.y *= p.acc_scale; d4.z *= p.acc_scale; d4.w *= p.acc_scale;
      float4 b4 = bptr ? __ldg(reinterpret_cast<const float4*>(bptr) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      float t0 = fmaf(v[4 * q + 0], d4.x, nz + b4.x);
      float t1 = fmaf(v[4 * q + 1], d4.y, nz + b4.y);
      float t2 = fmaf(v[4 * q + 2], d4.z, nz + b4.z);
      float t3 = fmaf(v[4 * q + 3], d4.w, nz + b4.w);
      if (p.act) {
        t0 = fmaxf(t0, 0.2f * t0);
        t1 = fmaxf(t1, 0.2f * t1);
        t2 = fmaxf(t2, 0.2f * t2);
        t3 = fmaxf(t3, 0.2f * t3);
      }
      v[4 * q + 0] = t0;
      v[4 * q + 1] = t1;
      v[4 * q + 2] = t2;
      v[4 * q + 3] = t3;
    }
    if (p.rgb_coef) {
      const float* cptr = p.rgb_coef + static_cast<size_t>(b) * 3 * p.cout + o0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(cptr) + q);
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(cptr + p.cout) + q);
        const float4 c2 = __ldg(reinterpret_cast<const float4*>(cptr + 2 * p.cout) + q);
        rgb0 = fmaf(v[4 * q], c0.x, fmaf(v[4 * q + 1], c0.y, fmaf(v[4 * q + 2], c0.z, fmaf(v[4 * q + 3], c0.w, rgb0))));
        rgb1 = fmaf(v[4 * q], c1.x, fmaf(v[4 * q + 1], c1.y, fmaf(v[4 * q + 2], c1.z, fmaf(v[4 * q + 3], c1.w, rgb1))));
        rgb2 = fmaf(v[4 * q], c2.x, fmaf(v[4 * q + 1], c2.y, fmaf(v[4 * q + 2], c2.z, fmaf(v[4 * q + 3], c2.w, rgb2))));
      }
    }
    if (p.out_f32) {
      float* optr = p.out_f32 + ((static_cast<size_t>(b) * p.cout + o0) * p.Hout + oy) * p.Wout + ox;
      const size_t cs = static_cast<size_t>(p.Hout) * p.Wout;
#pragma unroll
      for (int j = 0; j < 32; ++j) optr[j * cs] = v[j] * p.act_gain;
    }
    if (p.out_c8) {
      const float* sptr = p.s2 ? p.s2 + static_cast<size_t>(b) * p.cout + o0 : nullptr;
      // element offset of channel chunk (o0/8) of this pixel inside the hi plane
      __nv_bfloat16* optr = p.out_c8 +
          (((static_cast<size_t>(b) * (p.cout >> 3) + (o0 >> 3)) * p.Hout + oy) * p.Wout + ox) * 8;
      const size_t chunk_stride = static_cast<size_t>(p.Hout) * p.Wout * 8;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float g[8];
        if (sptr) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(sptr) + 2 * q);
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(sptr) + 2 * q + 1);
          g[0] = s0.x; g[1] = s0.y; g[2] = s0.z; g[3] = s0.w;
          g[4] = s1.x; g[5] = s1.y; g[6] = s1.z; g[7] = s1.w;
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] *= p.out_scale;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = p.act_gain * p.out_scale;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          split2(v[8 * q + 2 * e] * g[2 * e], v[8 * q + 2 * e + 1] * g[2 * e + 1], p.out_fmt, hi[e], lo[e]);
        *reinterpret_cast<uint4*>(optr + q * chunk_stride) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (!p.single_out) *reinterpret_cast<uint4*>(optr + plane_stride + q * chunk_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
}

// one partial-sum slot per column tile (summed in a fixed order by torgb_tail_kernel: deterministic)
__device__ __forceinline__ void rgb_store(const ConvKernelParams& p, int n_tile, int b, int y, int x, float rgb0,
                                          float rgb1, float rgb2) {
  const size_t cs = static_cast<size_t>(p.Hout) * p.Wout;
  float* rptr = p.rgb_part + ((static_cast<size_t>(n_tile) * p.B + b) * 3) * cs + static_cast<size_t>(y) * p.Wout + x;
  rptr[0] = rgb0;
  rptr[cs] = rgb1;
  rptr[2 * cs] = rgb2;
}

}  // namespace sgr

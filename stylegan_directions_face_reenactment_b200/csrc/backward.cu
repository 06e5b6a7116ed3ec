// Backward of the synthesis network with respect to the W+ latent (the only gradient the A-matrix training of the
// reference consumes: Adam steps A alone, libs/trainer.py:144,187-189).
//
// Per styled layer l (input a_{l-1}, style s_l, demod d_l, pre-activation t_l, output a_l = sqrt2*lrelu(t_l)):
//   ga_l  = gx_{l+1} * s_{l+1}  +  sum_c grgb_r[c] * Wrgb[c,o]/sqrt(C) * s_rgb[b,o]        (next conv + ToRGB branches)
//   gt_l  = ga_l * sqrt2 * (a_l > 0 ? 1 : 0.2)                     (op/fused_bias_act_kernel.cu:43, keyed on the OUTPUT)
//   gz_l  = gt_l * d_l[b,o]                                          -> bf16 hi/lo C8 planes, operand of the dgrad GEMM
//   q_l   = sum_p gt_l * (t_l - noise - bias)   ( = gd * d )         demodulation branch
//   gx_l  = conv^T_l(gz_l)                                           tcgen05 kernel with the adjoint-packed shared weights
//   ds_l  = sum_p a_{l-1} * gx_l  -  s_l * ((q_l * d_l^2) @ Wsq_l)   conv term + demod term
//   dlatent[:, row_l] += ds_l @ Wm_l / sqrt(512)                     EqualLinear modulation backward
// No per-sample weight gradient is ever formed; generator weight gradients (optimize_g) are out of this kernel set.
#include <string.h>

#include <algorithm>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

struct BwdActParams {
  int B, C, H, W;
  const float* gx;        // [B,C,H,W] dgrad output of the NEXT layer, or NULL
  const float* s_next;    // [B,C] next layer's style (scale of gx)
  const float* a;         // [B,C,H,W] saved forward output of this layer (batch stride a_bstride)
  long long a_bstride;
  const float* grgb;      // [B,3,H,W] or NULL
  const float* wrgb;      // [3,C]
  const float* s_rgb;     // [B,C]
  const float* demod;     // [B,C]
  const float* bias;      // [C]
  const float* noise;     // [H,W] (+ batch stride)
  long long noise_bstride;
  const float* noise_w;
  __nv_bfloat16* out_c8;  // gz planes or NULL
  float* out_ga;          // dL/d(a) [B,C,H,W] fp32 (generator-parameter gradients) or NULL
  float* out_gz4;         // gz as fp32 [B][C/4][H][W][4] (scatter-form up layers: input of up_bwd_prepare_kernel) or NULL
  int s2d;
  int act;                // 0: constant-input pseudo layer (only the ds reduction)
  int iters;              // pixel iterations per thread (set by bwd_act_launch)
  float* ds_next;         // [B,C] += sum_p a * gx
  float* q;               // [B,C] += sum_p gt * (t - noise - bias)
  float* ds_rgb;          // [B,C] += sum_p a * sum_c grgb*wrgb/sqrt(C)
  // PARAMS variant (generator-parameter gradients, optimize_g): per-channel sums of the same operands, formed in the
  // same pass (they used to be a second kernel that re-read a and a stored fp32 copy of ga: 12 of 24 bytes per element)
  //   d(activate.bias)[o] = sum_{b,p} gt        d(noise.weight) = sum_{b,o,p} gt * noise[p]
  //   d(ToRGB.conv.weight)[c,o] = sum_{b,p} grgb[b,c,p] a[b,o,p] * s_rgb[b,o] / sqrt(C)      d(ToRGB.bias)[c] = sum_{b,p} grgb
  float* g_act_bias;      // [C]
  float* g_noise_w;       // [1]
  float* g_wrgb;          // [3,C] or NULL
  float* g_rgb_bias;      // [3]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// grid: (pixel spans, C/8, B); a thread handles 1 pixel x 8 channels per iteration and `iters` iterations
// (block stride), keeps the three per-(b,c) sums in registers, and the block issues ONE atomic per sum and channel at
// the end (one warp-reduction + atomic per iteration put 1.6 M atomics on 1024 addresses at 256^2: 4x off the HBM time).
constexpr int kBwdActThreads = 256;
template <bool PARAMS>
__global__ void __launch_bounds__(kBwdActThreads, PARAMS ? 2 : 3) bwd_act_kernel(const BwdActParams p) {
  constexpr int kSums = PARAMS ? 60 : 24;      // + 8 bias, 24 ToRGB weight, 3 ToRGB bias, 1 noise weight
  __shared__ float red[kBwdActThreads / 32][kSums];
  const int HW = p.H * p.W;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float kSqrt2 = 1.4142135623730951f, kInvSqrt2 = 0.7071067811865476f;
  const float rs = rsqrtf(static_cast<float>(p.C));
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  // per-channel constants live in shared memory (56 registers otherwise: 25 % occupancy on an HBM-latency-bound kernel)
  __shared__ float sn[8], w0[8], w1[8], w2[8], sr[8], dm[8], bi[8];
  if (threadIdx.x < 8) {
    const int e = threadIdx.x, c = c0 + e;
    sn[e] = p.gx ? __ldg(p.s_next + static_cast<size_t>(b) * p.C + c) : 0.f;
    w0[e] = w1[e] = w2[e] = sr[e] = 0.f;
    if (p.grgb) {
      w0[e] = __ldg(p.wrgb + c) * rs;
      w1[e] = __ldg(p.wrgb + p.C + c) * rs;
      w2[e] = __ldg(p.wrgb + 2 * p.C + c) * rs;
      sr[e] = __ldg(p.s_rgb + static_cast<size_t>(b) * p.C + c);
    }
    dm[e] = p.act ? __ldg(p.demod + static_cast<size_t>(b) * p.C + c) : 1.f;
    bi[e] = (p.act && p.bias) ? __ldg(p.bias + c) : 0.f;
  }
  __syncthreads();
  float dsn[8], qa[8], dsr[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) dsn[e] = qa[e] = dsr[e] = 0.f;
  float gb[PARAMS ? 8 : 1], gw3[PARAMS ? 24 : 1], gb3[PARAMS ? 3 : 1], gnw = 0.f;
  if (PARAMS) {
#pragma unroll
    for (int e = 0; e < 8; ++e) gb[e] = 0.f;
#pragma unroll
    for (int e = 0; e < 24; ++e) gw3[e] = 0.f;
    gb3[0] = gb3[1] = gb3[2] = 0.f;
  }

  // one pixel per thread and iteration: consecutive lanes read consecutive floats of every channel plane (128 B per warp
  // and channel) and write consecutive 16 B operand chunks (4 pixels per thread made every store instruction touch
  // 32 sectors for 512 B: 28 % excess sectors)
  for (int it = 0; it < p.iters; ++it) {
    const int pix = (blockIdx.x * p.iters + it) * kBwdActThreads + threadIdx.x;
    if (pix >= HW) break;
    float g3[3] = {0.f, 0.f, 0.f};
    if (p.grgb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) g3[c] = __ldg(p.grgb + (static_cast<size_t>(b) * 3 + c) * HW + pix);
    }
    const float nraw = p.noise ? __ldg(p.noise + static_cast<size_t>(b) * p.noise_bstride + pix) : 0.f;
    const float nz = nw * nraw;
    if (PARAMS && p.grgb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gb3[c] += g3[c];
    }
    float gtsum = 0.f;
    float av[8], gxv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      av[e] = __ldg(p.a + static_cast<size_t>(b) * p.a_bstride + static_cast<size_t>(c0 + e) * HW + pix);
      gxv[e] = p.gx ? __ldg(p.gx + (static_cast<size_t>(b) * p.C + c0 + e) * HW + pix) : 0.f;
    }
    float gzv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = av[e];
      float ga = gxv[e] * sn[e];
      dsn[e] = fmaf(a, gxv[e], dsn[e]);
      if (p.grgb) {
        const float r = fmaf(g3[0], w0[e], fmaf(g3[1], w1[e], g3[2] * w2[e]));
        ga = fmaf(r, sr[e], ga);
        dsr[e] = fmaf(a, r, dsr[e]);
        if (PARAMS) {
#pragma unroll
          for (int c = 0; c < 3; ++c) gw3[c * 8 + e] = fmaf(a, g3[c], gw3[c * 8 + e]);
        }
      }
      if (p.out_ga) p.out_ga[(static_cast<size_t>(b) * p.C + c0 + e) * HW + pix] = ga;
      gzv[e] = 0.f;
      if (p.act) {
        const bool pos = a > 0.f;
        const float gt = ga * (pos ? kSqrt2 : 0.2f * kSqrt2);
        const float t = pos ? a * kInvSqrt2 : a * (5.f * kInvSqrt2);
        qa[e] = fmaf(gt, t - nz - bi[e], qa[e]);
        gzv[e] = gt * dm[e];
        if (PARAMS) {
          gb[e] += gt;
          gtsum += gt;
        }
      }
    }
    if (PARAMS) gnw = fmaf(gtsum, nraw, gnw);
    if (p.out_gz4 && p.act) {
      float* dst = p.out_gz4 + ((static_cast<size_t>(b) * (p.C / 4) + blockIdx.y * 2) * HW + pix) * 4;
      *reinterpret_cast<float4*>(dst) = make_float4(gzv[0], gzv[1], gzv[2], gzv[3]);
      *reinterpret_cast<float4*>(dst + static_cast<size_t>(HW) * 4) = make_float4(gzv[4], gzv[5], gzv[6], gzv[7]);
    }
    if (p.out_c8 && p.act) {
      const int chunks = p.C / 8;
      size_t off;
      if (!p.s2d) {
        off = (static_cast<size_t>(b) * chunks + blockIdx.y) * HW + pix;
      } else {
        const int y = pix / p.W, x = pix - y * p.W;
        const int phase = (y & 1) * 2 + (x & 1);
        const int H2 = p.H / 2, W2 = p.W / 2;
        off = ((static_cast<size_t>(b) * (4 * chunks) + phase * chunks + blockIdx.y) * H2 + (y >> 1)) * W2 + (x >> 1);
      }
      const size_t plane = static_cast<size_t>(p.B) * chunks * HW;        // 16-byte rows per hi/lo plane
      uint32_t hp[4], lp[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split2(gzv[2 * e], gzv[2 * e + 1], kFmtBF16, hp[e], lp[e]);
      uint4* o4 = reinterpret_cast<uint4*>(p.out_c8);
      o4[off] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      o4[plane + off] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
  }

  // block reduction of the 24 sums, then one atomic each
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    dsn[e] = warp_sum(dsn[e]);
    qa[e] = warp_sum(qa[e]);
    dsr[e] = warp_sum(dsr[e]);
  }
  if (PARAMS) {
#pragma unroll
    for (int e = 0; e < 8; ++e) gb[e] = warp_sum(gb[e]);
    if (p.grgb) {
#pragma unroll
      for (int e = 0; e < 24; ++e) gw3[e] = warp_sum(gw3[e]);
#pragma unroll
      for (int c = 0; c < 3; ++c) gb3[c] = warp_sum(gb3[c]);
    }
    gnw = warp_sum(gnw);
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][e] = dsn[e];
      red[warp][8 + e] = qa[e];
      red[warp][16 + e] = dsr[e];
    }
    if (PARAMS) {
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][24 + e] = gb[e];
#pragma unroll
      for (int e = 0; e < 24; ++e) red[warp][32 + e] = gw3[e];
#pragma unroll
      for (int c = 0; c < 3; ++c) red[warp][56 + c] = gb3[c];
      red[warp][59] = gnw;
    }
  }
  __syncthreads();
  if (threadIdx.x < kSums) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdActThreads / 32; ++w) v += red[w][threadIdx.x];
    const int k = threadIdx.x;
    if (k < 24) {
      const int kind = k >> 3, e = k & 7;
      const size_t o = static_cast<size_t>(b) * p.C + c0 + e;
      if (kind == 0 && p.ds_next && p.gx) atomicAdd(p.ds_next + o, v);
      if (kind == 1 && p.act) atomicAdd(p.q + o, v);
      if (kind == 2 && p.grgb) atomicAdd(p.ds_rgb + o, v);
    } else if (PARAMS && p.act) {
      if (k < 32) {
        atomicAdd(p.g_act_bias + c0 + (k - 24), v);
      } else if (k < 56) {
        const int c = (k - 32) >> 3, e = (k - 32) & 7;
        if (p.grgb && p.g_wrgb) atomicAdd(p.g_wrgb + c * p.C + c0 + e, v * sr[e] * rs);
      } else if (k < 59) {
        if (p.grgb && p.g_rgb_bias && blockIdx.y == 0) atomicAdd(p.g_rgb_bias + (k - 56), v);
      } else if (p.noise) {
        atomicAdd(p.g_noise_w, v);
      }
    }
  }
}

static int bwd_act_launch(BwdActParams p, cudaStream_t st) {
  const int HW = p.H * p.W;
  const int groups = (HW + kBwdActThreads - 1) / kBwdActThreads;           // thread-iterations along the pixels
  // enough blocks to fill the machine, as few atomics as possible
  const long long others = static_cast<long long>(p.C / 8) * p.B;
  int iters = 1;
  while (iters < 16 && (groups / (iters * 2)) * others >= 2048) iters *= 2;
  p.iters = iters;
  dim3 grid((groups + iters - 1) / iters, p.C / 8, p.B);
  if (p.g_act_bias) bwd_act_kernel<true><<<grid, kBwdActThreads, 0, st>>>(p);
  else bwd_act_kernel<false><<<grid, kBwdActThreads, 0, st>>>(p);
  count_launch();
  return check_launch("bwd_act_kernel") ? 0 : 1;
}

// Scatter-form up layers, backward: G = FIR^T(gz) on the (2H+1)^2 grid of conv_transpose2d outputs, split into its four
// parity planes [ee|eo|oe|oo] on the (H+1) x (W+1) grid and written as the bf16 hi/lo operand (4*C channels, plane-major)
// of the 9-tap gather GEMM (sgr_conv_args.up == 3).  Adjoint of up_finish_kernel's FIR (upfirdn2d backward,
// op/upfirdn2d.py:112-117): G[u][v] = sum_{a,b} fir[3-a][3-b] gz[u-a+1][v-b+1].
struct UpBwdPrepParams {
  int B, C, H, W;            // H, W = INPUT resolution of the up layer; gz is [B][C/4][2H][2W][4]
  const float* gz;
  const float* fir;          // [4][4] blur.kernel
  __nv_bfloat16* planes;     // [2][B][4*C/8][H+1][W+1][8]
};

__global__ void __launch_bounds__(128) up_bwd_prepare_kernel(const UpBwdPrepParams p) {
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = __ldg(p.fir + (3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4));
  __syncthreads();
  const int Hp = p.H + 1, Wp = p.W + 1, Ho = 2 * p.H, Wo = 2 * p.W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= Hp * Wp) return;
  const int J = pix % Wp, I = pix / Wp;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const float* g0 = p.gz + (static_cast<size_t>(b) * (p.C / 4) + chunk * 2) * Ho * Wo * 4;
  const float* g1 = g0 + static_cast<size_t>(Ho) * Wo * 4;
  float acc[4][8];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[o][e] = 0.f;
#pragma unroll
  for (int dY = -2; dY <= 2; ++dY) {
    const int Y = 2 * I + dY;
    if (Y < 0 || Y >= Ho) continue;
#pragma unroll
    for (int dX = -2; dX <= 2; ++dX) {
      const int X = 2 * J + dX;
      if (X < 0 || X >= Wo) continue;
      const size_t off = (static_cast<size_t>(Y) * Wo + X) * 4;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(g0 + off));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(g1 + off));
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int pu = 0; pu < 2; ++pu) {
        const int a = pu - dY + 1;
        if (a < 0 || a > 3) continue;
#pragma unroll
        for (int pv = 0; pv < 2; ++pv) {
          const int bq = pv - dX + 1;
          if (bq < 0 || bq > 3) continue;
          const float w = sk[a * 4 + bq];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[pu * 2 + pv][e] = fmaf(w, v[e], acc[pu * 2 + pv][e]);
        }
      }
    }
  }
  const size_t plane_elems = static_cast<size_t>(p.B) * 4 * p.C * Hp * Wp;         // elements per hi/lo plane
  const int chunks4 = 4 * (p.C / 8);
#pragma unroll
  for (int pl = 0; pl < 4; ++pl) {
    const bool exists = (2 * I + (pl >> 1) <= 2 * p.H) && (2 * J + (pl & 1) <= 2 * p.W);   // inside the (2H+1)^2 grid
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      split2(exists ? acc[pl][2 * e] : 0.f, exists ? acc[pl][2 * e + 1] : 0.f, kFmtBF16, hp[e], lp[e]);
    const size_t off = (((static_cast<size_t>(b) * chunks4 + pl * (p.C / 8) + chunk) * Hp + I) * Wp + J) * 8;
    *reinterpret_cast<uint4*>(p.planes + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(p.planes + plane_elems + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
  }
}

// ds[b,i] = ds_conv[b,i] - s[b,i] * sum_o q[b,o] d[b,o]^2 wsq[o,i]      (in place on ds_conv)
struct DsJob {
  float* ds;            // [B,cin] in/out
  const float* s;       // [B,cin]
  const float* q;       // [B,cout]
  const float* d;       // [B,cout]
  const float* wsq;     // [cout,cin]
  int cin, cout;
};
struct DsJobs {
  DsJob job[SGR_MAX_STYLED];
  int n;
};
// 128 outputs per block, the reduction index cut into kRedSplit parts (one per group of 128 threads): a single 512-step
// dependent chain per thread made these two kernels 70-80 us of pure latency.
constexpr int kRedSplit = 4;
__global__ void __launch_bounds__(128 * kRedSplit) ds_finish_kernel(const DsJobs jobs) {
  const DsJob& j = jobs.job[blockIdx.z];
  const int b = blockIdx.y;
  const int part = threadIdx.x >> 7, t = threadIdx.x & 127;
  const int i = blockIdx.x * 128 + t;
  __shared__ float qd[512];
  __shared__ float red[kRedSplit][128];
  for (int o = threadIdx.x; o < j.cout; o += blockDim.x) {
    const float d = j.d[static_cast<size_t>(b) * j.cout + o];
    qd[o] = j.q[static_cast<size_t>(b) * j.cout + o] * d * d;
  }
  __syncthreads();
  float acc = 0.f;
  if (i < j.cin) {
    const int o0 = part * j.cout / kRedSplit, o1 = (part + 1) * j.cout / kRedSplit;
#pragma unroll 8
    for (int o = o0; o < o1; ++o) acc = fmaf(qd[o], __ldg(j.wsq + static_cast<size_t>(o) * j.cin + i), acc);
  }
  red[part][t] = acc;
  __syncthreads();
  if (part != 0 || i >= j.cin) return;
#pragma unroll
  for (int q = 1; q < kRedSplit; ++q) acc += red[q][t];
  const size_t k = static_cast<size_t>(b) * j.cin + i;
  j.ds[k] = j.ds[k] - j.s[k] * acc;
}

// dlatent[b,row,k] += sum_i ds[b,i] * Wm[i,k] / sqrt(512)
struct LatJob {
  const float* ds;      // [B,cin]
  const float* wm;      // [cin,512]
  int cin, row;
};
struct LatJobs {
  LatJob job[SGR_MAX_STYLED + SGR_MAX_RGB];
  int n;
};
__global__ void __launch_bounds__(128 * kRedSplit) dlatent_kernel(const LatJobs jobs, float* __restrict__ dlatent, int latent_stride) {
  const LatJob& j = jobs.job[blockIdx.z];
  const int b = blockIdx.y;
  const int part = threadIdx.x >> 7, t = threadIdx.x & 127;
  const int k = blockIdx.x * 128 + t;                        // 0..511
  __shared__ float dsv[512];
  __shared__ float red[kRedSplit][128];
  for (int i = threadIdx.x; i < j.cin; i += blockDim.x) dsv[i] = j.ds[static_cast<size_t>(b) * j.cin + i];
  __syncthreads();
  float acc = 0.f;
  const int i0 = part * j.cin / kRedSplit, i1 = (part + 1) * j.cin / kRedSplit;
#pragma unroll 8
  for (int i = i0; i < i1; ++i) acc = fmaf(dsv[i], __ldg(j.wm + static_cast<size_t>(i) * SGR_STYLE_DIM + k), acc);
  red[part][t] = acc;
  __syncthreads();
  if (part != 0) return;
#pragma unroll
  for (int q = 1; q < kRedSplit; ++q) acc += red[q][t];
  atomicAdd(dlatent + static_cast<size_t>(b) * latent_stride + static_cast<size_t>(j.row) * SGR_STYLE_DIM + k,
            acc * 0.044194173824159216f);
}

struct ZeroJobs {
  float* ptr[2 * (SGR_MAX_STYLED + SGR_MAX_RGB)];
  int count[2 * (SGR_MAX_STYLED + SGR_MAX_RGB)];
  int n;
};
__global__ void zero_jobs_kernel(const ZeroJobs jobs) {
  float* p = jobs.ptr[blockIdx.x];
  for (int i = threadIdx.x; i < jobs.count[blockIdx.x]; i += blockDim.x) p[i] = 0.f;
}

// d(modulation.weight)[i,k] = sum_b ds[b,i] * latent[b,row,k] / sqrt(512),  d(modulation.bias)[i] = sum_b ds[b,i]
// (EqualLinear backward, model.py:148-157) for every styled and ToRGB layer in one launch.  grid (4, cin_max, jobs).
struct ModGradJob {
  const float* ds;      // [B,cin]
  float* g_weight;      // [cin,512]
  float* g_bias;        // [cin]
  int cin, row;
};
struct ModGradJobs {
  ModGradJob job[SGR_MAX_STYLED + SGR_MAX_RGB];
  int n;
};
__global__ void __launch_bounds__(128) modgrad_kernel(const ModGradJobs jobs, const float* __restrict__ latent,
                                                      int latent_stride, int batch) {
  const ModGradJob& j = jobs.job[blockIdx.z];
  const int i = blockIdx.y;
  if (i >= j.cin) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f, sum = 0.f;
  for (int b = 0; b < batch; ++b) {
    const float d = __ldg(j.ds + static_cast<size_t>(b) * j.cin + i);
    sum += d;
    acc = fmaf(d, __ldg(latent + static_cast<size_t>(b) * latent_stride + static_cast<size_t>(j.row) * SGR_STYLE_DIM + k), acc);
  }
  j.g_weight[static_cast<size_t>(i) * SGR_STYLE_DIM + k] = acc * 0.044194173824159216f;
  if (k == 0) j.g_bias[i] = sum;
}

// d(ConstantInput.input)[c,p] = sum_b g_input[b,c,p] * s_0[b,c]      (model.py:290-300; the input of conv1 is input * s_0)
__global__ void const_grad_kernel(const float* __restrict__ g_input, const float* __restrict__ s0, int batch, int C,
                                  float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * 16) return;
  const int c = idx / 16;
  float acc = 0.f;
  for (int b = 0; b < batch; ++b)
    acc = fmaf(__ldg(g_input + static_cast<size_t>(b) * C * 16 + idx), __ldg(s0 + static_cast<size_t>(b) * C + c), acc);
  out[idx] = acc;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct BwdPlan {
  size_t style_off[SGR_MAX_STYLED], rgbstyle_off[SGR_MAX_RGB], demod_off[SGR_MAX_STYLED];
  size_t s2_off, coef_off;                       // scratch outputs of the table kernel (unused by backward)
  size_t ds_off[SGR_MAX_STYLED], q_off[SGR_MAX_STYLED], dsrgb_off[SGR_MAX_RGB];
  size_t acc_begin, acc_end;
  size_t grgb_off[SGR_MAX_RGB];
  size_t gz_off, gx_off[2];
  size_t splitk_off;
  size_t planes_off;
  size_t total;
};

static int plan_backward(const sgr_synthesis* net, int batch, BwdPlan* pl) {
  if (!net || batch <= 0 || net->n_styled < 1 || net->n_styled > SGR_MAX_STYLED || net->n_styled != 2 * net->n_rgb - 1) {
    set_error("synthesis_backward: bad network descriptor");
    return 1;
  }
  const size_t B = static_cast<size_t>(batch);
  size_t off = 0, cmax = 0;
  for (int l = 0; l < net->n_styled; ++l) {
    const sgr_styled_layer& L = net->styled[l];
    pl->style_off[l] = off; off = align_up(off + B * L.cin * 4, 256);
    pl->demod_off[l] = off; off = align_up(off + B * L.cout * 4, 256);
    cmax = cmax > static_cast<size_t>(L.cout) ? cmax : L.cout;
  }
  for (int r = 0; r < net->n_rgb; ++r) {
    pl->rgbstyle_off[r] = off; off = align_up(off + B * net->rgb[r].cin * 4, 256);
  }
  pl->s2_off = off; off = align_up(off + B * cmax * 4, 256);
  pl->coef_off = off; off = align_up(off + B * 3 * cmax * 4, 256);
  pl->acc_begin = off;
  for (int l = 0; l < net->n_styled; ++l) {
    const sgr_styled_layer& L = net->styled[l];
    pl->ds_off[l] = off; off = align_up(off + B * L.cin * 4, 256);
    pl->q_off[l] = off; off = align_up(off + B * L.cout * 4, 256);
  }
  for (int r = 0; r < net->n_rgb; ++r) {
    pl->dsrgb_off[r] = off; off = align_up(off + B * net->rgb[r].cin * 4, 256);
  }
  pl->acc_end = off;
  size_t max_out = 0, max_in = 0;
  for (int r = 0; r < net->n_rgb; ++r) {
    const size_t res = static_cast<size_t>(4) << r;
    pl->grgb_off[r] = off; off = align_up(off + B * 3 * res * res * 4, 256);
  }
  for (int l = 0; l < net->n_styled; ++l) {
    const size_t res_out = static_cast<size_t>(4) << ((l + 1) / 2);
    const size_t res_in = net->styled[l].up ? res_out / 2 : res_out;
    const size_t eo = B * net->styled[l].cout * res_out * res_out;
    const size_t ei = B * net->styled[l].cin * res_in * res_in;
    max_out = max_out > eo ? max_out : eo;
    max_in = max_in > ei ? max_in : ei;
  }
  pl->gz_off = off; off = align_up(off + max_out * 4, 256);
  for (int i = 0; i < 2; ++i) {
    pl->gx_off[i] = off; off = align_up(off + max_in * 4, 256);
  }
  pl->splitk_off = off; off = align_up(off + kSplitKScratchBytes, 256);
  size_t max_planes = 0;
  for (int l = 0; l < net->n_styled; ++l)
    if (net->styled[l].up == 2) {
      const size_t res_in = static_cast<size_t>(4) << ((l + 1) / 2 - 1);
      const size_t e = B * 4 * net->styled[l].cout * (res_in + 1) * (res_in + 1) * 4;
      max_planes = max_planes > e ? max_planes : e;
    }
  pl->planes_off = off; off = align_up(off + max_planes, 256);
  pl->total = off;
  return 0;
}

}  // namespace sgr

using namespace sgr;

extern "C" {

// caller-allocated scratch of the parameter-gradient path:
//   [modulated layer input as a C8 operand | slice partials of the weight-gradient GEMM]
static size_t wgrad_xs_bytes(const sgr_synthesis* net, int batch) {
  size_t max_in = 0;
  for (int l = 0; l < net->n_styled; ++l) {
    const size_t res_out = static_cast<size_t>(4) << ((l + 1) / 2);
    const size_t res_in = net->styled[l].up ? res_out / 2 : res_out;
    max_in = std::max(max_in, static_cast<size_t>(batch) * net->styled[l].cin * res_in * res_in);
  }
  return align_up(max_in * 4, 256);
}

size_t sgr_synthesis_wgrad_scratch_bytes(const sgr_synthesis* net, int batch) {
  if (!net || batch <= 0 || net->n_styled < 1 || net->n_styled > SGR_MAX_STYLED) return 0;
  size_t part = 0;
  for (int l = 0; l < net->n_styled; ++l) part = std::max(part, wgrad_scratch_bytes(net->styled[l].cout, net->styled[l].cin));
  return wgrad_xs_bytes(net, batch) + align_up(part, 256);
}

size_t sgr_synthesis_backward_workspace_bytes(const sgr_synthesis* net, int batch) {
  BwdPlan pl;
  if (plan_backward(net, batch, &pl)) return 0;
  return pl.total;
}

int sgr_synthesis_backward(const sgr_synthesis* net, const float* latent, int batch, const float* const* feats,
                           const float* grad_image, float* dlatent, void* workspace, size_t workspace_bytes,
                           void* stream) {
  return sgr_synthesis_backward_ex(net, latent, batch, feats, grad_image, dlatent, workspace, workspace_bytes, nullptr, stream);
}

int sgr_synthesis_backward_ex(const sgr_synthesis* net, const float* latent, int batch, const float* const* feats,
                              const float* grad_image, float* dlatent, void* workspace, size_t workspace_bytes,
                              const sgr_backward_extras* extras, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available: libsgr has no CPU fallback");
    return 1;
  }
  BwdPlan pl;
  if (plan_backward(net, batch, &pl)) return 1;
  if (!latent || !feats || !grad_image || !dlatent || !workspace || workspace_bytes < pl.total ||
      (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("synthesis_backward: null pointer, unaligned or too small workspace (%zu < %zu)", workspace_bytes, pl.total);
    return 1;
  }
  for (int l = 0; l < net->n_styled; ++l)
    if (!feats[l] || !net->styled[l].w_packed_t) {
      set_error("synthesis_backward: layer %d lacks its saved activation or adjoint-packed weights", l);
      return 1;
    }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  const int latent_stride = net->n_latent * SGR_STYLE_DIM;
  const int L = net->n_styled, R = net->n_rgb;
  const sgr_param_grads* pg = extras ? extras->params : nullptr;
  const size_t xs_bytes = wgrad_xs_bytes(net, batch);
  if (pg) {
    if (!extras->wgrad_scratch || (reinterpret_cast<uintptr_t>(extras->wgrad_scratch) & 255) != 0 ||
        extras->wgrad_scratch_bytes < sgr_synthesis_wgrad_scratch_bytes(net, batch)) {
      set_error("synthesis_backward: parameter gradients need %zu bytes of 256-byte aligned wgrad_scratch",
                sgr_synthesis_wgrad_scratch_bytes(net, batch));
      return 1;
    }
    bool ok = pg->g_const_input != nullptr;
    for (int l = 0; l < net->n_styled && ok; ++l) {
      const sgr_styled_param_grads& g = pg->styled[l];
      ok = g.weight && g.g_weight && g.g_mod_weight && g.g_mod_bias && g.g_noise_weight && g.g_act_bias &&
           net->styled[l].up != 1;
    }
    for (int r = 0; r < net->n_rgb && ok; ++r) {
      const sgr_rgb_param_grads& g = pg->rgb[r];
      ok = g.g_weight && g.g_mod_weight && g.g_mod_bias && g.g_bias;
    }
    if (!ok) {
      set_error("synthesis_backward: sgr_param_grads has a null pointer, or a layer is packed in polyphase mode (up == 1)");
      return 1;
    }
    // the atomically accumulated sums start from zero: one launch for all of them (40 memsets cost 0.1 ms of host time,
    // which a batch-1 optimize_g step is bound by)
    ZeroJobs zj;
    zj.n = 0;
    for (int l = 0; l < net->n_styled; ++l) {
      zj.ptr[zj.n] = pg->styled[l].g_noise_weight; zj.count[zj.n++] = 1;
      zj.ptr[zj.n] = pg->styled[l].g_act_bias; zj.count[zj.n++] = net->styled[l].cout;
    }
    for (int r = 0; r < net->n_rgb; ++r) {
      zj.ptr[zj.n] = pg->rgb[r].g_weight; zj.count[zj.n++] = net->rgb[r].cin * 3;
      zj.ptr[zj.n] = pg->rgb[r].g_bias; zj.count[zj.n++] = 3;
    }
    zero_jobs_kernel<<<zj.n, 128, 0, static_cast<cudaStream_t>(stream)>>>(zj);
    count_launch();
    if (!check_launch("zero_jobs_kernel")) return 1;
  }

  // styles + demod (recomputed: cheaper than keeping them alive between forward and backward)
  StyleJobs sj;
  sj.n = 0;
  for (int l = 0; l < L; ++l)
    sj.job[sj.n++] = StyleJob{net->styled[l].mod_weight, net->styled[l].mod_bias, F(pl.style_off[l]), net->styled[l].cin,
                              net->styled[l].latent_row};
  for (int r = 0; r < R; ++r)
    sj.job[sj.n++] = StyleJob{net->rgb[r].mod_weight, net->rgb[r].mod_bias, F(pl.rgbstyle_off[r]), net->rgb[r].cin,
                              net->rgb[r].latent_row};
  if (style_jobs_launch(sj, latent, latent_stride, batch, st)) return 1;
  TableJobs tj;
  tj.n = L;
  for (int l = 0; l < L; ++l) {
    TableJob& j = tj.job[l];
    memset(&j, 0, sizeof(j));
    j.s = F(pl.style_off[l]);
    j.wsq = net->styled[l].wsq;
    j.demod = F(pl.demod_off[l]);
    j.s2 = F(pl.s2_off);
    j.rgb_coef = F(pl.coef_off);
    j.cin = net->styled[l].cin;
    j.cout = net->styled[l].cout;
  }
  if (table_jobs_launch(tj, batch, st)) return 1;
  if (cudaMemsetAsync(ws + pl.acc_begin, 0, pl.acc_end - pl.acc_begin, st) != cudaSuccess ||
      cudaMemsetAsync(dlatent, 0, static_cast<size_t>(batch) * latent_stride * 4, st) != cudaSuccess) {
    set_error("synthesis_backward: memset failed");
    return 1;
  }

  // gradient of every ToRGB output: grgb_R = gimage; grgb_{r-1} = adjoint of the 2x FIR upsample of the skip
  // (upfirdn2d with flipped taps, down 2, pad (1,1); op/upfirdn2d.py:32-43,112-117)
  const float* grgb[SGR_MAX_RGB];
  grgb[R - 1] = grad_image;
  for (int r = R - 1; r >= 1; --r) {
    const int res = 4 << r;
    // sgr upfirdn2d applies the TRUE convolution with its taps, so the adjoint needs the flipped FIR
    if (!net->rgb[r].fir_flipped) {
      set_error("synthesis_backward: rgb %d lacks fir_flipped", r);
      return 1;
    }
    if (upfirdn2d_launch(grgb[r], F(pl.grgb_off[r - 1]), net->rgb[r].fir_flipped, batch * 3, res, res, 1, 2, 1, 1, 4, 4, st))
      return 1;
    grgb[r - 1] = F(pl.grgb_off[r - 1]);
  }

  // layer chain, last to first
  const float* gx_next = nullptr;      // dgrad output of layer l+1 == gradient wrt (a_l * s_{l+1})
  int gx_cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    const sgr_styled_layer& Ly = net->styled[l];
    const int res_out = 4 << ((l + 1) / 2);
    const int res_in = Ly.up ? res_out / 2 : res_out;
    const int r = (l + 1) / 2;
    BwdActParams p;
    memset(&p, 0, sizeof(p));
    p.B = batch; p.C = Ly.cout; p.H = res_out; p.W = res_out;
    p.gx = gx_next;
    p.s_next = gx_next ? F(pl.style_off[l + 1]) : nullptr;
    p.a = feats[l];
    p.a_bstride = static_cast<long long>(Ly.cout) * res_out * res_out;
    if (!Ly.up) {
      p.grgb = grgb[r];
      p.wrgb = net->rgb[r].weight;
      p.s_rgb = F(pl.rgbstyle_off[r]);
      p.ds_rgb = F(pl.dsrgb_off[r]);
    }
    p.demod = F(pl.demod_off[l]);
    p.bias = Ly.act_bias;
    p.noise = Ly.noise;
    p.noise_bstride = Ly.noise_batch_stride;
    p.noise_w = Ly.noise_weight;
    if (extras && extras->gfeats) p.out_ga = extras->gfeats[l];
    if (pg) {      // per-channel parameter sums ride on this pass (bwd_act_kernel<true>)
      p.g_act_bias = pg->styled[l].g_act_bias;
      p.g_noise_w = pg->styled[l].g_noise_weight;
      if (!Ly.up) {
        p.g_wrgb = pg->rgb[r].g_weight;
        p.g_rgb_bias = pg->rgb[r].g_bias;
      }
    }
    const bool scatter = Ly.up == 2;      // gather adjoint: FIR^T to parity planes, then the 9 real taps
    if (scatter) p.out_gz4 = F(pl.gz_off);
    else p.out_c8 = reinterpret_cast<__nv_bfloat16*>(ws + pl.gz_off);
    p.s2d = (Ly.up && !scatter) ? 1 : 0;
    p.act = 1;
    p.ds_next = gx_next ? F(pl.ds_off[l + 1]) : nullptr;
    p.q = F(pl.q_off[l]);
    if (bwd_act_launch(p, st)) return 1;

    // data gradient through the shared-weight convolution (adjoint-packed weights)
    sgr_conv_args a;
    memset(&a, 0, sizeof(a));
    a.batch = batch;
    if (scatter) {
      if (!Ly.fir) {
        set_error("synthesis_backward: layer %d (scatter up-conv) lacks its blur kernel", l);
        return 1;
      }
      UpBwdPrepParams q;
      q.B = batch; q.C = Ly.cout; q.H = res_in; q.W = res_in;
      q.gz = F(pl.gz_off);
      q.fir = Ly.fir;
      q.planes = reinterpret_cast<__nv_bfloat16*>(ws + pl.planes_off);
      dim3 grid(((res_in + 1) * (res_in + 1) + 127) / 128, Ly.cout / 8, batch);
      up_bwd_prepare_kernel<<<grid, 128, 0, st>>>(q);
      count_launch();
      if (!check_launch("up_bwd_prepare_kernel")) return 1;
    }
    // generator-parameter gradients (optimize_g): the weight-gradient GEMM of the gz operand just built x the modulated
    // layer input, re-split from the saved fp32 activation of the previous layer (conv1: the constant)
    if (pg) {
      char* xs = static_cast<char*>(extras->wgrad_scratch);
      if (l == 0) {
        if (const_input_launch(net->const_input, F(pl.style_off[0]), batch, Ly.cin, SGR_FMT_BF16, xs, st)) return 1;
      } else if (nchw_to_c8_launch(feats[l - 1], F(pl.style_off[l]), xs, batch, Ly.cin, res_in, res_in, 0, SGR_FMT_BF16, st)) {
        return 1;
      }
      sgr_wgrad_args wa;
      memset(&wa, 0, sizeof(wa));
      wa.batch = batch; wa.cin = Ly.cin; wa.cout = Ly.cout; wa.h_in = res_in; wa.w_in = res_in;
      wa.up = scatter ? 2 : 0;
      wa.x_c8 = xs;
      wa.gz_c8 = scatter ? ws + pl.planes_off : ws + pl.gz_off;
      wa.gw = pg->styled[l].g_weight;
      wa.scratch = xs + xs_bytes;
      wa.scratch_bytes = extras->wgrad_scratch_bytes - xs_bytes;
      WgradFinish wf;
      wf.weight = pg->styled[l].weight; wf.q = F(pl.q_off[l]); wf.demod = F(pl.demod_off[l]); wf.style = F(pl.style_off[l]);
      wf.batch = batch; wf.scale = 1.f / sqrtf(static_cast<float>(Ly.cin) * 9.f);
      if (wgrad_launch(&wa, &wf, st)) return 1;
    }
    a.cin = scatter ? Ly.cout : (Ly.up ? 4 * Ly.cout : Ly.cout);
    a.cout = Ly.cin;
    a.h_in = res_in;
    a.w_in = res_in;
    a.ksize = 3;
    a.up = scatter ? 3 : 0;
    a.act = 0;
    a.act_gain = 1.f;
    a.operand_format = SGR_FMT_BF16;      // gradients have no a-priori range: bf16 split (fp32 exponent)
    a.out_format = SGR_FMT_BF16;
    a.column_tile = Ly.column_tile_t;
    a.x_c8 = scatter ? ws + pl.planes_off : ws + pl.gz_off;
    a.w_packed = Ly.w_packed_t;
    a.out_f32 = F(pl.gx_off[gx_cur]);
    a.splitk_scratch = ws + pl.splitk_off;
    a.splitk_scratch_bytes = kSplitKScratchBytes;
    if (sgr_modconv_forward(&a, stream)) return 1;
    gx_next = F(pl.gx_off[gx_cur]);
    gx_cur = 1 - gx_cur;
  }
  if (extras && extras->g_input && gx_next &&
      cudaMemcpyAsync(extras->g_input, gx_next, static_cast<size_t>(batch) * net->styled[0].cin * 16 * 4, cudaMemcpyDeviceToDevice,
                      st) != cudaSuccess) {
    set_error("synthesis_backward: copy of g_input failed");
    return 1;
  }
  if (pg && gx_next) {
    const int C0 = net->styled[0].cin;
    const_grad_kernel<<<(C0 * 16 + 127) / 128, 128, 0, st>>>(gx_next, F(pl.style_off[0]), batch, C0, pg->g_const_input);
    count_launch();
    if (!check_launch("const_grad_kernel")) return 1;
  }
  // conv term of ds_0: the input of conv1 is the constant 4x4 tensor (batch stride 0)
  {
    BwdActParams p;
    memset(&p, 0, sizeof(p));
    p.B = batch; p.C = net->styled[0].cin; p.H = 4; p.W = 4;
    p.gx = gx_next;
    p.s_next = F(pl.style_off[0]);
    p.a = net->const_input;
    p.a_bstride = 0;
    p.act = 0;
    p.ds_next = F(pl.ds_off[0]);
    if (bwd_act_launch(p, st)) return 1;
  }

  // demodulation term, then the modulation linears back to the latent rows
  DsJobs dj;
  dj.n = L;
  int cin_max = 0;
  for (int l = 0; l < L; ++l) {
    dj.job[l] = DsJob{F(pl.ds_off[l]), F(pl.style_off[l]), F(pl.q_off[l]), F(pl.demod_off[l]), net->styled[l].wsq,
                      net->styled[l].cin, net->styled[l].cout};
    cin_max = cin_max > net->styled[l].cin ? cin_max : net->styled[l].cin;
    if (net->styled[l].cout > 512) {
      set_error("synthesis_backward: cout > 512 unsupported");
      return 1;
    }
  }
  ds_finish_kernel<<<dim3((cin_max + 127) / 128, batch, L), 128 * kRedSplit, 0, st>>>(dj);
  count_launch();
  if (!check_launch("ds_finish_kernel")) return 1;
  if (extras) {
    for (int l = 0; l < L; ++l)
      if (extras->ds_styled && extras->ds_styled[l] &&
          cudaMemcpyAsync(extras->ds_styled[l], F(pl.ds_off[l]), static_cast<size_t>(batch) * net->styled[l].cin * 4,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("synthesis_backward: copy of ds failed");
        return 1;
      }
    for (int r = 0; r < R; ++r)
      if (extras->ds_rgb && extras->ds_rgb[r] &&
          cudaMemcpyAsync(extras->ds_rgb[r], F(pl.dsrgb_off[r]), static_cast<size_t>(batch) * net->rgb[r].cin * 4,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("synthesis_backward: copy of ds_rgb failed");
        return 1;
      }
  }
  if (pg) {
    ModGradJobs mj;
    mj.n = 0;
    int cmax = 0;
    for (int l = 0; l < L; ++l) {
      mj.job[mj.n++] = ModGradJob{F(pl.ds_off[l]), pg->styled[l].g_mod_weight, pg->styled[l].g_mod_bias, net->styled[l].cin,
                                  net->styled[l].latent_row};
      cmax = std::max(cmax, net->styled[l].cin);
    }
    for (int r = 0; r < R; ++r) {
      mj.job[mj.n++] = ModGradJob{F(pl.dsrgb_off[r]), pg->rgb[r].g_mod_weight, pg->rgb[r].g_mod_bias, net->rgb[r].cin,
                                  net->rgb[r].latent_row};
      cmax = std::max(cmax, net->rgb[r].cin);
    }
    modgrad_kernel<<<dim3(SGR_STYLE_DIM / 128, cmax, mj.n), 128, 0, st>>>(mj, latent, latent_stride, batch);
    count_launch();
    if (!check_launch("modgrad_kernel")) return 1;
  }
  LatJobs lj;
  lj.n = 0;
  for (int l = 0; l < L; ++l)
    lj.job[lj.n++] = LatJob{F(pl.ds_off[l]), net->styled[l].mod_weight, net->styled[l].cin, net->styled[l].latent_row};
  for (int r = 0; r < R; ++r)
    lj.job[lj.n++] = LatJob{F(pl.dsrgb_off[r]), net->rgb[r].mod_weight, net->rgb[r].cin, net->rgb[r].latent_row};
  dlatent_kernel<<<dim3(SGR_STYLE_DIM / 128, batch, lj.n), 128 * kRedSplit, 0, st>>>(lj, dlatent, latent_stride);
  count_launch();
  return check_launch("dlatent_kernel") ? 0 : 1;
}

}  // extern "C"

// Small HBM/latency-bound kernels around the tensor-core convolution:
//   weight packing (fp32 -> bf16 hi/lo slabs, polyphase folding of conv_transpose2d + FIR), per-sample styles
//   (EqualLinear modulation), demodulation coefficients, epilogue tables, constant input, layout conversion.
#include <math.h>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

// ------------------------------------------------------------------------------------------------ weight packing
// Effective kernel value for GEMM column n / reduction index (tap, i).
//   plain:      Wp[o][tap=(ky,kx)][i] = scale * W[o][i][ky][kx]                         (F.conv2d, model.py:269)
//   up:         Wp[phase*cout+o][tap=(ty,tx)][i] = Kc[o][i][py+2-2dy][px+2-2dx],  Kc = (scale*W) (*) fir (6x6 full conv)
//               reproduces conv_transpose2d(stride 2) followed by upfirdn2d(pad=(1,1))   (model.py:254-257)
//   transpose:  the adjoint operators (data gradient): columns = cin, reduction over (tap, cout[*4 phases]).
__device__ float effective_weight(const float* __restrict__ w, const float* __restrict__ fir, int cout, int cin,
                                  int ks, int up, int transpose, float scale, int n, int tap, int kidx) {
  const int dy = ks == 3 ? tap / 3 - 1 : 0;
  const int dx = ks == 3 ? tap % 3 - 1 : 0;
  if (!transpose) {
    if (!up) {
      return scale * w[((static_cast<size_t>(n) * cin + kidx) * ks + (dy + ks / 2)) * ks + (dx + ks / 2)];
    }
    const int phase = n / cout, o = n % cout;
    const int a = (phase >> 1) + 2 - 2 * dy, b = (phase & 1) + 2 - 2 * dx;   // index into the 6x6 merged kernel
    float acc = 0.f;
    const float* wk = w + (static_cast<size_t>(o) * cin + kidx) * 9;
    for (int u = 0; u < 3; ++u)
      for (int v = 0; v < 3; ++v) {
        const int fa = a - u, fb = b - v;
        if (fa >= 0 && fa < 4 && fb >= 0 && fb < 4) acc += scale * wk[u * 3 + v] * fir[fa * 4 + fb];
      }
    return acc;
  }
  // adjoint: column n = input channel i; reduction index = (phase,) output channel o; tap offsets negate.
  if (up == 2) {
    // gather adjoint of the scatter up-conv: tap t reads parity plane (ee x4, eo x2, oe x2, oo) of the FIR^T-filtered
    // gradient at shift (a,b) and multiplies W[o][i][ky][kx] with (ky,kx) = (pu + 2a, pv + 2b) (kScatterAdjTaps)
    const int ky = tap < 4 ? 2 * (tap >> 1) : (tap < 6 ? 2 * (tap - 4) : 1);
    const int kx = tap < 4 ? 2 * (tap & 1) : (tap < 6 ? 1 : (tap < 8 ? 2 * (tap - 6) : 1));
    return scale * w[((static_cast<size_t>(kidx) * cin + n) * 3 + ky) * 3 + kx];
  }
  if (!up) {
    const int o = kidx;
    return scale * w[((static_cast<size_t>(o) * cin + n) * ks + (ks / 2 - dy)) * ks + (ks / 2 - dx)];
  }
  const int phase = kidx / cout, o = kidx % cout;
  // forward: y[2m+py] += Kc[py+2-2d] x[m+d]  =>  gx[m'] = sum_d Kc[py+2-2d] gy_phase[m'-d]; as a correlation with
  // offset e = -d reading gy_phase[m'+e]: Kc index = py+2+2e.
  const int a = (phase >> 1) + 2 + 2 * dy, b = (phase & 1) + 2 + 2 * dx;
  float acc = 0.f;
  const float* wk = w + (static_cast<size_t>(o) * cin + n) * 9;
  for (int u = 0; u < 3; ++u)
    for (int v = 0; v < 3; ++v) {
      const int fa = a - u, fb = b - v;
      if (fa >= 0 && fa < 4 && fb >= 0 && fb < 4) acc += scale * wk[u * 3 + v] * fir[fa * 4 + fb];
    }
  return acc;
}

// One thread per 8-element (16 B) row of the packed image; writes the hi and the lo plane entry.
// packed layout: [n_tile][tap][kc][plane][chunk(4)][n_local(NT)][8] for NT > 64, [n_tile][tap][kc][chunk][plane][n_local][8] for NT <= 64
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ fir, int cout, int cin,
                                   int ks, int up, int transpose, int n_total, int k_total, int nt, float scale,
                                   int fmt, float wscale, __nv_bfloat16* __restrict__ packed) {
  const int ntaps = ks * ks;
  const int kchunks = k_total / kBlockK;
  const long long rows = static_cast<long long>(n_total / nt) * ntaps * kchunks * 4 * nt;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows) return;
  long long t = idx;
  const int nl = static_cast<int>(t % nt); t /= nt;
  const int chunk = static_cast<int>(t % 4); t /= 4;
  const int kc = static_cast<int>(t % kchunks); t /= kchunks;
  const int tap = static_cast<int>(t % ntaps); t /= ntaps;
  const int ntile = static_cast<int>(t);
  const int n = ntile * nt + nl;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k0 = kc * kBlockK + chunk * 8 + 2 * e;
    const float v0 = effective_weight(w, fir, cout, cin, ks, up, transpose, scale, n, tap, k0);
    const float v1 = effective_weight(w, fir, cout, cin, ks, up, transpose, scale, n, tap, k0 + 1);
    split2(v0 * wscale, v1 * wscale, fmt, hi[e], lo[e]);
  }
  const size_t slab = (static_cast<size_t>(ntile) * ntaps + tap) * kchunks + kc;        // stage index
  const size_t row_in_plane = static_cast<size_t>(chunk) * nt + nl;
  uint4* dst = reinterpret_cast<uint4*>(packed) + slab * (2 * 4 * nt);
  if (nt > 64) {                     // [plane][chunk][n]
    dst[row_in_plane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dst[4 * nt + row_in_plane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  } else {                           // [chunk][plane][n]: hi and lo rows of a chunk are adjacent, so [W_hi | W_lo] is ONE
    dst[(2 * chunk) * nt + nl] = make_uint4(hi[0], hi[1], hi[2], hi[3]);          // 2*nt-row operand (modconv_halo_sm100.cu)
    dst[(2 * chunk + 1) * nt + nl] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Scatter up-conv packing (modconv_sm100.cu, mode 2).  Column tile = 4 parity blocks [oe|ee|eo|oo] of CT = nt/4 output
// channels; shift s = (a,b) (a = s>>1 rows, b = s&1 columns) multiplies x[I-a, J-b] and owns the tap (ky,kx) =
// (pu + 2a, pv + 2b) of every block whose parity (pu,pv) keeps that tap inside the 3x3 kernel:
//   s=0: all four blocks (nt rows)   s=1: [oe|ee] (nt/2)   s=2: [ee|eo] (nt/2)   s=3: [ee] (nt/4)
// conv_transpose2d semantics out[2i+ky, 2j+kx] += x[i,j] W[o,i,ky,kx] (model.py:248-254, no flip).
// packed layout: [n_tile][kc][shift][plane][chunk(4)][n_local(n_s)][8]; one thread per 16-byte row.
__global__ void pack_weight_up2_kernel(const float* __restrict__ w, int cout, int cin, int nt, float scale, int fmt,
                                       float wscale, __nv_bfloat16* __restrict__ packed) {
  const int ct = nt / 4;
  const int kchunks = cin / kBlockK;
  const int rows_per_kc = 9 * ct;                              // nt + nt/2 + nt/2 + nt/4
  const long long rows = static_cast<long long>(cout / ct) * kchunks * rows_per_kc * 4;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows) return;
  // decode: idx = ((ntile * 4chunks... ) keep it simple: enumerate (ntile, kc, chunk, r) with r in [0, 9ct)
  long long t = idx;
  const int r = static_cast<int>(t % rows_per_kc); t /= rows_per_kc;
  const int chunk = static_cast<int>(t % 4); t /= 4;
  const int kc = static_cast<int>(t % kchunks); t /= kchunks;
  const int ntile = static_cast<int>(t);
  int sft, nl;
  if (r < nt) { sft = 0; nl = r; }
  else if (r < nt + nt / 2) { sft = 1; nl = r - nt; }
  else if (r < 2 * nt) { sft = 2; nl = r - nt - nt / 2; }
  else { sft = 3; nl = r - 2 * nt; }
  const int n_s = nt >> ((sft + 1) >> 1);
  const int block = (sft >= 2 ? 1 : 0) + nl / ct;              // index into [oe, ee, eo, oo]
  const int pu = (block == 0 || block == 3) ? 1 : 0;
  const int pv = (block >= 2) ? 1 : 0;
  const int ky = pu + 2 * (sft >> 1), kx = pv + 2 * (sft & 1);
  const int o = ntile * ct + nl % ct;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i0 = kc * kBlockK + chunk * 8 + 2 * e;
    const float v0 = scale * w[((static_cast<size_t>(o) * cin + i0) * 3 + ky) * 3 + kx];
    const float v1 = scale * w[((static_cast<size_t>(o) * cin + i0 + 1) * 3 + ky) * 3 + kx];
    split2(v0 * wscale, v1 * wscale, fmt, hi[e], lo[e]);
  }
  const int prefix = sft == 0 ? 0 : (sft == 1 ? nt : (sft == 2 ? nt + nt / 2 : 2 * nt));
  // 16-byte row index of the slab (ntile, kc, sft); a slab row is 8 x 16 B (2 planes x 4 chunks)
  const size_t slab = ((static_cast<size_t>(ntile) * kchunks + kc) * rows_per_kc + prefix) * 8;
  uint4* dst = reinterpret_cast<uint4*>(packed) + slab;
  dst[chunk * n_s + nl] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  dst[4 * n_s + chunk * n_s + nl] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void wsq_kernel(const float* __restrict__ w, int cout, int cin, int ntaps, float scale,
                           float* __restrict__ wsq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * cin) return;
  const float* wk = w + static_cast<size_t>(idx) * ntaps;
  float acc = 0.f;
  for (int t = 0; t < ntaps; ++t) {
    const float v = scale * wk[t];
    acc = fmaf(v, v, acc);
  }
  wsq[idx] = acc;
}

// ------------------------------------------------------------------------------------------------ styles

// One warp per (job, i, group of 4 samples): s[b,i] = <latent[b,row,:], W[i,:]> / sqrt(512) + bias[i]
// (model.py:148-157 with lr_mul = 1).  The weight row stays in registers; the four samples are independent chains.
constexpr int kStyleBatchGroup = 4;
constexpr int kStyleGroupsPerWarp = 4;
__global__ void style_kernel(const StyleJobs jobs, const float* __restrict__ latent, int latent_stride, int batch) {
  const StyleJob& j = jobs.job[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= j.cin) return;
  const float4* wrow = reinterpret_cast<const float4*>(j.mod_weight + static_cast<size_t>(i) * SGR_STYLE_DIM);
  float4 wv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) wv[q] = __ldg(wrow + q * 32 + lane);
  const float bias = __ldg(j.mod_bias + i);
  const float scale = 0.044194173824159216f;   // 1/sqrt(512)
  // the weight row in registers serves kStyleGroupsPerWarp groups of samples (one group per warp re-read the 12 MB of
  // modulation weights once per four samples: 35 us per step at B = 32)
  for (int g = 0; g < kStyleGroupsPerWarp; ++g) {
  const int b0 = (blockIdx.z * kStyleGroupsPerWarp + g) * kStyleBatchGroup;
  if (b0 >= batch) break;
  float acc[kStyleBatchGroup];
#pragma unroll
  for (int u = 0; u < kStyleBatchGroup; ++u) {
    acc[u] = 0.f;
    const int b = min(b0 + u, batch - 1);
    const float4* lrow = reinterpret_cast<const float4*>(latent + static_cast<size_t>(b) * latent_stride +
                                                         static_cast<size_t>(j.latent_row) * SGR_STYLE_DIM);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 l = __ldg(lrow + q * 32 + lane);
      acc[u] = fmaf(l.x, wv[q].x, fmaf(l.y, wv[q].y, fmaf(l.z, wv[q].z, fmaf(l.w, wv[q].w, acc[u]))));
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int u = 0; u < kStyleBatchGroup; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], off);
  if (lane < kStyleBatchGroup && b0 + lane < batch) {
    float v = acc[0];
#pragma unroll
    for (int u = 1; u < kStyleBatchGroup; ++u) v = lane == u ? acc[u] : v;
    j.out[static_cast<size_t>(b0 + lane) * j.cin + i] = fmaf(v, scale, bias);
  }
  }
}

// ------------------------------------------------------------------------------------------------ demod + tables

// Demodulation + epilogue tables: d[b,o] = rsqrt(sum_i s[b,i]^2 wsq[o,i] + 1e-8) as a small GEMM per layer.
// Block = (layer, 32 output channels, up to 32 samples): s^2 of the samples is staged once in shared memory; a warp owns
// one output at a time, its lanes hold coalesced slices of the wsq row (cin <= 512: 16 registers) and accumulate all 32
// samples in registers; a 31-shuffle reduce-scatter then leaves sample b's sum in lane b.
// (One warp per (b, o) re-read each wsq row once per sample: 270 MB of L2 traffic at B = 32.)
constexpr int kTableOutPerBlock = 16;
constexpr int kTableMaxCinPerLane = 16;          // cin <= 512
// NB = samples per block: 32, or 4 for the small-batch calls (batch <= 4: a block of 32 mostly-empty sample slots made the
// first launch of a batch-1 frame 16 us long)
template <int NB>
__global__ void __launch_bounds__(256) table_kernel(const TableJobs jobs, int batch) {
  extern __shared__ float s2[];                        // [NB][cin]
  const TableJob& j = jobs.job[blockIdx.y];
  const int o0 = blockIdx.x * kTableOutPerBlock;
  if (o0 >= j.cout) return;
  const int b0 = blockIdx.z * NB;
  const int nb = min(NB, batch - b0);
  for (int i = threadIdx.x; i < j.cin; i += blockDim.x) {           // no integer division: column i, all samples
#pragma unroll (NB < 8 ? NB : 8)
    for (int b = 0; b < NB; ++b) {
      const float v = b < nb ? __ldg(j.s + static_cast<size_t>(b0 + b) * j.cin + i) : 0.f;
      s2[b * j.cin + i] = v * v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float kSqrt2 = 1.4142135623730951f;
#pragma unroll 1
  for (int oo = warp; oo < kTableOutPerBlock / 2; oo += 8) {
    // two outputs per pass share every s^2 read (the loop is shared-memory-load bound)
    const int oa = o0 + oo, ob = oa + kTableOutPerBlock / 2;
    if (oa >= j.cout) break;
    const bool has_b = ob < j.cout;
    const float* qa = j.wsq + static_cast<size_t>(oa) * j.cin;
    const float* qb = j.wsq + static_cast<size_t>(has_b ? ob : oa) * j.cin;
    float acc[NB], bcc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = bcc[b] = 0.f;
#pragma unroll
    for (int t = 0; t < kTableMaxCinPerLane; ++t) {
      if (32 * t < j.cin) {                          // warp-uniform
        const float wa = __ldg(qa + lane + 32 * t), wb = __ldg(qb + lane + 32 * t);
        const float* col = s2 + lane + 32 * t;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float sv = col[b * j.cin];
          acc[b] = fmaf(sv, wa, acc[b]);
          bcc[b] = fmaf(sv, wb, bcc[b]);
        }
      }
    }
    // reduce-scatter over the lanes: after the step with offset `off` a lane keeps the half of the samples whose bit
    // `off` equals its own, so lane b ends with the total of sample b in acc[0] / bcc[0]
    if constexpr (NB == 32) {
#pragma unroll
      for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < n / 2; ++k) {
          const float send = up ? acc[k] : acc[k + n / 2], keep = up ? acc[k + n / 2] : acc[k];
          acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          const float send2 = up ? bcc[k] : bcc[k + n / 2], keep2 = up ? bcc[k + n / 2] : bcc[k];
          bcc[k] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
        }
      }
    } else {                                           // few samples: plain butterfly per sample, lane b keeps sample b
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
          bcc[k] += __shfl_xor_sync(0xffffffffu, bcc[k], off);
        }
#pragma unroll
      for (int k = 1; k < NB; ++k) {
        acc[0] = lane == k ? acc[k] : acc[0];
        bcc[0] = lane == k ? bcc[k] : bcc[0];
      }
    }
    if (lane < nb) {
      const int b = b0 + lane;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !has_b) break;
        const int o = h ? ob : oa;
        const float a = h ? bcc[0] : acc[0];
        j.demod[static_cast<size_t>(b) * j.cout + o] = rsqrtf(a + 1e-8f);
        if (j.s_next) j.s2[static_cast<size_t>(b) * j.cout + o] = kSqrt2 * __ldg(j.s_next + static_cast<size_t>(b) * j.cout + o);
        if (j.s_rgb) {
          const float sr = kSqrt2 * __ldg(j.s_rgb + static_cast<size_t>(b) * j.cout + o) * rsqrtf(static_cast<float>(j.cout));
          for (int c = 0; c < 3; ++c)
            j.rgb_coef[(static_cast<size_t>(b) * 3 + c) * j.cout + o] = sr * __ldg(j.w_rgb + c * j.cout + o);
        }
      }
    }
  }
}

__global__ void demod_kernel(const float* __restrict__ s, const float* __restrict__ wsq, int cin, int cout,
                             float* __restrict__ d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  const int b = blockIdx.y;
  if (o >= cout) return;
  float acc = 0.f;
  for (int i = lane; i < cin; i += 32) {
    const float sv = __ldg(s + static_cast<size_t>(b) * cin + i);
    acc = fmaf(sv * sv, __ldg(wsq + static_cast<size_t>(o) * cin + i), acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) d[static_cast<size_t>(b) * cout + o] = rsqrtf(acc + 1e-8f);
}

// ------------------------------------------------------------------------------------------------ layout conversion
// NCHW fp32 (* scale[b,c]) -> C8 hi/lo planes.  One thread per (b, chunk, y, x) writes 16 B to each plane.
// s2d: channel' = phase*C + c at (y/2, x/2), phase = (y&1)*2 + (x&1).
__global__ void nchw_to_c8_kernel(const float* __restrict__ x, const float* __restrict__ scale, int batch, int C, int H,
                                  int W, int s2d, int fmt, float ascale, __nv_bfloat16* __restrict__ out) {
  const long long total = static_cast<long long>(batch) * (C / 8) * H * W;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  long long t = idx;
  const int xx = static_cast<int>(t % W); t /= W;
  const int yy = static_cast<int>(t % H); t /= H;
  const int ch = static_cast<int>(t % (C / 8)); t /= (C / 8);
  const int b = static_cast<int>(t);
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = ch * 8 + 2 * e + h;
      float val = __ldg(x + ((static_cast<size_t>(b) * C + c) * H + yy) * W + xx);
      if (scale) val *= __ldg(scale + static_cast<size_t>(b) * C + c);
      v[h] = val;
    }
    split2(v[0] * ascale, v[1] * ascale, fmt, hi[e], lo[e]);
  }
  size_t off, plane;
  if (!s2d) {
    off = ((static_cast<size_t>(b) * (C / 8) + ch) * H + yy) * W + xx;
    plane = static_cast<size_t>(batch) * (C / 8) * H * W;
  } else {
    const int phase = (yy & 1) * 2 + (xx & 1);
    const int H2 = H / 2, W2 = W / 2, C4 = 4 * C;
    off = ((static_cast<size_t>(b) * (C4 / 8) + phase * (C / 8) + ch) * H2 + (yy >> 1)) * W2 + (xx >> 1);
    plane = static_cast<size_t>(batch) * (C4 / 8) * H2 * W2;
  }
  uint4* o4 = reinterpret_cast<uint4*>(out);
  o4[off] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  o4[plane + off] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ConstantInput (model.py:290-300) times the first layer's style, straight into C8 planes.
__global__ void const_input_kernel(const float* __restrict__ cinput, const float* __restrict__ s, int batch, int C,
                                   int fmt, float ascale, __nv_bfloat16* __restrict__ out) {
  const int total = batch * (C / 8) * 16;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pix = idx % 16;
  const int ch = (idx / 16) % (C / 8);
  const int b = idx / (16 * (C / 8));
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = ch * 8 + 2 * e;
    const float v0 = __ldg(cinput + c * 16 + pix) * __ldg(s + static_cast<size_t>(b) * C + c);
    const float v1 = __ldg(cinput + (c + 1) * 16 + pix) * __ldg(s + static_cast<size_t>(b) * C + c + 1);
    split2(v0 * ascale, v1 * ascale, fmt, hi[e], lo[e]);
  }
  uint4* o4 = reinterpret_cast<uint4*>(out);
  o4[idx] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  o4[static_cast<size_t>(total) + idx] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------------ host wrappers
int pack_weight_launch(const float* w, const float* fir, int cout, int cin, int ks, int up, int transpose, int fmt,
                       int nt_req, void* packed, float* wsq, cudaStream_t st) {
  const float scale = 1.f / sqrtf(static_cast<float>(cin) * ks * ks);
  if (up == 2 && !transpose) {
    const int nt2 = up2_nt(cout);
    const long long rows2 = static_cast<long long>(cout / (nt2 / 4)) * (cin / kBlockK) * (9 * nt2 / 4) * 4;
    pack_weight_up2_kernel<<<static_cast<unsigned>((rows2 + 255) / 256), 256, 0, st>>>(
        w, cout, cin, nt2, scale, fmt, w_scale(fmt), static_cast<__nv_bfloat16*>(packed));
    count_launch();
    if (!check_launch("pack_weight_up2_kernel")) return 1;
    if (wsq) {
      wsq_kernel<<<(cout * cin + 255) / 256, 256, 0, st>>>(w, cout, cin, ks * ks, scale, wsq);
      count_launch();
      if (!check_launch("wsq_kernel")) return 1;
    }
    return 0;
  }
  if (up && !(up == 2 && transpose)) up = 1;
  const int n_total = transpose ? cin : cout * (up ? 4 : 1);
  const int k_total = transpose ? cout * (up == 1 ? 4 : 1) : cin;
  const int nt = nt_req > 0 ? nt_req : pick_nt(n_total);
  const long long rows = static_cast<long long>(n_total) * ks * ks * (k_total / 8);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((rows + threads - 1) / threads);
  pack_weight_kernel<<<blocks, threads, 0, st>>>(w, fir, cout, cin, ks, up, transpose, n_total, k_total, nt, scale, fmt,
                                                 w_scale(fmt), static_cast<__nv_bfloat16*>(packed));
  count_launch();
  if (!check_launch("pack_weight_kernel")) return 1;
  if (wsq) {
    wsq_kernel<<<(cout * cin + 255) / 256, 256, 0, st>>>(w, cout, cin, ks * ks, scale, wsq);
    count_launch();
    if (!check_launch("wsq_kernel")) return 1;
  }
  return 0;
}

int style_jobs_launch(const StyleJobs& jobs, const float* latent, int latent_stride, int batch, cudaStream_t st) {
  int cmax = 0;
  for (int i = 0; i < jobs.n; ++i) cmax = cmax > jobs.job[i].cin ? cmax : jobs.job[i].cin;
  dim3 grid((cmax + 7) / 8, jobs.n, (batch + kStyleBatchGroup * kStyleGroupsPerWarp - 1) / (kStyleBatchGroup * kStyleGroupsPerWarp));
  style_kernel<<<grid, 256, 0, st>>>(jobs, latent, latent_stride, batch);
  count_launch();
  return check_launch("style_kernel") ? 0 : 1;
}

int table_jobs_launch(const TableJobs& jobs, int batch, cudaStream_t st) {
  int cmax = 0;
  for (int i = 0; i < jobs.n; ++i) cmax = cmax > jobs.job[i].cout ? cmax : jobs.job[i].cout;
  int cin_max = 0;
  for (int i = 0; i < jobs.n; ++i) {
    cin_max = cin_max > jobs.job[i].cin ? cin_max : jobs.job[i].cin;
    if (jobs.job[i].cin % 32 != 0 || jobs.job[i].cin > 32 * kTableMaxCinPerLane) {
      set_error("table_kernel: cin %d must be a multiple of 32 and <= %d", jobs.job[i].cin, 32 * kTableMaxCinPerLane);
      return 1;
    }
  }
  if (batch <= 4) {                                    // [4][cin <= 512] floats: within the default shared-memory limit
    dim3 grid4((cmax + kTableOutPerBlock - 1) / kTableOutPerBlock, jobs.n, 1);
    table_kernel<4><<<grid4, 256, 4 * cin_max * 4, st>>>(jobs, batch);
    count_launch();
    return check_launch("table_kernel") ? 0 : 1;
  }
  const int smem = 32 * cin_max * 4;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    if (cudaFuncSetAttribute(table_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("table_kernel: cin %d needs %d bytes of shared memory", cin_max, smem);
      return 1;
    }
    configured_smem = smem;
  }
  dim3 grid((cmax + kTableOutPerBlock - 1) / kTableOutPerBlock, jobs.n, (batch + 31) / 32);
  table_kernel<32><<<grid, 256, smem, st>>>(jobs, batch);
  count_launch();
  return check_launch("table_kernel") ? 0 : 1;
}

int demod_launch(const float* s, const float* wsq, int batch, int cin, int cout, float* d, cudaStream_t st) {
  dim3 grid((cout + 7) / 8, batch);
  demod_kernel<<<grid, 256, 0, st>>>(s, wsq, cin, cout, d);
  count_launch();
  return check_launch("demod_kernel") ? 0 : 1;
}

int nchw_to_c8_launch(const float* x, const float* scale, void* out, int batch, int C, int H, int W, int s2d, int fmt,
                      cudaStream_t st) {
  const long long total = static_cast<long long>(batch) * (C / 8) * H * W;
  nchw_to_c8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, scale, batch, C, H, W, s2d, fmt,
                                                                                act_scale(fmt), static_cast<__nv_bfloat16*>(out));
  count_launch();
  return check_launch("nchw_to_c8_kernel") ? 0 : 1;
}

int const_input_launch(const float* cinput, const float* s, int batch, int C, int fmt, void* out, cudaStream_t st) {
  const int total = batch * (C / 8) * 16;
  const_input_kernel<<<(total + 127) / 128, 128, 0, st>>>(cinput, s, batch, C, fmt, act_scale(fmt), static_cast<__nv_bfloat16*>(out));
  count_launch();
  return check_launch("const_input_kernel") ? 0 : 1;
}

}  // namespace sgr

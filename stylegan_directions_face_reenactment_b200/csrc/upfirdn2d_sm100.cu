// HBM-bound kernels: upfirdn2d (replaces op/upfirdn2d_kernel.cu:52-272), fused bias + leaky-relu
// (replaces op/fused_bias_act_kernel.cu:18-99) and the ToRGB tail (bias + 2x FIR-upsampled skip, model.py:355-359).
#include <stdlib.h>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

// ------------------------------------------------------------------------------------------------ upfirdn2d
// y[p, oy, ox] = sum_{ky,kx} taps[kh-1-ky][kw-1-kx] * u[oy*down + ky - pad0][ox*down + kx - pad0]
// where u is x zero-inserted by `up` (u[i*up][j*up] = x[i][j]) and zero outside (upfirdn2d_kernel.cu:77,100,125-129).
//
// Vectorised over x: each thread produces VX adjacent outputs of one row; a warp covers 32*VX contiguous outputs, so
// loads and stores are fully coalesced.  For up == down == 1 the horizontal window of a thread overlaps its lane
// neighbour's: each lane loads only its own VX-aligned span (+ the right-most lane the halo) and the KW-1 halo values
// come from the next lane through warp shuffles, so every input element is read from L1/L2 once per row pass.
template <int UP, int DOWN, int KH, int KW, int VX>
__global__ void upfirdn2d_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ taps,
                                 int planes, int in_h, int in_w, int out_h, int out_w, int pad0) {
  __shared__ float sk[KH * KW];
  if (threadIdx.x < KH * KW) sk[threadIdx.x] = taps[(KH - 1 - threadIdx.x / KW) * KW + (KW - 1 - threadIdx.x % KW)];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * VX;
  const int oy = blockIdx.y;
  const bool col_ok = ox0 < out_w;
  for (int pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const float* xp = x + static_cast<size_t>(pl) * in_h * in_w;
    float acc[VX];
#pragma unroll
    for (int v = 0; v < VX; ++v) acc[v] = 0.f;
    if (UP == 1 && DOWN == 1) {
      // window columns ox0 - pad0 .. ox0 - pad0 + VX + KW - 2
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const int iy = oy + ky - pad0;
        const bool row_ok = iy >= 0 && iy < in_h;          // warp-uniform
        float own[VX];
#pragma unroll
        for (int v = 0; v < VX; ++v) {
          const int ix = ox0 + v - pad0;
          own[v] = (row_ok && ix >= 0 && ix < in_w) ? __ldg(xp + static_cast<size_t>(iy) * in_w + ix) : 0.f;
        }
        float win[VX + KW - 1];
#pragma unroll
        for (int v = 0; v < VX; ++v) win[v] = own[v];
#pragma unroll
        for (int h = 0; h < KW - 1; ++h) {
          // halo value h = neighbour lane's own[h]; the last lane of the warp loads it directly
          float nb = __shfl_down_sync(0xffffffffu, own[h % VX], 1 + h / VX);
          if (lane + 1 + h / VX > 31) {
            const int ix = ox0 + VX + h - pad0;
            nb = (row_ok && ix >= 0 && ix < in_w) ? __ldg(xp + static_cast<size_t>(iy) * in_w + ix) : 0.f;
          }
          win[VX + h] = nb;
        }
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          const float t = sk[ky * KW + kx];
#pragma unroll
          for (int v = 0; v < VX; ++v) acc[v] = fmaf(t, win[v + kx], acc[v]);
        }
      }
    } else {
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const int uy = oy * DOWN + ky - pad0;
        if (uy < 0 || uy % UP != 0) continue;
        const int iy = uy / UP;
        if (iy >= in_h) continue;
#pragma unroll
        for (int v = 0; v < VX; ++v) {
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) {
            const int ux = (ox0 + v) * DOWN + kx - pad0;
            if (ux < 0 || ux % UP != 0) continue;
            const int ix = ux / UP;
            if (ix >= in_w) continue;
            acc[v] = fmaf(sk[ky * KW + kx], __ldg(xp + static_cast<size_t>(iy) * in_w + ix), acc[v]);
          }
        }
      }
    }
    if (col_ok) {
      float* yp = y + (static_cast<size_t>(pl) * out_h + oy) * out_w + ox0;
      if (VX == 4 && (out_w % 4) == 0) {
        *reinterpret_cast<float4*>(yp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
#pragma unroll
        for (int v = 0; v < VX; ++v)
          if (ox0 + v < out_w) yp[v] = acc[v];
      }
    }
  }
}

// up == down == 1 with a 4x4 FIR (the Blur after an up-conv, model.py:72-88; the mode that moves the most bytes): row-walking
// variant.  A thread owns VX = 4 adjacent output columns and walks ROWS output rows downwards with a 4-row register window,
// so every input element is loaded ONCE (the kernel above re-reads each input row for its four output rows and re-does the
// halo shuffles: ~30 instructions per output, issue-bound at 27 % of the HBM rate); per output row: 4 loads, 3 halo shuffles from the
// lane to the right, the FMAs and one 16-byte store.  When the taps are an outer product gy (x) gx (checked in the kernel; always the
// case for make_kernel's FIRs) the FIR runs separably: horizontal pass on the incoming row (4 FMAs / output), vertical
// pass over the window of row results (4 FMAs / output) instead of 16.
template <bool SEP, int ROWS>
__device__ __forceinline__ void upfirdn2d_rows_body(const float* __restrict__ x, float* __restrict__ y, const float* sk, int planes,
                                                    int in_h, int in_w, int out_h, int out_w, int pad0) {
  constexpr int VX = 4, K = 4;
  const int lane = threadIdx.x & 31;
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * VX;
  const int oy0 = (blockIdx.y * blockDim.y + threadIdx.y) * ROWS;
  if (oy0 >= out_h) return;                                  // whole warps (blockDim.x is a multiple of 32)
  const bool col_ok = ox0 < out_w;
  const bool vec_ok = (out_w % 4) == 0;
  float gy[K], gx[K], t2[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    gy[i] = SEP ? sk[K * K + i] : 0.f;
    gx[i] = SEP ? sk[K * K + K + i] : 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) t2[i][j] = SEP ? 0.f : sk[i * K + j];
  }
  const int rows = min(ROWS, out_h - oy0);
  for (int pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const float* xp = x + static_cast<size_t>(pl) * in_h * in_w;
    float* yp = y + static_cast<size_t>(pl) * out_h * out_w;
    // window of the last K rows: SEP keeps horizontally filtered rows (VX values), otherwise raw rows (VX + K - 1 values)
    float win[K][SEP ? VX : VX + K - 1];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < (SEP ? VX : VX + K - 1); ++j) win[i][j] = 0.f;
    // input rows oy0 - pad0 .. oy0 + rows + K - 2 - pad0; output row oy is complete once input row oy + K - 1 - pad0 is in.
    // (unroll 2: the loads of the next row are independent of this row's arithmetic and get hoisted above it; an explicit
    //  three-row prefetch queue cost 120 registers and ran 1.5x slower)
#pragma unroll 2
    for (int r = 0; r < rows + K - 1; ++r) {
      const int iy = oy0 + r - pad0;
      const bool row_ok = iy >= 0 && iy < in_h;            // warp-uniform
      float own[VX];
#pragma unroll
      for (int v = 0; v < VX; ++v) {
        const int ix = ox0 + v - pad0;
        own[v] = (row_ok && ix >= 0 && ix < in_w) ? __ldg(xp + static_cast<size_t>(iy) * in_w + ix) : 0.f;
      }
      float raw[VX + K - 1];
#pragma unroll
      for (int v = 0; v < VX; ++v) raw[v] = own[v];
#pragma unroll
      for (int h = 0; h < K - 1; ++h) {                     // halo: the first K-1 values of the lane to the right
        float nb = __shfl_down_sync(0xffffffffu, own[h], 1);
        if (lane == 31) {
          const int ix = ox0 + VX + h - pad0;
          nb = (row_ok && ix >= 0 && ix < in_w) ? __ldg(xp + static_cast<size_t>(iy) * in_w + ix) : 0.f;
        }
        raw[VX + h] = nb;
      }
#pragma unroll
      for (int i = 0; i < K - 1; ++i)
#pragma unroll
        for (int j = 0; j < (SEP ? VX : VX + K - 1); ++j) win[i][j] = win[i + 1][j];
      if (SEP) {
#pragma unroll
        for (int v = 0; v < VX; ++v)
          win[K - 1][v] = fmaf(gx[3], raw[v + 3], fmaf(gx[2], raw[v + 2], fmaf(gx[1], raw[v + 1], gx[0] * raw[v])));
      } else {
#pragma unroll
        for (int j = 0; j < VX + K - 1; ++j) win[K - 1][j] = raw[j];
      }
      if (r >= K - 1 && col_ok) {
        const int oy = oy0 + r - (K - 1);
        float acc[VX];
#pragma unroll
        for (int v = 0; v < VX; ++v) {
          if (SEP) {
            acc[v] = fmaf(gy[3], win[3][v], fmaf(gy[2], win[2][v], fmaf(gy[1], win[1][v], gy[0] * win[0][v])));
          } else {
            float a = 0.f;
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
              for (int kx = 0; kx < K; ++kx) a = fmaf(t2[ky][kx], win[ky][v + kx], a);
            acc[v] = a;
          }
        }
        float* dst = yp + static_cast<size_t>(oy) * out_w + ox0;
        if (vec_ok) {
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
#pragma unroll
          for (int v = 0; v < VX; ++v)
            if (ox0 + v < out_w) dst[v] = acc[v];
        }
      }
    }
  }
}

template <int ROWS>
__global__ void __launch_bounds__(256) upfirdn2d_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             const float* __restrict__ taps, int planes, int in_h, int in_w,
                                                             int out_h, int out_w, int pad0, int allow_sep) {
  constexpr int K = 4;
  __shared__ float sk[K * K + 2 * K];
  __shared__ int sep;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    // flipped taps; separable form gy[ky] * gx[kx] with gy = row sums, gx = column sums / total; exact rank-1 test (tolerance
    // 1e-6 of the largest tap, as the host check of the scatter path, model.py up_mode())
    float tot = 0.f, mx = 0.f;
    for (int i = 0; i < K * K; ++i) {
      sk[i] = taps[(K - 1 - i / K) * K + (K - 1 - i % K)];
      tot += sk[i];
      mx = fmaxf(mx, fabsf(sk[i]));
    }
    for (int a = 0; a < K; ++a) {
      float rs = 0.f, cs = 0.f;
      for (int i = 0; i < K; ++i) {
        rs += sk[a * K + i];
        cs += sk[i * K + a];
      }
      sk[K * K + a] = rs;
      sk[K * K + K + a] = tot != 0.f ? cs / tot : 0.f;
    }
    float dev = 0.f;
    for (int i = 0; i < K * K; ++i) dev = fmaxf(dev, fabsf(sk[i] - sk[K * K + i / K] * sk[K * K + K + i % K]));
    sep = (allow_sep && tot != 0.f && dev <= 1e-6f * mx) ? 1 : 0;
  }
  __syncthreads();
  if (sep)
    upfirdn2d_rows_body<true, ROWS>(x, y, sk, planes, in_h, in_w, out_h, out_w, pad0);
  else
    upfirdn2d_rows_body<false, ROWS>(x, y, sk, planes, in_h, in_w, out_h, out_w, pad0);
}

static void launch_ufd_rows(const float* x, float* y, const float* taps, int planes, int in_h, int in_w, int out_h, int out_w,
                            int pad0, int allow_sep, cudaStream_t st) {
  constexpr int ROWS = 32;
  const int tx = out_w >= 4 * 64 ? 64 : 32;                   // threads along x (4 outputs each)
  const int strips = (out_h + ROWS - 1) / ROWS;
  const int ty = strips >= 4 ? 4 : (strips >= 2 ? 2 : 1);
  dim3 block(tx, ty);
  dim3 grid((out_w + tx * 4 - 1) / (tx * 4), (strips + ty - 1) / ty, planes < 65535 ? planes : 65535);
  upfirdn2d_rows_kernel<ROWS><<<grid, block, 0, st>>>(x, y, taps, planes, in_h, in_w, out_h, out_w, pad0, allow_sep);
}

// Any factors / tap counts (correctness path for shapes outside the generator's four modes).
__global__ void upfirdn2d_generic_kernel(const float* __restrict__ x, float* __restrict__ y,
                                         const float* __restrict__ taps, int planes, int in_h, int in_w, int out_h,
                                         int out_w, int up, int down, int pad0, int kh, int kw) {
  const long long total = static_cast<long long>(planes) * out_h * out_w;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % out_w);
    const int oy = static_cast<int>((idx / out_w) % out_h);
    const long long pl = idx / (static_cast<long long>(out_w) * out_h);
    const float* xp = x + pl * in_h * in_w;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int uy = oy * down + ky - pad0;
      if (uy < 0 || uy % up != 0 || uy / up >= in_h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ux = ox * down + kx - pad0;
        if (ux < 0 || ux % up != 0 || ux / up >= in_w) continue;
        acc = fmaf(__ldg(taps + (kh - 1 - ky) * kw + (kw - 1 - kx)), __ldg(xp + static_cast<size_t>(uy / up) * in_w + ux / up), acc);
      }
    }
    y[idx] = acc;
  }
}

template <int UP, int DOWN>
static void launch_ufd44(const float* x, float* y, const float* taps, int planes, int in_h, int in_w, int out_h,
                         int out_w, int pad0, cudaStream_t st) {
  constexpr int VX = 4;
  const int threads = out_w >= 512 ? 128 : (out_w >= 256 ? 64 : 32);
  dim3 grid((out_w + threads * VX - 1) / (threads * VX), out_h, planes < 65535 ? planes : 65535);
  upfirdn2d_kernel<UP, DOWN, 4, 4, VX><<<grid, threads, 0, st>>>(x, y, taps, planes, in_h, in_w, out_h, out_w, pad0);
}

int upfirdn2d_launch(const float* x, float* y, const float* taps, int planes, int in_h, int in_w, int up, int down,
                     int pad0, int pad1, int kh, int kw, cudaStream_t st) {
  const int out_h = (in_h * up + pad0 + pad1 - kh + down) / down;
  const int out_w = (in_w * up + pad0 + pad1 - kw + down) / down;
  if (planes <= 0 || out_h <= 0 || out_w <= 0 || up < 1 || down < 1 || kh < 1 || kw < 1) {
    set_error("upfirdn2d: empty or invalid shape (planes=%d out=%dx%d up=%d down=%d k=%dx%d)", planes, out_h, out_w,
              up, down, kh, kw);
    return 1;
  }
  if (kh == 4 && kw == 4 && out_h <= 65535 && up == 1 && down == 1) {
    // SGR_UPFIRDN_ROWS: 0 = the per-row kernel, 1 = row-walking kernel with the 16-tap arithmetic only, default 2 = row-walking,
    // separable arithmetic when the taps are an outer product
    static const int rows_mode = [] { const char* e = getenv("SGR_UPFIRDN_ROWS"); return e ? atoi(e) : 2; }();
    if (rows_mode == 0)
      launch_ufd44<1, 1>(x, y, taps, planes, in_h, in_w, out_h, out_w, pad0, st);
    else
      launch_ufd_rows(x, y, taps, planes, in_h, in_w, out_h, out_w, pad0, rows_mode == 2 ? 1 : 0, st);
  } else if (kh == 4 && kw == 4 && out_h <= 65535 && up == 2 && down == 1) {
    launch_ufd44<2, 1>(x, y, taps, planes, in_h, in_w, out_h, out_w, pad0, st);
  } else if (kh == 4 && kw == 4 && out_h <= 65535 && up == 1 && down == 2) {
    launch_ufd44<1, 2>(x, y, taps, planes, in_h, in_w, out_h, out_w, pad0, st);
  } else {
    const long long total = static_cast<long long>(planes) * out_h * out_w;
    const long long blocks = (total + 255) / 256;
    upfirdn2d_generic_kernel<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, st>>>(
        x, y, taps, planes, in_h, in_w, out_h, out_w, up, down, pad0, kh, kw);
  }
  count_launch();
  return check_launch("upfirdn2d_kernel") ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ fused bias act
__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                const float* __restrict__ ref, float* __restrict__ y, long long total, int channels,
                                long long inner, int grad, float slope, float scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = __ldg(x + i);
    if (grad == 0) {
      if (bias) v += __ldg(bias + (i / inner) % channels);
      y[i] = (v > 0.f ? v : v * slope) * scale;
    } else {
      y[i] = (__ldg(ref + i) > 0.f ? v : v * slope) * scale;
    }
  }
}

__global__ void bias_act_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ bias,
                                     const float4* __restrict__ ref, float4* __restrict__ y, long long total4,
                                     int channels, long long inner, int grad, float slope, float scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = __ldg(x + i);
    if (grad == 0) {
      const float b = bias ? __ldg(bias + ((i * 4) / inner) % channels) : 0.f;   // inner % 4 == 0: one channel per vec
      v.x += b; v.y += b; v.z += b; v.w += b;
      v.x = (v.x > 0.f ? v.x : v.x * slope) * scale;
      v.y = (v.y > 0.f ? v.y : v.y * slope) * scale;
      v.z = (v.z > 0.f ? v.z : v.z * slope) * scale;
      v.w = (v.w > 0.f ? v.w : v.w * slope) * scale;
    } else {
      const float4 r = __ldg(ref + i);
      v.x = (r.x > 0.f ? v.x : v.x * slope) * scale;
      v.y = (r.y > 0.f ? v.y : v.y * slope) * scale;
      v.z = (r.z > 0.f ? v.z : v.z * slope) * scale;
      v.w = (r.w > 0.f ? v.w : v.w * slope) * scale;
    }
    y[i] = v;
  }
}

int bias_act_launch(const float* x, const float* bias, const float* ref, float* y, long long outer, int channels,
                    long long inner, int grad, float slope, float scale, cudaStream_t st) {
  const long long total = outer * channels * inner;
  if (total <= 0) return 0;
  if (grad != 0 && !ref) {
    set_error("fused_bias_act: grad=1 needs the saved output");
    return 1;
  }
  const bool vec = (inner % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                         reinterpret_cast<uintptr_t>(ref)) % 16 == 0);
  if (vec) {
    const long long t4 = total / 4;
    const long long blocks = (t4 + 255) / 256;
    bias_act_vec4_kernel<<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), bias, reinterpret_cast<const float4*>(ref), reinterpret_cast<float4*>(y), t4,
        channels, inner, grad, slope, scale);
  } else {
    const long long blocks = (total + 255) / 256;
    bias_act_kernel<<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
        x, bias, ref, y, total, channels, inner, grad, slope, scale);
  }
  count_launch();
  return check_launch("bias_act_kernel") ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ ToRGB tail
// skip_out[b,c,y,x] = sum_slots rgb_partial[slot][b,c,y,x] + bias[c] + upfirdn2d(skip_in, fir, up=2, pad=(2,1))[b,c,y,x]   (model.py:355-359)
// skip_in is [B,3,H/2,W/2] (or NULL for to_rgb1).  4 outputs per thread along x.
__global__ void torgb_tail_kernel(const float* __restrict__ rgb_acc, int slots, const float* __restrict__ bias,
                                  const float* __restrict__ skip_in, const float* __restrict__ fir,
                                  float* __restrict__ out, int planes, int H, int W) {
  __shared__ float sk[16];
  pdl_launch_dependents();
  pdl_wait();                                        // the ToRGB partial sums come from the convolution right before
  if (threadIdx.x < 16) sk[threadIdx.x] = fir ? fir[(3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4)] : 0.f;
  __syncthreads();
  // grid: (x groups of 4 pixels, y, plane) - no integer divisions in the kernel
  const int h2 = H / 2, w2 = W / 2;
  const int xg = blockIdx.x * blockDim.x + threadIdx.x;
  if (xg >= W / 4) return;
  for (int pl = blockIdx.z; pl < planes; pl += gridDim.z) {
    const int x0 = xg * 4;
    const int y = blockIdx.y;
    const float b = __ldg(bias + pl % 3);
    float4 a = __ldg(reinterpret_cast<const float4*>(rgb_acc + (static_cast<size_t>(pl) * H + y) * W + x0));
    for (int s = 1; s < slots; ++s) {           // per-column-tile partial sums, fixed order
      const float4 a2 = __ldg(reinterpret_cast<const float4*>(rgb_acc + ((static_cast<size_t>(s) * planes + pl) * H + y) * W + x0));
      a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
    }
    float acc[4] = {a.x + b, a.y + b, a.z + b, a.w + b};
    if (skip_in) {
      const float* sp = skip_in + static_cast<size_t>(pl) * h2 * w2;
      // u[j] = skip[j/2] for even j; out[y] = sum_k flipped_taps[k] * u[y + k - 2]
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int uy = y + ky - 2;
        if (uy < 0 || (uy & 1) || (uy >> 1) >= h2) continue;
        const float* row = sp + static_cast<size_t>(uy >> 1) * w2;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            const int ux = x0 + v + kx - 2;
            if (ux < 0 || (ux & 1) || (ux >> 1) >= w2) continue;
            acc[v] = fmaf(sk[ky * 4 + kx], __ldg(row + (ux >> 1)), acc[v]);
          }
        }
      }
    }
    *reinterpret_cast<float4*>(out + (static_cast<size_t>(pl) * H + y) * W + x0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

int torgb_tail_launch(const float* rgb_acc, int slots, const float* bias, const float* skip_in, const float* fir,
                      float* out, int batch, int H, int W, cudaStream_t st) {
  if (W % 4 != 0 || H > 65535) {
    set_error("torgb_tail: unsupported image size %dx%d", H, W);
    return 1;
  }
  const int threads = W / 4 >= 128 ? 128 : (W / 4 >= 64 ? 64 : 32);
  const int planes = batch * 3;
  dim3 grid((W / 4 + threads - 1) / threads, H, planes < 65535 ? planes : 65535);
  launch_pdl(torgb_tail_kernel, grid, dim3(threads), 0, st, rgb_acc, slots, bias, skip_in, fir, out, planes, H, W);
  count_launch();
  return check_launch("torgb_tail_kernel") ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ ToRGB tail + output stage
// Last level, when the caller wants bytes: the same slot sum + bias + 2x FIR skip as torgb_tail_kernel (same
// operation order, so the frame is bit-identical to the fp32 one), then generate_image's 256-pooling (generic.py:146-148),
// tensor_to_image's clamp / scale (image_utils.py:97-111) and np.uint8 (utils_inference.py:16) — the fp32 frame is never
// written.  One thread per OUTPUT pixel: lanes run along x, the three channels leave as 3 adjacent bytes.
__global__ void torgb_tail_u8_kernel(const float* __restrict__ rgb_acc, int slots, const float* __restrict__ bias,
                                     const float* __restrict__ skip_in, const float* __restrict__ fir,
                                     unsigned char* __restrict__ out, int batch, int H, int W, int out_h, int out_w) {
  __shared__ float sk[16];
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < 16) sk[threadIdx.x] = fir ? fir[(3 - threadIdx.x / 4) * 4 + (3 - threadIdx.x % 4)] : 0.f;
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= out_w) return;
  const int oy = blockIdx.y, b = blockIdx.z;
  const int fy = H / out_h, fx = W / out_w;
  const int h2 = H / 2, w2 = W / 2;
  const int planes = batch * 3;
  const float inv_area = 1.f / static_cast<float>(fy * fx);
  unsigned char px[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int pl = b * 3 + c;
    const float bc = __ldg(bias + c);
    const float* sp = skip_in ? skip_in + static_cast<size_t>(pl) * h2 * w2 : nullptr;
    float v = 0.f;
    for (int dy = 0; dy < fy; ++dy) {
      const int y = oy * fy + dy;
      for (int dx = 0; dx < fx; ++dx) {
        const int x = ox * fx + dx;
        float a = __ldg(rgb_acc + (static_cast<size_t>(pl) * H + y) * W + x);
        for (int s = 1; s < slots; ++s) a += __ldg(rgb_acc + ((static_cast<size_t>(s) * planes + pl) * H + y) * W + x);
        float acc = a + bc;
        if (sp) {
#pragma unroll
          for (int ky = 0; ky < 4; ++ky) {
            const int uy = y + ky - 2;
            if (uy < 0 || (uy & 1) || (uy >> 1) >= h2) continue;
            const float* row = sp + static_cast<size_t>(uy >> 1) * w2;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
              const int ux = x + kx - 2;
              if (ux < 0 || (ux & 1) || (ux >> 1) >= w2) continue;
              acc = fmaf(sk[ky * 4 + kx], __ldg(row + (ux >> 1)), acc);
            }
          }
        }
        v = (fy == 1 && fx == 1) ? acc : v + acc;
      }
    }
    if (fy != 1 || fx != 1) v *= inv_area;
    v = fminf(fmaxf(v, -1.f), 1.f);
    v = __fdiv_rn(v + 1.f, 2.00001f) * 255.f;
    px[c] = static_cast<unsigned char>(static_cast<int>(v));
  }
  unsigned char* dst = out + ((static_cast<size_t>(b) * out_h + oy) * out_w + ox) * 3;
  dst[0] = px[0]; dst[1] = px[1]; dst[2] = px[2];
}

int torgb_tail_u8_launch(const float* rgb_acc, int slots, const float* bias, const float* skip_in, const float* fir,
                         unsigned char* out, int batch, int H, int W, int out_h, int out_w, cudaStream_t st) {
  if (out_h <= 0 || out_w <= 0 || H % out_h != 0 || W % out_w != 0 || out_h > 65535 || batch > 65535) {
    set_error("torgb_tail_u8: output %dx%d must divide the frame size %dx%d", out_h, out_w, H, W);
    return 1;
  }
  const int threads = out_w >= 128 ? 128 : (out_w >= 64 ? 64 : 32);
  dim3 grid((out_w + threads - 1) / threads, out_h, batch);
  launch_pdl(torgb_tail_u8_kernel, grid, dim3(threads), 0, st, rgb_acc, slots, bias, skip_in, fir, out, batch, H, W, out_h, out_w);
  count_launch();
  return check_launch("torgb_tail_u8_kernel") ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ output stage
// Frames [B,3,H,W] fp32 in [-1,1] -> uint8 [B,H,W,3] exactly as the reference post-processing does on the way to the
// video writer (libs/utilities/image_utils.py:97-111 tensor_to_image + np.uint8 at utils_inference.py:16):
//   v = clamp(x, -1, 1);  v = (v + 1) / (2 + 1e-5) * 255;  uint8(v)  (truncation toward zero)
// optionally after the AdaptiveAvgPool2d(256) of generate_image for larger nets (generic.py:146-148; H, W multiples of
// out_h, out_w: each output pixel is the mean of an (H/out_h) x (W/out_w) block).  One thread per output pixel, the three
// channels of a pixel are written as 3 adjacent bytes (coalesced over x).
__global__ void frames_to_uint8_kernel(const float* __restrict__ x, unsigned char* __restrict__ y, int batch, int H, int W,
                                       int out_h, int out_w) {
  const long long total = static_cast<long long>(batch) * out_h * out_w;
  const int fy = H / out_h, fx = W / out_w;
  const float inv_area = 1.f / static_cast<float>(fy * fx);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % out_w);
    const int oy = static_cast<int>((idx / out_w) % out_h);
    const int b = static_cast<int>(idx / (static_cast<long long>(out_w) * out_h));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* src = x + ((static_cast<size_t>(b) * 3 + c) * H + static_cast<size_t>(oy) * fy) * W + static_cast<size_t>(ox) * fx;
      float v = 0.f;
      if (fy == 1 && fx == 1) {
        v = __ldg(src);
      } else {
        for (int dy = 0; dy < fy; ++dy)
          for (int dx = 0; dx < fx; ++dx) v += __ldg(src + static_cast<size_t>(dy) * W + dx);
        v *= inv_area;
      }
      v = fminf(fmaxf(v, -1.f), 1.f);
      v = __fdiv_rn(v + 1.f, 2.00001f) * 255.f;
      y[idx * 3 + c] = static_cast<unsigned char>(static_cast<int>(v));
    }
  }
}

int frames_to_uint8_launch(const float* x, unsigned char* y, int batch, int H, int W, int out_h, int out_w, cudaStream_t st) {
  if (batch <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || H % out_h != 0 || W % out_w != 0) {
    set_error("frames_to_uint8: output %dx%d must divide the frame size %dx%d", out_h, out_w, H, W);
    return 1;
  }
  const long long total = static_cast<long long>(batch) * out_h * out_w;
  const long long blocks = (total + 255) / 256;
  frames_to_uint8_kernel<<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(x, y, batch, H, W, out_h,
                                                                                                  out_w);
  count_launch();
  return check_launch("frames_to_uint8_kernel") ? 0 : 1;
}

}  // namespace sgr

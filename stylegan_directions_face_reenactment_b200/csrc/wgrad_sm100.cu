// Weight gradient of the modulated 3x3 convolutions on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the cuDNN wgrad that ATen autograd runs for F.conv2d / F.conv_transpose2d(groups = batch) in
// ModulatedConv2d.forward (libs/gan/StyleGAN2/model.py:254,263,269) when the generator itself is fine-tuned
// (optimize_g, libs/optimization.py:25-72).  In the shared-weight form the
// gradient with respect to W_bar = scale * W is ONE dense contraction over all pixels of all samples:
//     plain   gW[o,i,ky,kx] = sum_{b,y,x} gz[b,o,y,x] * xs[b,i,y+ky-1,x+kx-1]                       (zero padding)
//     up      gW[o,i,ky,kx] = sum_{b,y,x} G[b,o,2y+ky,2x+kx] * xs[b,i,y,x]
// with xs = a_{l-1} * s_l (the modulated input), gz = dL/d(conv output) (demodulation already applied) and, for the
// upsampling layers, G = FIR^T(gz) on the (2H+1)^2 grid of conv_transpose2d, read through its four parity planes
// G[2I+pu, 2J+pv] = plane_{2pu+pv}[I,J] — the operand up_bwd_prepare_kernel already builds for the data gradient.
//
// GEMM view: D_tap[o, i] = sum_pix A_tap[pix, o] * B[pix, i]  — M = 128 output channels, N = NT input channels,
// K = pixels.  Both operands are the C8 activation layout [plane(hi,lo)][B][C/8][H][W][8] bf16, i.e. "MN-major" for this
// GEMM: a TMA box [chunk][y][x][8] lands in shared memory as the canonical no-swizzle MN-major UMMA layout (16 B = 8
// channels contiguous, 8 consecutive pixels = one 128 B core matrix, channel chunks SBO apart, the second 8-pixel group of
// a K = 16 step LBO apart), so no transposition pass is needed.  The 3x3 taps are relative shifts between the two
// operands: the xs tile is loaded unshifted and the gz box carries the halo, a tap is a start-address offset into it.
// A work item = (tap group, 128-channel row tile, NT-channel column tile, slice of the pixel tiles); a tap group shares
// one gz box (plain: the three kx of a kernel row; up: the taps of one parity plane) and owns one TMEM accumulator per
// tap.  The work items are planned as ONE wave of equal-work items: the pixel tiles of a tap group are cut into a number of
// slices proportional to its taps (wgrad_plan).  Slices write fp32 partials and wgrad_finish_kernel adds them in slice order
// (deterministic, no atomics).
// fp32 parity: bf16 hi/lo operands, three MMAs per product, like every other gradient GEMM of the library.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over work items): warp0 TMA producer, warp1 MMA issuer, warp2 TMEM
// allocator, warps4-7 epilogue (TMEM -> registers -> partial slot).
#include <algorithm>
#include <string.h>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

constexpr int kWgPix = 64;           // pixels (GEMM K) per pipeline stage: bw x bh = 16 x 4 or 8 x 8
constexpr int kWgMaxStages = 4;

struct WgradGroup {
  int chunk_base;      // first 8-channel chunk of the gz-side operand (parity plane of an up layer)
  int ay, ax;          // origin of the gz box relative to the xs tile origin
  int ntaps;           // <= 4
  int tap_off[4];      // byte offset of the tap's first pixel inside the gz box
  int tap_out[4];      // ky * 3 + kx
  int slices;          // pixel slices of this group (proportional to ntaps: every work item does the same number of MMAs)
  int item_base;       // first work item of this group
};

struct WgradParams {
  int H, W;                      // xs pixel grid
  int bw, bh;                    // xs tile
  int tiles_x, tiles_y, n_ptiles;
  int o_tiles, i_tiles, nt, cout, cin;
  int n_groups, total_items, stages;
  int tap_base[9], tap_slices[9]; // partial slots of output tap t: part[tap_base[t] + s], s < tap_slices[t]
  uint32_t a_bytes, b_bytes;     // gz box / xs box, both planes
  uint32_t a_sbo, a_lbo, a_kstep;
  WgradGroup g[4];
  float* part;                   // [tap slot][cout][cin]
};

struct WgItem {
  int g, ot, it, s, p0, p1;
};
__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
  WgItem w;
  w.g = 0;
#pragma unroll
  for (int g = 1; g < 4; ++g)
    if (g < p.n_groups && item >= p.g[g].item_base) w.g = g;
  const int slices = p.g[w.g].slices;
  item -= p.g[w.g].item_base;
  w.s = item % slices; item /= slices;
  w.it = item % p.i_tiles;
  w.ot = item / p.i_tiles;
  w.p0 = static_cast<int>(static_cast<long long>(w.s) * p.n_ptiles / slices);
  w.p1 = static_cast<int>(static_cast<long long>(w.s + 1) * p.n_ptiles / slices);
  return w;
}

// idesc of umma_idesc with both operands MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t wg_idesc(int n) { return umma_idesc(kFmtBF16, kTileM, n) | (1u << 15) | (1u << 16); }

__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                       const __grid_constant__ CUtensorMap tmap_b, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* tfull = empty + kWgMaxStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  uint8_t* stage_base = smem + 1024;
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  const int S = p.stages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 128);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_items = p.total_items;
  const int tiles_per_image = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const WgItem w = wg_decode(p, item);
        const WgradGroup& G = p.g[w.g];
        for (int pt = w.p0; pt < w.p1; ++pt, ++it) {
          const int b = pt / tiles_per_image;
          const int r = pt - b * tiles_per_image;
          const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
          const int x0 = tx * p.bw, y0 = ty * p.bh;
          const uint32_t s = it % S;
          mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
          uint8_t* sa = stage_base + s * stage_bytes;
          mbar_expect_tx(&full[s], stage_bytes);
          tma_load_5d(sa, &tmap_a, &full[s], (x0 + G.ax) * 8, y0 + G.ay, b, G.chunk_base + w.ot * 16, 0);
          tma_load_5d(sa + p.a_bytes, &tmap_b, &full[s], x0 * 8, y0, b, w.it * (p.nt / 8), 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    // The issuing thread is the limiter once the operands are resident (about 8 cycles per instruction, DESIGN.md
    // section 4): the four descriptors of a K step differ only in their 14-bit start-address field, so they are built once
    // per launch and advanced with 32-bit adds on the low word (16-byte units) — 4 adds + 3 MMAs per K step.
    if (elect_one_sync()) {
      const uint32_t idesc = wg_idesc(p.nt);
      const uint32_t b_sbo = kWgPix * 16;
      const uint64_t a_d0 = umma_desc(smem_u32(stage_base), p.a_lbo, p.a_sbo);
      const uint64_t b_d0 = umma_desc(smem_u32(stage_base) + p.a_bytes, 128, b_sbo);
      const uint32_t a_hiw = static_cast<uint32_t>(a_d0 >> 32), b_hiw = static_cast<uint32_t>(b_d0 >> 32);
      const uint32_t a_low0 = static_cast<uint32_t>(a_d0), b_low0 = static_cast<uint32_t>(b_d0);
      const uint32_t stage16 = stage_bytes >> 4;
      const uint32_t a_plane16 = p.a_sbo;                          // 16 chunks per plane of the gz box: 16 * a_sbo bytes
      const uint32_t b_plane16 = static_cast<uint32_t>(p.nt / 8) * (b_sbo >> 4);
      const uint32_t a_k16 = p.a_kstep >> 4;
      auto desc = [](uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; };
      uint32_t s = 0, phase = 0, icount = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++icount) {
        const WgItem w = wg_decode(p, item);
        const WgradGroup& G = p.g[w.g];
        const int ntaps = G.ntaps;
        mbar_wait(tempty, (icount & 1) ^ 1);
        tc_fence_after();
        for (int pt = w.p0; pt < w.p1; ++pt) {
          mbar_wait(&full[s], phase);
          tc_fence_after();
          const uint32_t a_s = a_low0 + s * stage16, b_s = b_low0 + s * stage16;
          const uint32_t acc0 = pt > w.p0 ? 1u : 0u;
#pragma unroll 1
          for (int t = 0; t < ntaps; ++t) {
            const uint32_t d_tmem = tmem_base + t * p.nt;
            uint32_t a_hi = a_s + (static_cast<uint32_t>(G.tap_off[t]) >> 4), b_hi = b_s;
#pragma unroll
            for (int j = 0; j < kWgPix / 16; ++j) {
              const uint32_t a_lo = a_hi + a_plane16, b_lo = b_hi + b_plane16;
              umma_bf16(d_tmem, desc(a_hiw, a_lo), desc(b_hiw, b_hi), idesc, j > 0 ? 1u : acc0);
              umma_bf16(d_tmem, desc(a_hiw, a_hi), desc(b_hiw, b_lo), idesc, 1);
              umma_bf16(d_tmem, desc(a_hiw, a_hi), desc(b_hiw, b_hi), idesc, 1);
              a_hi += a_k16;
              b_hi += 16;                                          // 256 bytes: the next two 8-pixel groups of the xs tile
            }
          }
          umma_commit(&empty[s]);
          if (++s == static_cast<uint32_t>(S)) {
            s = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int ew = warp - 4;
    uint32_t icount = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++icount) {
      const WgItem w = wg_decode(p, item);
      const WgradGroup& G = p.g[w.g];
      const int o = w.ot * kTileM + ew * 32 + lane;
      mbar_wait(tfull, icount & 1);
      tc_fence_after();
      for (int t = 0; t < G.ntaps; ++t) {
        float* dst = p.part + (static_cast<size_t>(p.tap_base[G.tap_out[t]] + w.s) * p.cout + o) * p.cin + w.it * p.nt;
#pragma unroll 1
        for (int c = 0; c < p.nt; c += 32) {
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + t * p.nt + c, v);
          tmem_ld_wait();
          if (o < p.cout) {
            float4* d4 = reinterpret_cast<float4*>(dst + c);
#pragma unroll
            for (int q = 0; q < 8; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct WgradSlots {
  int base[9], count[9];
};

// gw[o][i][tap] = sum_s part[slot(tap) + s][o][i], slices in order.  With `f` (the synthesis backward pass) the result is the
// finished gradient of the reference parameter conv.weight (model.py:216-218), demodulation term included:
//   dW[o,i,k] = scale * ( gw[o,i,k] - scale * W[o,i,k] * sum_b q[b,o] d[b,o]^2 s[b,i]^2 )
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ part, const WgradSlots slots, int cout,
                                                           int cin, float* __restrict__ gw, const WgradFinish f) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * cin) return;
  const size_t tap_stride = static_cast<size_t>(cout) * cin;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float* src = part + static_cast<size_t>(slots.base[t]) * tap_stride + idx;
    float a = 0.f;
    int s = 0;
    for (; s + 4 <= slots.count[t]; s += 4) {            // four loads in flight, added in slice order
      const float v0 = __ldcs(src + (s + 0) * tap_stride), v1 = __ldcs(src + (s + 1) * tap_stride);
      const float v2 = __ldcs(src + (s + 2) * tap_stride), v3 = __ldcs(src + (s + 3) * tap_stride);
      a = (((a + v0) + v1) + v2) + v3;
    }
    for (; s < slots.count[t]; ++s) a += __ldcs(src + s * tap_stride);
    acc[t] = a;
  }
  if (f.weight) {
    const int o = idx / cin, i = idx - o * cin;
    float dterm = 0.f;
    for (int b = 0; b < f.batch; ++b) {
      const float d = __ldg(f.demod + static_cast<size_t>(b) * cout + o), sv = __ldg(f.style + static_cast<size_t>(b) * cin + i);
      dterm = fmaf(__ldg(f.q + static_cast<size_t>(b) * cout + o) * d * d, sv * sv, dterm);
    }
    dterm *= f.scale;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = f.scale * fmaf(-dterm, __ldg(f.weight + static_cast<size_t>(idx) * 9 + t), acc[t]);
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) gw[static_cast<size_t>(idx) * 9 + t] = acc[t];
}

// ---------------------------------------------------------------------------------------------- host side
// Pixel slices per tap group, proportional to the group's tap count (equal MMAs per work item) and sized for ONE wave of
// work items on `sms` SMs (the finish pass reads every slice, so more slices than needed only cost traffic); limited by the
// number of pixel tiles and by the scratch.  Returns the number of work items, 0 if even one slice per tap does not fit.
static int wgrad_plan(WgradParams* p, int sms, size_t scratch_bytes) {
  const int pairs = p->o_tiles * p->i_tiles;
  int wsum = 0;
  for (int g = 0; g < p->n_groups; ++g) wsum += p->g[g].ntaps;
  const size_t slot_bytes = static_cast<size_t>(p->cout) * p->cin * 4;
  double k = static_cast<double>(sms) / (static_cast<double>(pairs) * wsum);      // slices per tap of weight
  for (int iter = 0; iter < 64; ++iter, k *= 0.9) {
    int items = 0, slots = 0;
    bool all_one = true;
    for (int g = 0; g < p->n_groups; ++g) {
      WgradGroup& G = p->g[g];
      int sl = static_cast<int>(k * G.ntaps);
      sl = std::max(1, std::min(std::min(sl, 128), p->n_ptiles));
      G.slices = sl;
      G.item_base = items;
      items += pairs * sl;
      slots += G.ntaps * sl;
      all_one = all_one && sl == 1;
    }
    if ((items <= sms && static_cast<size_t>(slots) * slot_bytes <= scratch_bytes) || all_one) {
      if (static_cast<size_t>(slots) * slot_bytes > scratch_bytes) return 0;
      int base = 0;
      for (int t = 0; t < 9; ++t) p->tap_slices[t] = 0;
      for (int g = 0; g < p->n_groups; ++g)
        for (int t = 0; t < p->g[g].ntaps; ++t) p->tap_slices[p->g[g].tap_out[t]] = p->g[g].slices;
      for (int t = 0; t < 9; ++t) {
        p->tap_base[t] = base;
        base += p->tap_slices[t];
      }
      p->total_items = items;
      return items;
    }
  }
  return 0;
}

size_t wgrad_scratch_bytes(int cout, int cin) {
  // room for one wave of work items on up to 192 SMs (at most 128 slices per tap; 38 MB on a 512 -> 512 layer)
  const size_t per_slice = static_cast<size_t>(9) * cout * cin * 4;
  const int base = 3 * ((cout + 127) / 128) * std::max(1, cin / 128);
  const int s = std::max(1, std::min(128, (192 + base - 1) / base));
  return per_slice * s;
}

int wgrad_launch(const sgr_wgrad_args* a, const WgradFinish* finish, cudaStream_t stream) {
  const int sms = num_sms();
  if (sms <= 0) {
    set_error("modconv_wgrad: no CUDA device");
    return 1;
  }
  if (!a->x_c8 || !a->gz_c8 || !a->gw || !a->scratch || a->batch <= 0 || a->cin < 32 || a->cin % 32 || a->cout < 8 ||
      a->cout % 8 || (a->up != 0 && a->up != 2) || a->h_in <= 0 || a->w_in <= 0) {
    set_error("modconv_wgrad: bad arguments (cin %d, cout %d, up %d)", a->cin, a->cout, a->up);
    return 1;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.H = a->h_in; p.W = a->w_in;
  p.bw = a->w_in > 8 ? 16 : 8;
  p.bh = kWgPix / p.bw;
  p.tiles_x = (a->w_in + p.bw - 1) / p.bw;
  p.tiles_y = (a->h_in + p.bh - 1) / p.bh;
  p.n_ptiles = a->batch * p.tiles_x * p.tiles_y;
  p.cout = a->cout; p.cin = a->cin;
  p.nt = std::min(a->cin, 128);
  if (a->cin % p.nt) {
    set_error("modconv_wgrad: cin %d is not a multiple of the column tile %d", a->cin, p.nt);
    return 1;
  }
  p.o_tiles = (a->cout + kTileM - 1) / kTileM;
  p.i_tiles = a->cin / p.nt;
  const int abw = a->up ? p.bw + 1 : p.bw + 2, abh = a->up ? p.bh + 1 : p.bh;
  p.a_sbo = static_cast<uint32_t>(abw) * abh * 16;
  p.a_bytes = 2 * 16 * p.a_sbo;
  p.b_bytes = 2 * static_cast<uint32_t>(p.nt / 8) * kWgPix * 16;
  p.a_lbo = p.bw == 16 ? 128 : static_cast<uint32_t>(abw) * 16;
  p.a_kstep = (p.bw == 16 ? 1 : 2) * static_cast<uint32_t>(abw) * 16;
  if (!a->up) {
    p.n_groups = 3;                                  // one kernel row per group: gz box rows y0 + 1 - ky .., columns x0 - 1 ..
    for (int ky = 0; ky < 3; ++ky) {
      WgradGroup& G = p.g[ky];
      G.chunk_base = 0; G.ay = 1 - ky; G.ax = -1; G.ntaps = 3;
      for (int kx = 0; kx < 3; ++kx) {
        G.tap_off[kx] = (2 - kx) * 16;
        G.tap_out[kx] = ky * 3 + kx;
      }
    }
  } else {
    p.n_groups = 4;                                  // one parity plane per group, heaviest (ee, 4 taps) first
    for (int pl = 0; pl < 4; ++pl) {
      WgradGroup& G = p.g[pl];
      const int pu = pl >> 1, pv = pl & 1;
      G.chunk_base = pl * (a->cout / 8); G.ay = 0; G.ax = 0; G.ntaps = 0;
      for (int ky = pu; ky < 3; ky += 2)
        for (int kx = pv; kx < 3; kx += 2) {
          G.tap_off[G.ntaps] = ((ky == 2 ? abw : 0) + (kx == 2 ? 1 : 0)) * 16;
          G.tap_out[G.ntaps] = ky * 3 + kx;
          ++G.ntaps;
        }
    }
  }
  const int total_items = wgrad_plan(&p, sms, a->scratch_bytes);
  if (total_items < 1) {
    set_error("modconv_wgrad: scratch too small (%zu bytes, one slice needs %zu)", a->scratch_bytes,
              static_cast<size_t>(9) * a->cout * a->cin * 4);
    return 1;
  }
  p.part = static_cast<float*>(a->scratch);
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  const int max_smem = 227 * 1024;
  p.stages = std::min<int>(kWgMaxStages, (max_smem - 1024) / stage_bytes);
  if (p.stages < 2) {
    set_error("modconv_wgrad: pipeline stage of %u bytes does not fit twice", stage_bytes);
    return 1;
  }
  const int smem_bytes = 1024 + p.stages * stage_bytes;
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) {
      set_error("modconv_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return 1;
    }
    configured = max_smem;
  }
  CUtensorMap tmap_a, tmap_b;
  const int a_ch = a->up ? 4 * a->cout : a->cout, a_h = a->up ? a->h_in + 1 : a->h_in, a_w = a->up ? a->w_in + 1 : a->w_in;
  if (make_act_tensor_map(&tmap_a, a->gz_c8, a->batch, a_ch, a_h, a_w, abw, abh, 1, 2, 16)) return 1;
  if (make_act_tensor_map(&tmap_b, a->x_c8, a->batch, a->cin, a->h_in, a->w_in, p.bw, p.bh, 1, 2, p.nt / 8)) return 1;
  wgrad_kernel<<<std::min(total_items, sms), 256, smem_bytes, stream>>>(tmap_a, tmap_b, p);
  count_launch();
  if (!check_launch("wgrad_kernel")) return 1;
  WgradFinish f;
  memset(&f, 0, sizeof(f));
  if (finish) f = *finish;
  WgradSlots slots;
  for (int t = 0; t < 9; ++t) {
    slots.base[t] = p.tap_base[t];
    slots.count[t] = p.tap_slices[t];
  }
  wgrad_finish_kernel<<<(a->cout * a->cin + 255) / 256, 256, 0, stream>>>(p.part, slots, a->cout, a->cin, a->gw, f);
  count_launch();
  return check_launch("wgrad_finish_kernel") ? 0 : 1;
}

}  // namespace sgr

using namespace sgr;

extern "C" {

size_t sgr_wgrad_scratch_bytes(int cout, int cin) { return wgrad_scratch_bytes(cout, cin); }

int sgr_modconv_wgrad(const sgr_wgrad_args* args, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available: libsgr has no CPU fallback");
    return 1;
  }
  if (!args) {
    set_error("modconv_wgrad: null arguments");
    return 1;
  }
  return wgrad_launch(args, nullptr, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// Modulated convolution as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces, for one layer, the whole chain of the reference's ModulatedConv2d.forward -> NoiseInjection ->
// FusedLeakyReLU (libs/gan/StyleGAN2/model.py:232-287,331-337, op/fused_bias_act_kernel.cu:18-49) and, for the
// upsampling layers, conv_transpose2d + Blur/upfirdn2d (model.py:246-257, op/upfirdn2d_kernel.cu:52-137), and
// optionally the following ToRGB 1x1 modulated conv (model.py:350-354).
//
// GEMM view:  D[m, n] = sum_{tap, i} A_tap[m, i] * Wp[n, tap, i]
//   m : 128 pixels of a (bw x bh x bb) box of the INPUT-resolution grid (TMEM lane = pixel)
//   n : output channel (plain) or phase*cout + channel (polyphase up-conv: 4 output pixels per input pixel)
//   A : activations already multiplied by the per-sample style, stored as bf16 hi/lo planes in the "C8" layout
//       [plane][B][C/8][H][W][8]; each 3x3 tap is one TMA box load at shifted coordinates (zero OOB fill = padding)
//   Wp: batch-shared weights, pre-split into bf16 hi/lo and pre-arranged in the shared-memory image order, so a
//       stage's weight slab is one 1-D bulk copy.
// fp32 parity needs more than one bf16 pass: every K step issues A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (bf16x3).
//
// Shared-memory operand layout (no swizzle, K-major): [plane][chunk of 8 channels][row][8 bf16]; a row is 16 B,
// rows are contiguous (SBO = 128 B per 8 rows), K chunks are LBO = rows*16 B apart.
//
// The scatter form of the upsampling layers (minimal flops) lives in modconv_scatter_sm100.cu; the resident-halo variant
// for wide 3x3 layers in modconv_halo_sm100.cu.  This kernel serves 1x1, small-resolution 3x3, polyphase and adjoint GEMMs.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles): warp0 = TMA producer, warp1 = MMA issuer,
// warp2 = TMEM allocator, warps4-7 = epilogue (TMEM -> registers -> fused demod/noise/bias/lrelu/style/ToRGB -> HBM).
// The accumulator is double-buffered in TMEM (2 x NT columns) so the epilogue of tile t overlaps the MMAs of t+1.
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"
#include "modconv_epilogue.cuh"

namespace sgr {

template <int NT>
struct ConvCfg {
  static constexpr int kBBytes = NT * kBlockK * 2 * 2;                 // hi + lo planes of one weight stage
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (192 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes;
};

struct TileCoord {
  int n_tile, tx, ty, tb;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int tile) {
  TileCoord t;
  t.n_tile = tile / p.m_tiles;
  int m = tile - t.n_tile * p.m_tiles;
  t.tx = m % p.tiles_x;
  m /= p.tiles_x;
  t.ty = m % p.tiles_y;
  t.tb = m / p.tiles_y;
  return t;
}

template <int NT>
__global__ void __launch_bounds__(256, 1) modconv_kernel(const __grid_constant__ CUtensorMap tmap,
                                                         const ConvKernelParams p) {
  using Cfg = ConvCfg<NT>;
  constexpr int S = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stage_base = smem + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
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
  pdl_launch_dependents();
  pdl_wait();

  const int out_tiles = p.m_tiles * p.n_tiles;
  const int total_tiles = out_tiles * p.ksplit;          // work item = (output tile, K slice); slice fastest
  const int k_iters = p.ntaps * p.kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ks = tile % p.ksplit;
        const TileCoord tc = decode_tile(p, tile / p.ksplit);
        const int x0 = tc.tx * p.bw, y0 = tc.ty * p.bh, b0 = tc.tb * p.bb;
        // single-pass mode reads only the hi plane of the activations and (plane-major slabs, NT > 64) of the weights
        const uint32_t a_bytes = static_cast<uint32_t>(p.rows) * (p.single ? 64u : 128u);
        const uint32_t b_bytes = (p.single && NT > 64) ? Cfg::kBBytes / 2 : Cfg::kBBytes;
        const __nv_bfloat16* wsrc = p.wpacked + static_cast<size_t>(tc.n_tile) * k_iters * (NT * 64);
        const int k0 = ks * k_iters / p.ksplit, k1 = (ks + 1) * k_iters / p.ksplit;
        int tap = k0 / p.kchunks, kc = k0 - tap * p.kchunks;
        for (int k = k0; k < k1; ++k, ++it) {
          const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
          const uint32_t s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = stage_base + s * Cfg::kStageBytes;
          mbar_expect_tx(&full[s], a_bytes + b_bytes);
          tma_load_5d(sa, &tmap, &full[s], (x0 + dx) * 8, y0 + dy, b0, p.tap_chunk[tap] + kc * 4, 0);
          bulk_g2s(sa + kABytes, wsrc + static_cast<size_t>(k) * (NT * 64), b_bytes, &full[s]);
          if (++kc == p.kchunks) { kc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one_sync()) {
      // A stage image: [plane][chunk][row][8] with `rows` dense rows per chunk (the TMA box), rows <= 128
      // (issuing thread = critical resource, see modconv_halo_sm100.cu: stage base descriptors by one multiply-add, the six
      //  MMAs of a stage at base + loop-invariant offsets)
      const uint32_t a_lbo = static_cast<uint32_t>(p.rows) * 16;
      const uint32_t idesc = umma_idesc(p.fmt, kTileM, NT);
      // weight slab: [plane][chunk][n][8] (NT > 64) or [chunk][plane][n][8] (NT <= 64, see pack_weight_kernel)
      constexpr uint32_t b_lbo = NT <= 64 ? 2 * NT * 16 : NT * 16, b_plane = NT <= 64 ? NT * 16 : NT * 64;
      const uint64_t a_ring = umma_desc(smem_u32(stage_base), a_lbo, 128);
      const uint64_t b_ring = umma_desc(smem_u32(stage_base) + kABytes, b_lbo, 128);
      const uint64_t a_j = (2 * a_lbo) >> 4, a_lo_off = (4 * a_lbo) >> 4;
      constexpr uint64_t b_j = (2 * b_lbo) >> 4, b_lo_off = b_plane >> 4;
      uint32_t it = 0, tcount = 0, s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * NT;
        const int ks = tile % p.ksplit;
        const int k0 = ks * k_iters / p.ksplit, k1 = (ks + 1) * k_iters / p.ksplit;
        for (int k = k0; k < k1; ++k, ++it) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t st = static_cast<uint64_t>(s) * (Cfg::kStageBytes >> 4);
          const uint64_t a_s = a_ring + st, b_s = b_ring + st;
          const uint32_t acc0 = k > k0 ? 1u : 0u;
#pragma unroll
          for (int j = 0; j < kBlockK / 16; ++j) {
            const uint64_t a_hi = a_s + j * a_j, a_lo = a_hi + a_lo_off;
            const uint64_t b_hi = b_s + j * b_j, b_lo = b_hi + b_lo_off;
            if (!p.single) {
              umma_bf16(d_tmem, a_lo, b_hi, idesc, j != 0 ? 1u : acc0);
              umma_bf16(d_tmem, a_hi, b_lo, idesc, 1);
              umma_bf16(d_tmem, a_hi, b_hi, idesc, 1);
            } else {
              umma_bf16(d_tmem, a_hi, b_hi, idesc, j != 0 ? 1u : acc0);
            }
          }
          umma_commit(&empty[s]);      // frees the smem stage once these MMAs have read it
          if (++s == S) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull[as]);       // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int ew = warp - 4;                       // == warp % 4: the TMEM lane quarter this warp may access
    const int r = ew * 32 + lane;                  // tile row == pixel
    const int xx = r % p.bw;
    const int yy = (r / p.bw) % p.bh;
    const int bl = r / (p.bw * p.bh);
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const size_t plane_stride = static_cast<size_t>(p.B) * p.cout * p.Hout * p.Wout;   // elements per C8 plane
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int ks = tile % p.ksplit, otile = tile / p.ksplit;
      const TileCoord tc = decode_tile(p, otile);
      const int b = tc.tb * p.bb + bl;
      const int y = tc.ty * p.bh + yy;
      const int x = tc.tx * p.bw + xx;
      const bool valid = r < p.rows && b < p.B && y < p.H && x < p.W;   // partial tiles: TMA zero-fills, stores are masked
      const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < NT; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * NT + c, v);
        tmem_ld_wait();
        if (p.ksplit > 1) {            // raw partial sums of this K slice (splitk_finish_kernel adds and finishes)
          float4* dst = reinterpret_cast<float4*>(p.kpart + ((static_cast<size_t>(ks) * out_tiles + otile) * kTileM + r) * NT + c);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          continue;
        }
        if (valid) epilogue_32cols(p, v, tc.n_tile * NT + c, b, y, x, nw, plane_stride, rgb0, rgb1, rgb2);
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);     // 128 arrivals release the accumulator buffer
      if (valid && p.rgb_coef && p.ksplit == 1) rgb_store(p, tc.n_tile, b, y, x, rgb0, rgb1, rgb2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

template <int NT>
static int launch_nt(const ConvKernelParams& p, const CUtensorMap& tmap, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(modconv_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ConvCfg<NT>::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("modconv: cudaFuncSetAttribute(smem=%d) failed: %s", ConvCfg<NT>::kSmemBytes, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int sms = num_sms();
  if (sms <= 0) {
    set_error("modconv: no CUDA device");
    return 1;
  }
  const int total = p.m_tiles * p.n_tiles * p.ksplit;
  const int grid = std::min(total, sms);
  const cudaError_t le = launch_pdl(modconv_kernel<NT>, dim3(grid), dim3(256), ConvCfg<NT>::kSmemBytes, stream, tmap, p);
  count_launch();
  if (le != cudaSuccess) {
    set_error("modconv_kernel: launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("modconv_kernel") ? 0 : 1;
}

int launch_modconv(const ConvKernelParams& p, const CUtensorMap& tmap, int nt, cudaStream_t stream) {
  switch (nt) {
    case 256: return launch_nt<256>(p, tmap, stream);
    case 128: return launch_nt<128>(p, tmap, stream);
    case 64: return launch_nt<64>(p, tmap, stream);
    case 32: return launch_nt<32>(p, tmap, stream);
    default: set_error("modconv: unsupported column tile %d", nt); return 1;
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_act_tensor_map(CUtensorMap* map, const void* base, int batch, int channels, int h, int w, int bw, int bh,
                        int bb, int planes, int chunk_box, bool wide) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  const cuuint64_t chunk_bytes = static_cast<cuuint64_t>(h) * w * 16;
  // `wide`: the same memory as 8-byte elements (two per 8-channel entry), for boxes of more than 32 entries per row (a box
  // dimension is limited to 256 elements); coordinates along x are then in half entries
  const int epe = wide ? 2 : 8;      // elements per entry
  const cuuint64_t dims[5] = {static_cast<cuuint64_t>(w) * epe, static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(batch),
                              static_cast<cuuint64_t>(channels / 8), 2};
  const cuuint64_t strides[4] = {static_cast<cuuint64_t>(w) * 16, chunk_bytes * (channels / 8), chunk_bytes,
                                 chunk_bytes * (channels / 8) * batch};
  const cuuint32_t box[5] = {static_cast<cuuint32_t>(bw * epe), static_cast<cuuint32_t>(bh), static_cast<cuuint32_t>(bb),
                             static_cast<cuuint32_t>(chunk_box), static_cast<cuuint32_t>(planes)};   // planes == 1: the hi plane only
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, wide ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for B=%d C=%d H=%d W=%d box=%dx%dx%d", static_cast<int>(r), batch,
              channels, h, w, bw, bh, bb);
    return 1;
  }
  return 0;
}

int make_plane_tensor_map(CUtensorMap* map, const float* base, int batch, int channels, int hp, int wp, int cols, int rows,
                          int box_groups, int box_planes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return 1;
  }
  const cuuint64_t group_bytes = static_cast<cuuint64_t>(hp) * wp * 16;
  const cuuint64_t groups = channels / 4;
  const cuuint64_t dims[5] = {static_cast<cuuint64_t>(wp) * 4, static_cast<cuuint64_t>(hp), groups, 4, static_cast<cuuint64_t>(batch)};
  const cuuint64_t strides[4] = {static_cast<cuuint64_t>(wp) * 16, group_bytes, group_bytes * groups, group_bytes * groups * 4};
  const cuuint32_t box[5] = {static_cast<cuuint32_t>(cols * 4), static_cast<cuuint32_t>(rows), static_cast<cuuint32_t>(box_groups),
                             static_cast<cuuint32_t>(box_planes), 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for parity planes B=%d C=%d %dx%d", static_cast<int>(r), batch, channels, hp, wp);
    return 1;
  }
  return 0;
}

bool acc_comp_enabled() {
  static const bool comp = [] { const char* e = getenv("SGR_ACC_COMP"); return !(e && e[0] == '0'); }();
  return comp;
}

// acc_scale for a K range cut into `ksplit` slices: each slice's accumulation chain is 1/ksplit as long
void set_ksplit(ConvKernelParams* p, int ksplit) {
  p->ksplit = ksplit < 1 ? 1 : ksplit;
  p->acc_scale = p->acc_base * (1.f + 1.16e-8f * p->acc_mmas / static_cast<float>(p->ksplit));
}

// Split-K slices: enough to occupy the SMs when a layer has few output tiles, each slice keeping at least
// `min_units_per_slice` K units; 1 (off) without scratch or when the tiles already fill the machine.
int choose_ksplit(const sgr_conv_args* a, int tiles, int k_units, int min_units_per_slice, size_t tile_bytes) {
  static const bool off = [] { const char* e = getenv("SGR_SPLITK"); return e && e[0] == '0'; }();
  if (off || !a->splitk_scratch || tiles <= 0) return 1;
  const int sms = num_sms();
  if (sms <= 0 || 2 * tiles > sms) return 1;
  int s = std::min(16, sms / tiles);
  s = std::min(s, k_units / min_units_per_slice);
  while (s > 1 && static_cast<size_t>(s) * tiles * tile_bytes > a->splitk_scratch_bytes) --s;
  return s < 2 ? 1 : s;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

void tile_box(int h, int w, int* bw, int* bh, int* bb) {
  int pw = 1, ph = 1;
  while (pw < w) pw <<= 1;
  while (ph < h) ph <<= 1;
  *bw = std::min(pw, 16);
  *bh = std::min(ph, kTileM / *bw);
  *bb = kTileM / (*bw * *bh);
}

// Any dense box with bw*bh*bb <= 128 rows is a legal M tile (rows beyond the box are never stored): for grids that are
// not powers of two (the (H+1)x(W+1) parity-plane grid of the scatter up-conv) pick the box that needs the fewest tiles.
void tile_box_search(int batch, int h, int w, int* bw, int* bh, int* bb) {
  // Measured on B200 (tools/gpu_layer_bench.py, SGR_UP_BOX sweep): time follows the tile count as long as a box row is
  // >= 128 B (8 pixels); boxes folding the batch over 2x2-pixel patches are 2.5x slower.  So: fewest tiles among boxes with
  // bw >= min(w, 8), batch folding only when a whole image fits one tile; ties -> wider rows.
  long long best_tiles = -1;
  int best_bw = 0;
  const int min_bw = std::min(w, 8);
  for (int cw = std::min(w, 32); cw >= min_bw; --cw) {
    for (int ch = std::min(h, kTileM / cw); ch >= 1; --ch) {
      int cb = 1;
      if (cw >= w && ch >= h) cb = std::max(1, std::min(batch, kTileM / (cw * ch)));
      const long long tiles = static_cast<long long>((w + cw - 1) / cw) * ((h + ch - 1) / ch) * ((batch + cb - 1) / cb);
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && cw > best_bw)) {
        best_tiles = tiles;
        best_bw = cw;
        *bw = cw; *bh = ch; *bb = cb;
      }
    }
  }
}

// Column tile minimising rounds x per-tile MMA time (cycles per K=16 step: 128/64/48/40 for N=256/128/64/32; the two
// small ones are bound by the shared-memory read of the 128-row A operand, not by the MMA; 80/52 are calibrated).
int choose_nt(int batch, int h, int w, int n_total) {
  int bw, bh, bb;
  tile_box(h, w, &bw, &bh, &bb);
  const long long m_tiles = static_cast<long long>((w + bw - 1) / bw) * ((h + bh - 1) / bh) * ((batch + bb - 1) / bb);
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  int best = pick_nt(n_total);
  double best_cost = 1e300;
  for (int nt = 256; nt >= 64; nt >>= 1) {
    if (nt > n_total || n_total % nt) continue;
    const double per_tile = nt == 256 ? 128 : nt == 128 ? 80 : 52;   // measured: N<=128 is operand-read bound
    const long long tiles = m_tiles * (n_total / nt);
    const double cost = static_cast<double>((tiles + sms - 1) / sms) * per_tile;
    if (cost < best_cost * 0.95) {       // prefer the larger tile unless clearly better
      best_cost = cost;
      best = nt;
    }
  }
  return best;
}

int conv_fill_params(const sgr_conv_args* a, ConvKernelParams* p, int* nt) {
  if (!a || !a->x_c8 || !a->w_packed) {
    set_error("modconv: null input");
    return 1;
  }
  if (a->batch <= 0 || a->h_in < 1 || a->w_in < 1 || a->h_in > 16384 || a->w_in > 16384) {
    set_error("modconv: unsupported geometry B=%d H=%d W=%d", a->batch, a->h_in, a->w_in);
    return 1;
  }
  if (a->cin % kBlockK != 0 || !is_pow2(a->cout) || a->cout < 32 || (a->ksize != 3 && a->ksize != 1)) {
    set_error("modconv: unsupported channels cin=%d cout=%d k=%d (cin %% 32 == 0, cout power of two >= 32)", a->cin,
              a->cout, a->ksize);
    return 1;
  }
  if (a->up && a->ksize != 3) {
    set_error("modconv: up requires ksize 3");
    return 1;
  }
  if (a->up < 0 || a->up > 3 || (a->up == 2 && (!a->t_scratch || !a->fir || !a->demod))) {
    set_error("modconv: up must be 0, 1 (polyphase), 2 (scatter + FIR; needs t_scratch, fir and demod) or 3 (gather adjoint)");
    return 1;
  }
  if (a->noise && !a->noise_weight) {
    set_error("modconv: noise without noise_weight");
    return 1;
  }
  if (a->rgb_coef && (!a->rgb_partial || a->up)) {
    set_error("modconv: fused ToRGB needs rgb_partial and a non-upsampling layer");
    return 1;
  }
  p->B = a->batch;
  p->mode = a->up == 3 ? 0 : a->up;
  { const char* e = getenv("SGR_DEBUG"); p->debug = e ? atoi(e) : 0; }
  if (a->up == 2) {               // tiles walk the (H+1) x (W+1) parity-plane grid
    p->H = a->h_in + 1;
    p->W = a->w_in + 1;
    tile_box_search(a->batch, p->H, p->W, &p->bw, &p->bh, &p->bb);
    if (const char* e = getenv("SGR_UP_BOX")) {        // experiments: "bw,bh,bb"
      int ebw = 0, ebh = 0, ebb = 0;
      if (sscanf(e, "%d,%d,%d", &ebw, &ebh, &ebb) == 3 && ebw >= 1 && ebw <= 32 && ebh >= 1 && ebb >= 1 &&
          ebw * ebh * ebb <= kTileM) {
        p->bw = ebw; p->bh = ebh; p->bb = ebb;
      }
    }
    p->halo = 0;
    p->box_rows = 0;
    static const bool halo_off = [] { const char* e = getenv("SGR_UP_HALO"); return e && e[0] == '0'; }();
    if (!halo_off && p->bb == 1 && !getenv("SGR_UP_BOX")) {
      // Wrapped-halo tiles: one dense box of hbw x (vr + 1) pixels starting one pixel up / left of the tile; M row r reads box
      // entry hbw + 1 + r - (a * hbw + b) for shift (a, b), so the entries of box column 0 produce garbage (their left
      // neighbour wraps to the previous row) and are discarded: hbw - 1 valid columns x vr valid rows, vr * hbw <= 129.
      // A traffic L2 -> shared memory drops from 4 boxes per channel block to 1.1 (the kernel was bound by exactly that).
      // The box has 144 or 140 entries (hbw x rows): a fixed chunk stride makes every MMA operand of the kernel base +
      // compile-time constant (the issuing thread is the critical resource); one kernel instantiation per entry count.
      long long best = -1;
      int best_w = 0, best_vr = 0, best_rows = 0;
      static const int shapes[6][2] = {{24, 6}, {18, 8}, {16, 9}, {14, 10}, {12, 12}, {9, 16}};
      for (const auto& sh : shapes) {
        const int hbw = sh[0];
        const int vr = std::min(sh[1] - 1, 129 / hbw);
        const long long tiles = static_cast<long long>((p->W + hbw - 2) / (hbw - 1)) * ((p->H + vr - 1) / vr);
        if (best < 0 || tiles < best) {
          best = tiles;
          best_w = hbw;
          best_vr = vr;
          best_rows = sh[1];
        }
      }
      // measured (tools/gpu_layer_bench.py, B=32): 129^2 grid 150 vs 143 tiles -5 %, 65^2 39 vs 42 -1 %, 33^2 10 vs 9 +6 %:
      // the single box wins unless it costs more than ~5 % extra tiles
      const long long legacy = static_cast<long long>((p->W + p->bw - 1) / p->bw) * ((p->H + p->bh - 1) / p->bh);
      if (best_w > 0 && best * 100 <= legacy * 106) {
        p->halo = 1;
        p->bw = best_w;
        p->bh = best_vr;                // valid rows = tile stride in y
        p->box_rows = best_rows;        // rows of the TMA box
        p->bb = 1;
      }
    }
    // Linear tiles (modconv_scatter_sm100.cu) where the wrapped-halo box lost: runs of 128 consecutive entries of the grid
    // padded to pitch P = W + 1, one P x R box per channel block.  Instantiated for the 33^2 and 65^2 grids (204 / 264 entries).
    static const bool lin_off = [] { const char* e = getenv("SGR_UP_LINEAR"); return e && e[0] == '0'; }();
    if (!lin_off && !halo_off && !p->halo && p->bb == 1 && !a->single_pass && !getenv("SGR_UP_BOX")) {
      const int P = p->W + 1, R = (2 * P - 1 + kTileM) / P + 1;
      if (P * R == 204 || P * R == 264) {
        p->halo = 2;
        p->bw = P;
        p->bh = 1;                      // unused: the tiles are runs of entries, not boxes
        p->box_rows = R;
      }
    }
  } else {
    p->halo = 0;
    p->box_rows = 0;
    p->H = a->h_in;
    p->W = a->w_in;
    tile_box(a->h_in, a->w_in, &p->bw, &p->bh, &p->bb);
  }
  p->rows = p->bw * p->bh * p->bb;
  p->tiles_x = (p->W + p->bw - 1) / p->bw;
  if (p->halo) {                  // rows = store slots of a tile (valid positions)
    p->rows = (p->bw - 1) * p->bh;
    p->tiles_x = (p->W + p->bw - 2) / (p->bw - 1);
  }
  p->tiles_y = (p->H + p->bh - 1) / p->bh;
  if (p->halo == 2) {             // entries P + 1 .. (H + 1) P - 1 of the padded grid in runs of 128
    p->rows = kTileM;
    p->tiles_x = 1;
    p->tiles_y = (p->H * p->bw - 1 + kTileM - 1) / kTileM;
  }
  p->tiles_b = (a->batch + p->bb - 1) / p->bb;
  p->m_tiles = p->tiles_x * p->tiles_y * p->tiles_b;
  const int n_total = a->cout * ((a->up == 1 || a->up == 2) ? 4 : 1);
  if (a->up == 2) {
    *nt = up2_nt(a->cout);        // [oe|ee|eo|oo] blocks of NT/4 channels; fixed by the packed layout
    if (a->column_tile > 0 && a->column_tile != *nt) {
      set_error("modconv: scatter up-conv uses column tile %d for cout %d (got %d)", *nt, a->cout, a->column_tile);
      return 1;
    }
  } else {
    *nt = a->column_tile > 0 ? a->column_tile : pick_nt(n_total);
  }
  if (*nt > n_total || n_total % *nt != 0 || (*nt != 32 && *nt != 64 && *nt != 128 && *nt != 256)) {
    set_error("modconv: column tile %d does not fit %d columns", *nt, n_total);
    return 1;
  }
  p->n_tiles = n_total / *nt;
  p->kchunks = a->cin / kBlockK;
  p->ntaps = a->up == 2 ? 4 : a->ksize * a->ksize;
  for (int t = 0; t < 9; ++t) {
    p->tap_dy[t] = static_cast<signed char>(p->ntaps == 9 ? t / 3 - 1 : 0);
    p->tap_dx[t] = static_cast<signed char>(p->ntaps == 9 ? t % 3 - 1 : 0);
    p->tap_chunk[t] = 0;
  }
  if (a->up == 3) {
    // gather adjoint of the scatter up-conv: x_c8 holds the four parity planes [ee|eo|oe|oo] of the FIR^T-filtered
    // gradient stacked as 4*cin channels on the (h_in+1) x (w_in+1) grid; tap t = (plane, shift (a,b))
    static const int plane[9] = {0, 0, 0, 0, 1, 1, 2, 2, 3};
    static const int sa[9] = {0, 0, 1, 1, 0, 1, 0, 0, 0};
    static const int sb[9] = {0, 1, 0, 1, 0, 0, 0, 1, 0};
    for (int t = 0; t < 9; ++t) {
      p->tap_dy[t] = static_cast<signed char>(sa[t]);
      p->tap_dx[t] = static_cast<signed char>(sb[t]);
      p->tap_chunk[t] = plane[t] * (a->cin / 8);
    }
  }
  p->cout = a->cout;
  p->up = a->up == 1 ? 1 : 0;
  p->t_out = a->t_scratch;
  p->Hout = (a->up == 1 || a->up == 2) ? 2 * a->h_in : a->h_in;
  p->Wout = (a->up == 1 || a->up == 2) ? 2 * a->w_in : a->w_in;
  p->act = a->act;
  p->act_gain = a->act_gain;
  if ((a->operand_format != SGR_FMT_BF16 && a->operand_format != SGR_FMT_FP16) ||
      (a->out_format != SGR_FMT_BF16 && a->out_format != SGR_FMT_FP16)) {
    set_error("modconv: unknown operand format");
    return 1;
  }
  p->fmt = a->operand_format;
  p->single = a->single_pass ? 1 : 0;
  p->single_out = a->single_pass ? 1 : 0;
  p->acc_base = 1.f / (act_scale(a->operand_format) * w_scale(a->operand_format));
  p->acc_scale = p->acc_base;
  p->ksplit = 1;
  p->halo_mt = 0;
  p->nt = *nt;
  p->kpart = static_cast<float*>(a->splitk_scratch);
  // The tensor core's fp32 accumulate truncates: measured on B200 the result shrinks by ~1.16e-8 per MMA accumulated into
  // the same TMEM cell (tools/gpu_debug.py "mean signed rel err": -1.0e-5 at 864 MMAs, -1.6e-6 at 108).  Undo the mean.
  const bool comp = acc_comp_enabled();
  // (scatter up-conv: the parity planes see 4/2/2/1 taps; 9/4 on average over the 4 planes feeding each output)
  const float per_product = a->single_pass ? 1.f : 3.f;
  const float mmas = a->up == 2 ? per_product * 2.25f * (a->cin / 16) : per_product * a->ksize * a->ksize * (a->cin / 16);
  p->acc_mmas = comp ? mmas : 0.f;
  set_ksplit(p, 1);
  p->out_fmt = a->out_format;
  p->out_scale = act_scale(a->out_format);
  p->wpacked = static_cast<const __nv_bfloat16*>(a->w_packed);
  p->demod = a->demod;
  p->bias = a->bias;
  p->noise = a->noise;
  p->noise_w = a->noise_weight;
  p->noise_bstride = a->noise_batch_stride;
  p->s2 = a->s2;
  p->out_c8 = static_cast<__nv_bfloat16*>(a->out_c8);
  p->out_f32 = a->out_f32;
  p->rgb_coef = a->rgb_coef;
  p->rgb_part = a->rgb_partial;
  return 0;
}

}  // namespace sgr

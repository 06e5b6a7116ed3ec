// 3x3 modulated convolution on the sm_100a tensor cores with a RESIDENT halo tile (same math and epilogue as
// modconv_sm100.cu; replaces the same reference code: model.py:232-287,331-337).
//
// modconv_sm100.cu loads the 128-pixel A tile once per tap (9 shifted TMA boxes per 32-channel block) and one weight
// slab per 128 pixels.  On the wide, low-channel layers that makes the kernel L2 -> shared-memory bound.  Here:
//   * one TMA box per 32-channel block brings the pixel tile WITH its one-pixel halo, (8*MT + 2) x 18 pixels, and all nine
//     taps read it in place: in the no-swizzle K-major layout a pixel is one 16-byte row, the eight pixels of an image row
//     are one core matrix (8 rows x 16 B, contiguous), image rows are SBO = (8*MT+2)*16 bytes apart and the 8-channel
//     chunks LBO = box bytes apart, so tap (dy,dx) is just the descriptor start address + (dy*(8*MT+2) + dx)*16;
//   * MT (1 or 2) horizontally adjacent 8x16-pixel tiles share every weight slab (MT accumulators of NT columns, double
//     buffered in the 512 TMEM columns), halving the weight traffic of the NT <= 128 layers.
// A-operand traffic drops from 9x to (8MT+2)*18 / (8MT*16) = 1.4x / 1.27x of the tile.
//
// Warp roles as in modconv_sm100.cu: warp0 TMA producer (A ring of 32-channel blocks + B ring of (block, tap) slabs),
// warp1 MMA issuer, warp2 TMEM allocator, warps4-7 epilogue.
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"
#include "modconv_epilogue.cuh"
#include "fir_producer.cuh"

namespace sgr {

constexpr int kFirRows = 12;      // plane rows m_first-1 .. m_first+10 feed the 18 halo-tile rows (fir_producer.cuh)
#ifndef SGR_FIR_WARPS
#define SGR_FIR_WARPS 8
#endif
constexpr int kFirWarps = SGR_FIR_WARPS;      // fused mode: warps 8.. build the A halo tiles from the parity planes (8 or 12)
constexpr int kFirParts = kFirWarps / 4;      // the 10 plane rows of a tile window are cut into 2 x 5 or 4 + 3 + 3
constexpr int kFusedThreads = 256 + 32 * kFirWarps;

template <int NT, int MT, bool FUSED = false>
struct HaloCfg {
  static constexpr int kTileW = 8 * MT, kTileH = 16;
  static constexpr int kHW = kTileW + 2, kHH = kTileH + 2;
  static constexpr int kChunkBytes = kHW * kHH * 16;          // one 8-channel chunk of the halo box
  static constexpr int kABytes = kChunkBytes * 4 * 2;         // 32 channels x (hi, lo)
  static constexpr int kBBytes = NT * kBlockK * 2 * 2;        // one (block, tap) weight slab, hi + lo
  // fused mode: two TMA-fed stages of parity planes, 16 channels each: [plane 4][group 4][row 12][col kPC][4 floats]
  // plane-window columns: kHW / 2 + 3 are needed; the 18-pixel tile takes 13 so that the three row sub-groups of a producer
  // warp (rows 384 B apart at 12 columns = the same banks) fall into different banks (ncu: 50 % of the wavefronts conflicted)
  // (the 10-pixel tile: five sub-groups of six lanes, rows one apart: 14 columns = 224 B pitch)
  static constexpr int kPC = kHW == 18 ? 13 : 14;
  static constexpr int kPlaneStageBytes = kPC * 16 * kFirRows * 4 * 4;
  static constexpr int kPlaneBytes = FUSED ? 2 * kPlaneStageBytes : 0;
  static constexpr int kAStages = FUSED ? 2 : 3;
  static constexpr int kBRaw = (226 * 1024 - 1024 - kAStages * kABytes - kPlaneBytes) / kBBytes;
  static constexpr int kBStages = kBRaw > 8 ? 8 : kBRaw;
  static constexpr int kSmemBytes = 1024 + kAStages * kABytes + kBStages * kBBytes + kPlaneBytes;
  static_assert(kABytes % 128 == 0, "TMA destination alignment");
  static_assert(kBStages >= 3, "weight ring too shallow");
  // NT <= 64: an N = 64 MMA costs as much as N = 128 (the 128-row A operand read bounds it), so the three products of the
  // hi/lo scheme are issued as A_hi x [W_hi | W_lo] (one N = 2 NT MMA into 2 NT columns) + A_lo x W_hi (N = NT); the
  // epilogue adds the two column halves.
  static constexpr bool kConcat = NT <= 64;
  static constexpr int kAccCols = kConcat ? 2 * NT : NT;      // TMEM columns per sub-tile
  static_assert(2 * MT * kAccCols <= 512, "TMEM columns");
  static_assert(kHH == 18, "fir_produce_chunk walks 10 plane rows in two halves of five");
};

template <int NT, int MT, bool FUSED>
__global__ void __launch_bounds__(FUSED ? kFusedThreads : 256, 1) modconv_halo_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                            const __grid_constant__ CUtensorMap tmap_planes,
                                                                            const ConvKernelParams p, const FusedFirParams f) {
  using Cfg = HaloCfg<NT, MT, FUSED>;
  constexpr int AS = Cfg::kAStages, BS = Cfg::kBStages;
#ifndef SGR_PACED_MODE
#define SGR_PACED_MODE 0
#endif
  constexpr bool kPacedIssue = SGR_PACED_MODE == 1 ? false : (FUSED && NT <= 128);      // see the two MMA issuers below
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_empty = a_full + AS;
  uint64_t* b_full = a_empty + AS;
  uint64_t* b_empty = b_full + BS;
  uint64_t* tfull = b_empty + BS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* fir_sc = reinterpret_cast<float*>(tmem_slot + 2);      // fused mode: separable FIR taps (gy | gx)
  uint64_t* p_full = reinterpret_cast<uint64_t*>(fir_sc + 8);   // fused mode: parity-plane stages
  uint64_t* p_empty = p_full + 2;
  uint8_t* a_base = smem + 1024;
  uint8_t* b_base = a_base + AS * Cfg::kABytes;
  uint8_t* plane_base = b_base + BS * Cfg::kBBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    if (FUSED) tma_prefetch_desc(&tmap_planes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AS; ++i) {
      mbar_init(&a_full[i], FUSED ? kFirWarps : 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < BS; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
      if (FUSED) {
        mbar_init(&p_full[i], 1);
        mbar_init(&p_empty[i], kFirWarps);
      }
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (FUSED && warp == 3 && lane < 4) {      // gy (vertical taps, flipped), gx (horizontal, flipped, / tap sum): up_finish_kernel
    const int a = lane;
    float rs = 0.f, cs = 0.f, tot = 0.f;
    for (int i = 0; i < 4; ++i) {
      rs += __ldg(f.fir + (3 - a) * 4 + i);
      cs += __ldg(f.fir + i * 4 + (3 - a));
      for (int j = 0; j < 4; ++j) tot += __ldg(f.fir + i * 4 + j);
    }
    fir_sc[a] = rs;
    fir_sc[4 + a] = cs / tot;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();      // the next kernel may set itself up on SMs this grid has left
  pdl_wait();                   // nothing above reads what the previous kernel wrote

  const int out_tiles = p.m_tiles * p.n_tiles;
  const int total_tiles = out_tiles * p.ksplit;          // work item = (output tile, K slice of channel blocks)

  // (setmaxnreg re-dealing of the 64 K registers between the service / epilogue / producer warp groups was tried: ptxas keeps
  //  the 128-register cap of the 512-thread launch bound in the `inc` regions and spills MORE in the `dec` region)

  if (FUSED && warp >= 8) {
    // ------------------------------------------------------------------ FIR producers (fused mode; ksplit == 1)
    // every 32-channel block is built in two phases of 16 channels (plane stage h = phase): all eight warps work on the same
    // stage — 4-channel group = pw / kFirParts, part of the ten plane rows = pw % kFirParts — while TMA refills the other one
    const int pw = warp - 8, gq = pw / kFirParts, part = pw % kFirParts;
    const int row0 = kFirParts == 2 ? part * 5 : (part == 0 ? 0 : 1 + part * 3);      // 0,5 | 0,4,7
    const int nrows = kFirParts == 2 ? 5 : (part == 0 ? 4 : 3);
    constexpr int kNrMax = kFirParts == 2 ? 5 : 4;
    uint32_t ai = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile / p.m_tiles;
      int m = tile - n_tile * p.m_tiles;
      const int tx = m % p.tiles_x;
      m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = m / p.tiles_y;
      const int X0 = tx * Cfg::kTileW - 1, Y0 = ty * Cfg::kTileH - 1;
      const int m_first = (Y0 - 1) >> 1;
      for (int kb = 0; kb < p.kchunks; ++kb, ++ai) {
        const uint32_t as = ai % AS;
        mbar_wait(&a_empty[as], ((ai / AS) & 1) ^ 1);      // (one polling lane + __syncwarp measured 2x SLOWER)
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&p_full[h], ai & 1);
          if (!(p.debug & 512)) {                                          // experiment 512: handshakes only, no FIR work
            const uint8_t* stage = plane_base + h * Cfg::kPlaneStageBytes;
            // group gq of this stage = channels kb*32 + h*16 + gq*4 ..: chunk h*2 + (gq >> 1), upper half when gq is odd
            uint8_t* dst8 = a_base + as * Cfg::kABytes + (h * 2 + (gq >> 1)) * Cfg::kChunkBytes + (gq & 1) * 8;
            const int group = kb * 8 + h * 4 + gq;
            if (p.fmt == kFmtBF16)
              fir_produce_group_smem<Cfg::kHW, Cfg::kHH, kFmtBF16, Cfg::kPC, kNrMax>(f, fir_sc, b, group, Y0, X0, stage, gq, m_first + row0,
                                                                                     1 + row0, nrows, dst8, Cfg::kChunkBytes * 4, lane);
            else
              fir_produce_group_smem<Cfg::kHW, Cfg::kHH, kFmtFP16, Cfg::kPC, kNrMax>(f, fir_sc, b, group, Y0, X0, stage, gq, m_first + row0,
                                                                                     1 + row0, nrows, dst8, Cfg::kChunkBytes * 4, lane);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_empty[h]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
      }
    }
  } else if (FUSED && warp == 3) {
    // ------------------------------------------------------------------ parity-plane TMA issuer (fused mode)
    if (elect_one_sync()) {
      uint32_t ai = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile / p.m_tiles;
        int m = tile - n_tile * p.m_tiles;
        const int tx = m % p.tiles_x;
        m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int b = m / p.tiles_y;
        const int n_first = (tx * Cfg::kTileW - 2) >> 1, m_first = (ty * Cfg::kTileH - 2) >> 1;
        for (int kb = 0; kb < p.kchunks; ++kb, ++ai) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&p_empty[h], (ai & 1) ^ 1);
            mbar_expect_tx(&p_full[h], Cfg::kPlaneStageBytes);
            tma_load_5d(plane_base + h * Cfg::kPlaneStageBytes, &tmap_planes, &p_full[h], (n_first - 1) * 4, m_first - 1,
                        kb * 8 + h * 4, 0, b);
          }
        }
      }
    }
  } else

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      uint32_t ai = 0, bi = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ks = tile % p.ksplit, otile = tile / p.ksplit;
        const int n_tile = otile / p.m_tiles;
        int m = otile - n_tile * p.m_tiles;
        const int tx = m % p.tiles_x;
        m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int b = m / p.tiles_y;
        const int x0 = tx * Cfg::kTileW - 1, y0 = ty * Cfg::kTileH - 1;
        const __nv_bfloat16* wsrc = p.wpacked + static_cast<size_t>(n_tile) * 9 * p.kchunks * (NT * 64);
        // single-pass mode: hi plane of the halo box; hi half of plane-major weight slabs (NT > 64)
        const uint32_t a_bytes = p.single ? Cfg::kABytes / 2 : Cfg::kABytes;
        const uint32_t b_bytes = (p.single && !Cfg::kConcat) ? Cfg::kBBytes / 2 : Cfg::kBBytes;
        const int kb0 = ks * p.kchunks / p.ksplit, kb1 = (ks + 1) * p.kchunks / p.ksplit;
        for (int kb = kb0; kb < kb1; ++kb, ++ai) {
          const uint32_t as = ai % AS;
          if (!FUSED && (!(p.debug & 16) || ai < AS)) {     // experiment 16: A ring loaded once; fused: FIR warps fill it
            mbar_wait(&a_empty[as], ((ai / AS) & 1) ^ 1);
            mbar_expect_tx(&a_full[as], a_bytes);
            tma_load_5d(a_base + as * Cfg::kABytes, &tmap, &a_full[as], x0 * 8, y0, b, kb * 4, 0);
          }
          for (int tap = 0; tap < 9; ++tap, ++bi) {
            if ((p.debug & 4) && bi >= BS) continue;   // experiment 4: weight ring loaded once
            const uint32_t bs = bi % BS;
            mbar_wait(&b_empty[bs], ((bi / BS) & 1) ^ 1);
            mbar_expect_tx(&b_full[bs], b_bytes);
            bulk_g2s(b_base + bs * Cfg::kBBytes, wsrc + static_cast<size_t>(tap * p.kchunks + kb) * (NT * 64), b_bytes,
                     &b_full[bs]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    // The issuing thread is the critical resource: its instructions run back to back at ~8 cycles each (uniform datapath),
    // and tcgen05.mma executes in 64 (N = 128) .. 128 (N = 256) cycles (tools/microbench/mma_rate.cu: a stream of precomputed
    // descriptors reaches exactly that), so an MMA may cost at most ~8 issue instructions.  The first version rebuilt both
    // shared-memory descriptors per MMA (~20 instructions: 170-180 cycles per N = 256 MMA, 100 per N = 128 MMA measured with
    // all loads switched off).  Here: ring base descriptors once, one per stage / slab by a multiply-add, the nine taps fully
    // unrolled so that every MMA's operands are base + compile-time constant.
    if (!kPacedIssue && elect_one_sync()) {
      const uint32_t idesc = umma_idesc(p.fmt, kTileM, NT);
      const uint32_t idesc_cat = umma_idesc(p.fmt, kTileM, Cfg::kConcat ? 2 * NT : NT);
      constexpr uint32_t kALbo = Cfg::kChunkBytes, kASbo = Cfg::kHW * 16, kAPlane = Cfg::kChunkBytes * 4;
      // weight slab: [plane][chunk][n][8] (NT > 64) or [chunk][plane][n][8] (NT <= 64)
      constexpr uint32_t kBLbo = Cfg::kConcat ? 2 * NT * 16 : NT * 16, kBPlane = Cfg::kConcat ? NT * 16 : NT * 64;
      // offsets in descriptor units (16 B)
      constexpr uint64_t kAJ = (2 * kALbo) >> 4, kALo = kAPlane >> 4, kBJ = (2 * kBLbo) >> 4, kBLo = kBPlane >> 4;
      const uint64_t a_ring = umma_desc(smem_u32(a_base), kALbo, kASbo);
      const uint64_t b_ring = umma_desc(smem_u32(b_base), kBLbo, 128);
      const bool acc_always = (p.debug & 1024) != 0;
      const uint32_t idesc_lo = (p.debug & 64) ? idesc_cat : idesc;
      // 0: concat scheme (NT <= 64), 1: sub-tile interleaved three products (NT >= 128), 2: single pass / experiment orders
      const int variant = p.single ? 2 : (Cfg::kConcat ? ((p.debug & 32) ? 2 : 0) : ((p.debug & 128) ? 2 : 1));
      uint32_t ai = 0, bi = 0, tcount = 0;
      uint32_t as = 0, aphase = 0, bs = 0, bphase = 0;          // ring slot + phase of the next stage / slab (no divisions)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * Cfg::kAccCols);
        const int ks = tile % p.ksplit;
        const int kb0 = ks * p.kchunks / p.ksplit, kb1 = (ks + 1) * p.kchunks / p.ksplit;
        for (int kb = kb0; kb < kb1; ++kb, ++ai) {
          if (!(p.debug & 16) || ai < AS) mbar_wait(&a_full[as], aphase);
          tc_fence_after();
          const uint64_t a_d = a_ring + static_cast<uint64_t>(as) * (Cfg::kABytes >> 4);
          const uint32_t first = (kb > kb0 || acc_always) ? 1u : 0u;      // accumulate flag of the very first MMA of the tile
          // one kernel row (dy) per trip, its three taps unrolled: the variant in use stays a compact, contiguous piece of
          // code (nine unrolled taps x four variants overflowed the instruction cache: the 64-channel layer ran 26 % slower)
#pragma unroll 1
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t a_row = a_d + static_cast<uint64_t>(dy * Cfg::kHW);
            const uint32_t first_row = dy == 0 ? first : 1u;
            auto slab_wait = [&]() {
              if (!(p.debug & 4) || bi < BS) mbar_wait(&b_full[bs], bphase);
              tc_fence_after();
              return b_ring + static_cast<uint64_t>(bs) * (Cfg::kBBytes >> 4);
            };
            auto slab_done = [&]() {
              if (!(p.debug & 256)) umma_commit(&b_empty[bs]);     // weight slab consumed (experiment 256 with 4|16: no ring commits)
              ++bi;
              if (++bs == BS) {
                bs = 0;
                bphase ^= 1;
              }
            };
            if (variant == 0) {
              // NT <= 64: A_hi x [W_hi | W_lo] (N = 2 NT) for both sub-tiles, then A_lo x W_hi (N = NT): MMAs of one shape are
              // issued together — alternating the two shapes on the same accumulator columns costs ~25 ns per switch
              // (0.64 -> 0.52 ms on the 64 -> 64 layer at 256^2; debug 32 = the old alternating order)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint64_t b_d = slab_wait(), a_t = a_row + dx;
#pragma unroll
                for (int j = 0; j < kBlockK / 16; ++j)
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt)
                    umma_bf16(d_tmem + mt * Cfg::kAccCols, a_t + (mt * 8 + j * kAJ), b_d + j * kBJ, idesc_cat,
                              (dx | j) != 0 ? 1u : first_row);                                 // hi*hi | hi*lo
#pragma unroll
                for (int j = 0; j < kBlockK / 16; ++j)
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt)
                    umma_bf16(d_tmem + mt * Cfg::kAccCols, a_t + (mt * 8 + j * kAJ + kALo), b_d + j * kBJ, idesc_lo, 1);   // lo*hi
                slab_done();
              }
            } else if (variant == 1) {
              // NT >= 128: consecutive MMAs alternate between the sub-tiles' accumulators (MT = 2): an MMA that accumulates
              // into the TMEM columns of its predecessor waits for it; debug 128 = three products of a sub-tile back to back
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint64_t b_d = slab_wait(), a_t = a_row + dx;
#pragma unroll
                for (int j = 0; j < kBlockK / 16; ++j)
#pragma unroll
                  for (int prod = 0; prod < 3; ++prod)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                      umma_bf16(d_tmem + mt * NT, a_t + (mt * 8 + j * kAJ + (prod == 0 ? kALo : 0)),
                                b_d + (j * kBJ + (prod == 1 ? kBLo : 0)), idesc,
                                (dx | j | prod) != 0 ? 1u : first_row);                        // lo*hi, hi*lo, hi*hi
                slab_done();
              }
            } else {
              // single-pass mode and the experiment orders
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint64_t b_d = slab_wait(), a_t = a_row + dx;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                  for (int j = 0; j < kBlockK / 16; ++j) {
                    const uint64_t a_hi = a_t + (mt * 8 + j * kAJ), a_lo = a_hi + kALo;
                    const uint64_t b_hi = b_d + j * kBJ;
                    const uint32_t acc0 = (dx | j) != 0 ? 1u : first_row;
                    if (p.single) {
                      umma_bf16(d_tmem + mt * Cfg::kAccCols, a_hi, b_hi, idesc, acc0);
                    } else if (Cfg::kConcat) {
                      umma_bf16(d_tmem + mt * Cfg::kAccCols, a_hi, b_hi, idesc_cat, acc0);     // hi*hi | hi*lo
                      umma_bf16(d_tmem + mt * Cfg::kAccCols, a_lo, b_hi, idesc_lo, 1);         // lo*hi (64: + lo*lo)
                    } else {
                      umma_bf16(d_tmem + mt * NT, a_lo, b_hi, idesc, acc0);
                      umma_bf16(d_tmem + mt * NT, a_hi, b_hi + kBLo, idesc, 1);
                      umma_bf16(d_tmem + mt * NT, a_hi, b_hi, idesc, 1);
                    }
                  }
                }
                slab_done();
              }
            }
          }
          if (!(p.debug & 256)) umma_commit(&a_empty[as]);       // halo block consumed by all nine taps
          if (++as == AS) {
            as = 0;
            aphase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);          // accumulators complete
      }
    }
    // ---- paced issuer (FUSED, NT <= 128): descriptors rebuilt per MMA, taps not unrolled.  The fused low-channel kernels are
    // bound by shared-memory bandwidth (MMA operand reads + FIR producers ~0.85 wavefronts / cycle); measured on the same B200,
    // the streamlined issuer above makes them SLOWER (64 -> 64 @ 256^2: 0.69 -> 0.80 ms, 128 -> 128 @ 128^2: 0.49 -> 0.50 ms):
    // bursts of operand reads starve the producer warps.  Everywhere else it wins (512 -> 512 @ 32^2: 0.40 -> 0.36 ms).
    if (kPacedIssue && (SGR_PACED_MODE == 2 ? elect_one_sync() : lane == 0)) {
      const uint32_t idesc = umma_idesc(p.fmt, kTileM, NT);
      const uint32_t idesc_cat = umma_idesc(p.fmt, kTileM, Cfg::kConcat ? 2 * NT : NT);
      constexpr uint32_t kALbo = Cfg::kChunkBytes, kASbo = Cfg::kHW * 16, kAPlane = Cfg::kChunkBytes * 4;
      // weight slab: [plane][chunk][n][8] (NT > 64) or [chunk][plane][n][8] (NT <= 64)
      constexpr uint32_t kBLbo = Cfg::kConcat ? 2 * NT * 16 : NT * 16, kBPlane = Cfg::kConcat ? NT * 16 : NT * 64;
      uint32_t ai = 0, bi = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * Cfg::kAccCols);
        const int ks = tile % p.ksplit;
        const int kb0 = ks * p.kchunks / p.ksplit, kb1 = (ks + 1) * p.kchunks / p.ksplit;
        for (int kb = kb0; kb < kb1; ++kb, ++ai) {
          const uint32_t as = ai % AS;
          if (!(p.debug & 16) || ai < AS) mbar_wait(&a_full[as], (ai / AS) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_base + as * Cfg::kABytes);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap, ++bi) {
            const uint32_t bs = bi % BS;
            if (!(p.debug & 4) || bi < BS) mbar_wait(&b_full[bs], (bi / BS) & 1);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(b_base + bs * Cfg::kBBytes);
            const uint32_t tap_off = static_cast<uint32_t>((tap / 3) * Cfg::kHW + (tap % 3)) * 16;
            if (Cfg::kConcat && !p.single && !(p.debug & 32)) {
              // MMAs of one shape are issued together: alternating the N = 2 NT and N = NT instructions on the same
              // accumulator columns costs ~25 ns per switch (measured: 0.64 -> 0.52 ms on the 64 -> 64 layer at 256^2 with
              // the pairs two MMAs apart, debug 32 = the old alternating order)
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j) {
                const uint64_t b_hi = umma_desc(b_addr + j * 2 * kBLbo, kBLbo, 128);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t a_off = a_addr + tap_off + mt * 128 + j * 2 * kALbo;
                  umma_bf16(d_tmem + mt * Cfg::kAccCols, umma_desc(a_off, kALbo, kASbo), b_hi, idesc_cat,
                            (kb > kb0 || (tap | j) != 0));                                   // hi*hi | hi*lo
                }
              }
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j) {
                const uint64_t b_hi = umma_desc(b_addr + j * 2 * kBLbo, kBLbo, 128);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t a_off = a_addr + tap_off + mt * 128 + j * 2 * kALbo;
                  umma_bf16(d_tmem + mt * Cfg::kAccCols, umma_desc(a_off + kAPlane, kALbo, kASbo), b_hi,
                            (p.debug & 64) ? idesc_cat : idesc, 1);                           // lo*hi
                }
              }
            } else if (!Cfg::kConcat && !p.single && MT > 1 && !(p.debug & 128)) {
              // consecutive MMAs alternate between the sub-tiles' accumulators: an MMA that accumulates into the TMEM columns
              // of its predecessor waits ~43 cycles for it (tools/microbench/mma_rate.cu: N=128 107 cycles in one chain, 64 when
              // two chains alternate); debug 128 = the old order (three products of one sub-tile back to back)
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j) {
                const uint64_t b_hi = umma_desc(b_addr + j * 2 * kBLbo, kBLbo, 128);
                const uint64_t b_lo = umma_desc(b_addr + kBPlane + j * 2 * kBLbo, kBLbo, 128);
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t a_off = a_addr + tap_off + mt * 128 + j * 2 * kALbo + (prod == 0 ? kAPlane : 0);
                    umma_bf16(d_tmem + mt * NT, umma_desc(a_off, kALbo, kASbo), prod == 1 ? b_lo : b_hi, idesc,
                              prod != 0 || kb > kb0 || (tap | j) != 0);          // lo*hi, hi*lo, hi*hi
                  }
                }
              }
            } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j) {
                const uint32_t a_off = a_addr + tap_off + mt * 128 + j * 2 * kALbo;
                const uint64_t a_hi = umma_desc(a_off, kALbo, kASbo);
                const uint64_t a_lo = umma_desc(a_off + kAPlane, kALbo, kASbo);
                const uint64_t b_hi = umma_desc(b_addr + j * 2 * kBLbo, kBLbo, 128);
                if (p.single) {
                  umma_bf16(d_tmem + mt * Cfg::kAccCols, a_hi, b_hi, idesc, (kb > kb0 || (tap | j) != 0));
                } else if (Cfg::kConcat) {
                  umma_bf16(d_tmem + mt * Cfg::kAccCols, a_hi, b_hi, idesc_cat, (kb > kb0 || (tap | j) != 0));   // hi*hi | hi*lo
                  umma_bf16(d_tmem + mt * Cfg::kAccCols, a_lo, b_hi, (p.debug & 64) ? idesc_cat : idesc, 1);   // lo*hi (64: + lo*lo)
                } else {
                  const uint64_t b_lo = umma_desc(b_addr + kBPlane + j * 2 * kBLbo, kBLbo, 128);
                  umma_bf16(d_tmem + mt * NT, a_lo, b_hi, idesc, (kb > kb0 || (tap | j) != 0));
                  umma_bf16(d_tmem + mt * NT, a_hi, b_lo, idesc, 1);
                  umma_bf16(d_tmem + mt * NT, a_hi, b_hi, idesc, 1);
                }
              }
            }
            }
            umma_commit(&b_empty[bs]);     // weight slab consumed
          }
          umma_commit(&a_empty[as]);       // halo block consumed by all nine taps
        }
        umma_commit(&tfull[acc]);          // accumulators complete
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int ew = warp - 4;
    const int r = ew * 32 + lane;                  // tile row: pixel (r & 7, r >> 3) of each 8x16 sub-tile
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const size_t plane_stride = static_cast<size_t>(p.B) * p.cout * p.Hout * p.Wout;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int ks = tile % p.ksplit, otile = tile / p.ksplit;
      const int n_tile = otile / p.m_tiles;
      int m = otile - n_tile * p.m_tiles;
      const int m_tile = m;
      const int tx = m % p.tiles_x;
      m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = m / p.tiles_y;
      const int y = ty * Cfg::kTileH + (r >> 3);
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < ((p.debug & 8) ? 0 : MT); ++mt) {      // experiment 8: no epilogue work
        const int x = tx * Cfg::kTileW + mt * 8 + (r & 7);
        const bool valid = y < p.H && x < p.W;
        float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < NT; c += 32) {
          float v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (acc * MT + mt) * Cfg::kAccCols + c;
          tmem_ld32(taddr, v);
          if (Cfg::kConcat && !p.single) {
            float w[32];
            tmem_ld32(taddr + NT, w);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += w[e];
          } else {
            tmem_ld_wait();
          }
          if (p.ksplit > 1) {          // raw partial sums of this K slice -> kpart[ks][n_tile][m_tile * MT + mt][row][NT]
            float4* dst = reinterpret_cast<float4*>(
                p.kpart + (((static_cast<size_t>(ks) * p.n_tiles + n_tile) * p.m_tiles + m_tile) * MT + mt) * (kTileM * NT) +
                static_cast<size_t>(r) * NT + c);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            continue;
          }
          if (valid) epilogue_32cols(p, v, n_tile * NT + c, b, y, x, nw, plane_stride, rgb0, rgb1, rgb2);
        }
        if (valid && p.rgb_coef && p.ksplit == 1) rgb_store(p, n_tile, b, y, x, rgb0, rgb1, rgb2);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
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
template <int NT, int MT, bool FUSED>
static int launch_halo_impl(const ConvKernelParams& p, const CUtensorMap& tmap, int sms, cudaStream_t stream,
                            const FusedFirParams& f) {
  using Cfg = HaloCfg<NT, MT, FUSED>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(modconv_halo_kernel<NT, MT, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("modconv_halo: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.ksplit;
  CUtensorMap tmap_planes = tmap;            // not fused: unused placeholder
  if (FUSED && make_plane_tensor_map(&tmap_planes, f.t, p.B, f.C, f.Hin + 1, f.Win + 1, Cfg::kPC)) return 1;
  const cudaError_t le = launch_pdl(modconv_halo_kernel<NT, MT, FUSED>, dim3(std::min(total, sms)), dim3(FUSED ? kFusedThreads : 256),
                                    Cfg::kSmemBytes, stream, tmap, tmap_planes, p, f);
  count_launch();
  if (le != cudaSuccess) {
    set_error("modconv_halo_kernel: launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("modconv_halo_kernel") ? 0 : 1;
}

template <int NT, int MT>
static int launch_halo(const ConvKernelParams& p, const CUtensorMap& tmap, int sms, cudaStream_t stream,
                       const FusedFirParams* fused) {
  if (fused) return launch_halo_impl<NT, MT, true>(p, tmap, sms, stream, *fused);
  FusedFirParams none;
  memset(&none, 0, sizeof(none));
  return launch_halo_impl<NT, MT, false>(p, tmap, sms, stream, none);
}

bool halo_eligible(const sgr_conv_args* a) {
  static const bool off = [] { const char* e = getenv("SGR_HALO"); return e && e[0] == '0'; }();
  if (off) return false;
  return a->ksize == 3 && a->up == 0 && a->h_in >= 16 && a->w_in >= 16 && a->cout >= 32;
}

// Fills the tile geometry of `p` (already filled by conv_fill_params) for the halo kernel and launches it.
// A halo convolution can take its input straight from the parity planes of the preceding scatter up-conv (fused FIR
// producer warps) when nothing else needs that activation: fp32-parity mode, no split-K, channel blocks of 32.
// On by default (SGR_FUSE_FIR=0 restores the separate up_finish_kernel pass): removes the FIR pass (0.51 ms of a 3.40 ms
// step at B=32, 256^2) and 2.1 GB of its traffic; the eight producer warps cost the five consumers 0.33 ms in total (shared-
// memory bandwidth: MMA operand reads + producer loads run at ~125 of 128 B/clk), net 3.22 vs 3.40 ms.  With the FIR work
// switched off (SGR_DEBUG=512, handshakes only) the step is 2.96 ms, the bound a cheaper producer can approach.
bool halo_fusable(const sgr_conv_args* a) {
  // SGR_FUSE_FIR: 0 = never, 1 = wherever possible, unset = consumers with at least 128 input channels.  Measured on one B200
  // (B = 32, 256^2 / cm1, tools/gpu/call18.sh), fused producers vs the separate FIR pass (up_finish_kernel) per consumer:
  //   512 ch @ 16^2 + 32^2: consumer +0.030 ms, pass 0.072 ms   256 ch @ 64^2: +0.029 vs 0.072   128 ch @ 128^2: +0.098 vs 0.132
  //   64 ch @ 256^2: +0.247 vs 0.239 — a wash: that consumer is bound by shared-memory bandwidth (operand reads + producers
  //   ~0.85 wavefronts per cycle), so it runs the plain kernel (0.45 ms, at the N = 64 MMA rate) behind the HBM-bound pass.
  // Whole step: all fused 3.10 .. 3.15 ms, none 3.18 ms, this policy the same as all fused.
  static const int min_cin = [] {
    const char* e = getenv("SGR_FUSE_FIR");
    if (!e) return 128;
    if (e[0] == '0' && !e[1]) return 1 << 30;
    if (e[0] == '1' && !e[1]) return 0;
    return atoi(e);                                   // any other number: the minimum input-channel count of a fused consumer
  }();
  if (!(a->cin >= min_cin && halo_eligible(a) && !a->single_pass && a->cin % 32 == 0 && a->h_in % 2 == 0 && a->w_in % 2 == 0)) return false;
  // layers with fewer tiles than half the SMs (small batches) run split-K, which the fused producers do not support
  const int nt = a->column_tile > 0 ? a->column_tile : pick_nt(a->cout);
  const int mt = nt <= 128 ? 2 : 1;
  const int tiles = ((a->w_in + 8 * mt - 1) / (8 * mt)) * ((a->h_in + 15) / 16) * a->batch * (a->cout / nt);
  return 2 * tiles > num_sms();
}

int launch_modconv_halo(const sgr_conv_args* a, ConvKernelParams p, cudaStream_t stream, const FusedFirParams* fused) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    set_error("modconv_halo: no CUDA device");
    return 1;
  }
  const int nt = a->column_tile > 0 ? a->column_tile : pick_nt(a->cout);
  const int mt = nt <= 128 ? 2 : 1;
  p.bw = 8 * mt; p.bh = 16; p.bb = 1; p.rows = 128;
  p.tiles_x = (a->w_in + p.bw - 1) / p.bw;
  p.tiles_y = (a->h_in + p.bh - 1) / p.bh;
  p.tiles_b = a->batch;
  p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  p.n_tiles = a->cout / nt;
  p.halo_mt = mt;
  p.nt = nt;
  set_ksplit(&p, fused ? 1 : choose_ksplit(a, p.m_tiles * p.n_tiles, p.kchunks, 2, static_cast<size_t>(mt) * kTileM * nt * 4));
  CUtensorMap tmap;
  if (make_act_tensor_map(&tmap, a->x_c8, a->batch, a->cin, a->h_in, a->w_in, p.bw + 2, p.bh + 2, 1, p.single ? 1 : 2)) return 1;
  int rc;
  switch (nt) {
    case 256: rc = launch_halo<256, 1>(p, tmap, sms, stream, fused); break;
    case 128: rc = launch_halo<128, 2>(p, tmap, sms, stream, fused); break;
    case 64: rc = launch_halo<64, 2>(p, tmap, sms, stream, fused); break;
    case 32: rc = launch_halo<32, 2>(p, tmap, sms, stream, fused); break;
    default: set_error("modconv_halo: unsupported column tile %d", nt); return 1;
  }
  if (rc == 0 && p.ksplit > 1) rc = splitk_finish_launch(p, stream);
  return rc;
}

}  // namespace sgr

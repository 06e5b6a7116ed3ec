// Upsampling modulated convolution, scatter form: conv_transpose2d(stride 2) (model.py:246-254) at its MINIMAL flop
// count on the sm_100a tensor cores.  The (2H+1)^2 intermediate t[2i+ky, 2j+kx] += x[i,j] W[ky,kx] splits into four
// parity planes on the (H+1) x (W+1) grid
//   ee[I,J] = sum_{a,b} x[I-a,J-b] W[2a,2b]   eo[I,J] = sum_a x[I-a,J] W[2a,1]
//   oe[I,J] = sum_b x[I,J-b] W[1,2b]          oo[I,J] = x[I,J] W[1,1]
// i.e. four shifted activation tiles (shift s = (a,b)) times column blocks [oe|ee|eo|oo] of one accumulator: shift (0,0)
// feeds all four blocks (N = NT), (0,1) feeds [oe|ee] (N = NT/2), (1,0) feeds [ee|eo] (N = NT/2), (1,1) feeds [ee]
// (N = NT/4): 9 Cin Cout MACs per input pixel instead of the 36 of the polyphase form.  The raw planes go to HBM (fp32)
// and up_finish_kernel applies the 4x4 FIR of the Blur (model.py:72-88) + the fused StyledConv epilogue.
//
// GEMM rows: a dense (bw x bh x bb) box of the parity grid, rows <= 128 (tile_box_search); TMA zero-fills x[-1], x[H].
// Pipeline: the kernel is bound by bytes in flight between L2 and shared memory, so the activation tiles (ring of
// 4 x 16 KiB, slot == shift) and the weight slabs (ring of kGroups channel blocks x [NT, NT/2, NT/2, NT/4] rows, exact
// sizes) have separate rings and barriers.  The single MMA-issuing thread is the critical resource on the N = NT/4 and
// NT/2 steps (6 MMAs of ~50 ns): its loop carries no division or descriptor construction - ring slots and phases follow
// from the running channel-block counter, descriptors are base + offset adds.
// Warp roles as in modconv_sm100.cu.
#include <algorithm>

#include <stdlib.h>

#include "sgr_internal.h"
#include "sgr_ptx.cuh"

namespace sgr {

template <int NT>
struct ScatterCfg {
  static constexpr int kGroupRows = NT / 4 * 9;                          // weight rows per 32-channel block
  static constexpr int kGroupBytes = kGroupRows * 128;
  static constexpr int kGroups = (144 * 1024) / kGroupBytes;             // 2 (NT=256), 4 (NT=128)
  static constexpr int kBSlabs = kGroups * 4;
  static constexpr int kAStages = 4;                                      // one slot per shift: slot index == shift
  static constexpr int kHaloStages = 3;                                   // wrapped-halo tiles: ring of channel blocks in the same 64 KiB
  static constexpr int kHaloStageBytes = 8 * 144 * 16 + 2304;             // 8 chunk-planes of <= 144 box entries + the over-read of the last one
  // [barriers | A ring | epilogue staging (8 channel groups x 128 pixels x 16 B) | B ring]; linear tiles store per lane and
  // use the staging area as part of their A ring
  static constexpr int kStageOff = 1024 + kAStages * kABytes;
  static constexpr int kARegion = kAStages * kABytes + 8 * kTileM * 16;
  static constexpr int kBOff = 1024 + kARegion;
  static constexpr int kSmemBytes = kBOff + kGroups * kGroupBytes;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory");
  static_assert(kBSlabs <= 16, "barrier block");
};

__device__ __forceinline__ int scatter_rows(int nt, int sft) { return nt >> ((sft + 1) >> 1); }          // NT, NT/2, NT/2, NT/4
__device__ __forceinline__ int scatter_prefix(int nt, int sft) {                                          // rows before shift
  return sft == 0 ? 0 : (sft == 1 ? nt : (sft == 2 ? nt + nt / 2 : 2 * nt));
}

// Linear tiles (LIN; the 33^2 and 65^2 grids, where no wrapped-halo box shape fits the grid without a third more tiles): the
// parity grid of one sample, padded by a halo row above and a halo column on the left (pitch P = W + 1 entries), is cut into
// runs of 128 CONSECUTIVE entries.  Tile t covers entries P + 1 + 128 t ..; shift (a, b) of row r is entry - (a P + b), so one
// box of P x R entries (R = 6 / 4 rows for P = 34 / 66) per channel block holds all four shifted operands at constant
// offsets, exactly like the wrapped-halo box with bw = P plus a per-tile start offset.  Against the four dense boxes per
// channel block this loads 26 / 34 KB instead of 62 KB and — what the layer was bound by — leaves room for a ring of 3 / 2
// channel blocks in flight where the four dense slots were refilled inside the channel block that consumed them.
template <int NT, int HE, bool LIN = false>
                               // HE > 0: box entries per chunk-plane of the wrapped-halo tiles (144 or 140) or linear tiles (204, 264);
                               // HE < 0: dense-box tiles of exactly -HE rows; 0: dense-box tiles of any shape
__global__ void __launch_bounds__(256, 1) upconv_scatter_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                const __grid_constant__ CUtensorMap tmap_out,
                                                                const ConvKernelParams p) {
  using Cfg = ScatterCfg<NT>;
  constexpr bool HALO = HE > 0;
  constexpr int kHE = HE > 0 ? HE : (HE < 0 ? -HE : 144);      // entries per chunk-plane when they are a compile-time constant
  constexpr bool kFixed = HE != 0;
  constexpr int kHStages = LIN ? Cfg::kARegion / (8 * kHE * 16) : Cfg::kHaloStages;          // 3 (204 entries), 2 (264)
  constexpr int kHStageBytes = LIN ? 8 * kHE * 16 : Cfg::kHaloStageBytes;
  static_assert(kHStages >= 2 && kHStages <= Cfg::kAStages, "A ring");
  constexpr int AS = Cfg::kAStages, BSL = Cfg::kBSlabs;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_empty = a_full + AS;
  uint64_t* b_full = a_empty + AS;
  uint64_t* b_empty = b_full + BSL;
  uint64_t* tfull = b_empty + BSL;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* a_base = smem + 1024;
  uint8_t* b_base = smem + Cfg::kBOff;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AS; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < BSL; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
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

  // work item = (output tile, K slice): with fewer tiles than SMs (small batches, the 5^2 .. 33^2 grids) the channel blocks
  // of a tile are cut into p.ksplit slices run by different CTAs; slice ks writes its raw planes to copy ks of the plane
  // tensor ([ksplit][B][4][cout/4][H][W][4]) and up_finish_kernel adds the copies in slice order (deterministic)
  const int total_tiles = p.m_tiles * p.n_tiles * p.ksplit;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      const uint32_t a_bytes = static_cast<uint32_t>(p.rows) * (p.single ? 64u : 128u);   // single pass: hi plane only
      const uint32_t halo_bytes = static_cast<uint32_t>(kHE) * (p.single ? 64u : 128u);      // box = bw x box_rows entries
      const uint32_t b_row = p.single ? 64u : 128u;                                       // slabs are plane-major: hi half first
      uint32_t g = 0;                                  // running channel-block counter: ring slots and phases
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int ks = item % p.ksplit, tile = item / p.ksplit;
        const int kb0 = ks * p.kchunks / p.ksplit, kb1 = (ks + 1) * p.kchunks / p.ksplit;
        const int n_tile = tile / p.m_tiles;
        int m = tile - n_tile * p.m_tiles;
        const int tx = m % p.tiles_x;
        m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int tb = m / p.tiles_y;
        const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
        const int hx0 = tx * (p.bw - 1), hy0 = ty * p.bh;                  // wrapped-halo tiles: bw - 1 valid columns
        const uint8_t* wkc = reinterpret_cast<const uint8_t*>(p.wpacked) +
                             (static_cast<size_t>(n_tile) * p.kchunks + kb0) * Cfg::kGroupBytes;
        for (int kc = kb0; kc < kb1; ++kc, ++g, wkc += Cfg::kGroupBytes) {
          const uint32_t grp = g % Cfg::kGroups;
          const uint32_t a_par = (g & 1) ^ 1, b_par = ((g / Cfg::kGroups) & 1) ^ 1;
          uint8_t* bdst = b_base + grp * Cfg::kGroupBytes;
          if (HALO && !((p.debug & 16) && g >= kHStages)) {   // one box per channel block: tile + top row / left column halo
            const uint32_t hs = g % kHStages;
            mbar_wait(&a_empty[hs], ((g / kHStages) & 1) ^ 1);
            mbar_expect_tx(&a_full[hs], halo_bytes);
            if (LIN)      // 8-byte elements (two per entry): rows (128 ty) / P - 1 .. of the sample, from the halo column on
              tma_load_5d(a_base + hs * kHStageBytes, &tmap, &a_full[hs], -2, (ty * kTileM) / p.bw - 1, b0, kc * 4, 0);
            else
              tma_load_5d(a_base + hs * kHStageBytes, &tmap, &a_full[hs], (hx0 - 1) * 8, hy0 - 1, b0, kc * 4, 0);
          }
#pragma unroll
          for (int sft = 0; sft < 4; ++sft) {
            const int n_s = scatter_rows(NT, sft), prefix = scatter_prefix(NT, sft);
            const uint32_t bs = grp * 4 + sft;
            if (!HALO) {
              mbar_wait(&a_empty[sft], a_par);
              mbar_expect_tx(&a_full[sft], a_bytes);
              tma_load_5d(a_base + sft * kABytes, &tmap, &a_full[sft], (x0 - (sft & 1)) * 8, y0 - (sft >> 1), b0, kc * 4, 0);
            }
            if ((p.debug & 4) && g >= Cfg::kGroups) continue;          // experiment 4: weight ring loaded once
            mbar_wait(&b_empty[bs], b_par);
            mbar_expect_tx(&b_full[bs], n_s * b_row);
            bulk_g2s(bdst + prefix * 128, wkc + prefix * 128, n_s * b_row, &b_full[bs]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    // The issuing thread is the critical resource (see modconv_halo_sm100.cu): ~8 cycles per instruction against 48 .. 128 cycles
    // per MMA.  Ring base descriptors are built once; per channel block one multiply-add per ring, per shift one add, and
    // every MMA's operands are base + compile-time constant (wrapped-halo tiles have a FIXED chunk stride of kHaloEntries
    // box entries for that; the dense-box tiles keep run-time strides).
    if (elect_one_sync()) {
      const uint32_t a_lbo = kFixed ? kHE * 16u : static_cast<uint32_t>(p.rows) * 16u;   // bytes per chunk-plane of a box
      const uint64_t a_ring = umma_desc(smem_u32(a_base), a_lbo, 128);
      // wrapped-halo tiles: M row r of shift (a, b) = box entry bw + 1 + r - (a * bw + b)
      const uint64_t h_off[4] = {static_cast<uint64_t>(p.bw + 1), static_cast<uint64_t>(p.bw), 1, 0};
      const uint64_t a_lo_off = (a_lbo * 4) >> 4, a_j_off = (a_lbo * 2) >> 4;
      constexpr uint64_t kHLo = (kHE * 16 * 4) >> 4, kHJ = (kHE * 16 * 2) >> 4;
      uint64_t b_ring[4];
      uint32_t idesc[4];
#pragma unroll
      for (int sft = 0; sft < 4; ++sft) {
        const uint32_t n_s = scatter_rows(NT, sft);
        b_ring[sft] = umma_desc(smem_u32(b_base) + scatter_prefix(NT, sft) * 128, n_s * 16, 128);
        idesc[sft] = umma_idesc(p.fmt, kTileM, static_cast<int>(n_s));
      }
      uint32_t g = 0, tcount = 0;
      uint32_t hs = 0, hphase = 0, grp = 0, gphase = 0;      // ring slot + phase of the next halo stage / weight group
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++tcount) {
        const int ks = item % p.ksplit, tile = item / p.ksplit;
        const int kb0 = ks * p.kchunks / p.ksplit, kb1 = (ks + 1) * p.kchunks / p.ksplit;
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NT;
        uint64_t lin_off = 0;            // linear tiles: first entry of the tile inside its box
        if (LIN) {
          const int t0 = ((tile % p.m_tiles) % p.tiles_y) * kTileM;
          lin_off = static_cast<uint64_t>(t0 - (t0 / p.bw) * p.bw);
        }
        for (int kc = kb0; kc < kb1; ++kc, ++g) {
          const uint32_t a_par = g & 1;
          const uint64_t grp_off = static_cast<uint64_t>(grp) * (Cfg::kGroupBytes >> 4);
          const bool b_skip = (p.debug & 4) && g >= Cfg::kGroups;
          const uint32_t acc0 = kc != kb0 ? 1u : 0u;
          if (kFixed && !p.single) {
            if (HALO && !((p.debug & 16) && g >= kHStages)) mbar_wait(&a_full[hs], hphase);
            const uint64_t a_st = a_ring + static_cast<uint64_t>(hs) * (kHStageBytes >> 4) + lin_off;
#pragma unroll
            for (int sft = 0; sft < 4; ++sft) {
              constexpr uint32_t kCol[4] = {0, 0, NT / 4, NT / 4};          // [oe|ee|eo|oo]: shifts (1,0), (1,1) start at ee
              const uint32_t n_s = scatter_rows(NT, sft);
              if (!HALO) mbar_wait(&a_full[sft], a_par);
              if (!b_skip) mbar_wait(&b_full[grp * 4 + sft], gphase);
              tc_fence_after();
              const uint64_t a_s = HALO ? a_st + h_off[sft] : a_ring + static_cast<uint64_t>((sft * kABytes) >> 4);
              const uint64_t b_s = b_ring[sft] + grp_off;
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j) {
                const uint64_t bj = static_cast<uint64_t>(j) * ((n_s * 32) >> 4), blo = (n_s * 64) >> 4;
                umma_bf16(d_tmem + kCol[sft], a_s + (j * kHJ + kHLo), b_s + bj, idesc[sft], (sft | j) != 0 ? 1u : acc0);
                umma_bf16(d_tmem + kCol[sft], a_s + j * kHJ, b_s + (bj + blo), idesc[sft], 1);
                umma_bf16(d_tmem + kCol[sft], a_s + j * kHJ, b_s + bj, idesc[sft], 1);
              }
              if (!HALO) umma_commit(&a_empty[sft]);
              else if (sft == 3) umma_commit(&a_empty[hs]);
              umma_commit(&b_empty[grp * 4 + sft]);
            }
          } else {
#pragma unroll
            for (int sft = 0; sft < 4; ++sft) {
              const uint32_t n_s = scatter_rows(NT, sft);
              const uint32_t coloff = sft >= 2 ? NT / 4 : 0;
              if (!HALO) mbar_wait(&a_full[sft], a_par);
              else if (sft == 0 && !((p.debug & 16) && g >= kHStages)) mbar_wait(&a_full[hs], hphase);
              if (!b_skip) mbar_wait(&b_full[grp * 4 + sft], gphase);
              tc_fence_after();
              const uint64_t a_hi0 = HALO ? a_ring + static_cast<uint64_t>(hs) * (kHStageBytes >> 4) + lin_off + h_off[sft]
                                            : a_ring + ((sft * kABytes) >> 4);
              const uint64_t b_hi0 = b_ring[sft] + grp_off;
#pragma unroll
              for (int j = 0; j < kBlockK / 16; ++j) {
                const uint64_t a_hi = a_hi0 + j * a_j_off, a_lo = a_hi + a_lo_off;
                const uint64_t b_hi = b_hi0 + j * ((n_s * 32) >> 4), b_lo = b_hi + ((n_s * 64) >> 4);
                if (!p.single) {
                  umma_bf16(d_tmem + coloff, a_lo, b_hi, idesc[sft], (sft | j) != 0 ? 1u : acc0);
                  umma_bf16(d_tmem + coloff, a_hi, b_lo, idesc[sft], 1);
                  umma_bf16(d_tmem + coloff, a_hi, b_hi, idesc[sft], 1);
                } else {
                  umma_bf16(d_tmem + coloff, a_hi, b_hi, idesc[sft], (sft | j) != 0 ? 1u : acc0);
                }
              }
              if (!HALO) umma_commit(&a_empty[sft]);
              else if (sft == 3) umma_commit(&a_empty[hs]);
              umma_commit(&b_empty[grp * 4 + sft]);
            }
          }
          if (HALO && ++hs == kHStages) {
            hs = 0;
            hphase ^= 1;
          }
          if (++grp == Cfg::kGroups) {
            grp = 0;
            gphase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: raw parity planes -> HBM
    const int ew = warp - 4;
    const int r = ew * 32 + lane;
    int xx = r % p.bw;
    int yy = (r / p.bw) % p.bh;
    int bl = r / (p.bw * p.bh);
    int slot = r;                                      // position inside the staged / stored box; < 0: nothing to store
    int sx = p.bw, sy = p.bh;                          // tile stride on the grid
    if (HALO) {                                      // box entry bw + 1 + r: column 0 of the box is halo (garbage rows)
      const int e = p.bw + 1 + r;
      xx = e % p.bw - 1;
      yy = e / p.bw - 1;
      bl = 0;
      sx = p.bw - 1;
      slot = (xx >= 0 && yy < p.bh) ? yy * (p.bw - 1) + xx : -1;
    } else if (r >= p.rows) {
      slot = -1;
    }
    constexpr int CT = NT / 4;
    const size_t group_stride = static_cast<size_t>(p.H) * p.W * 4;
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++tcount) {
      const int ks = item % p.ksplit, tile = item / p.ksplit;
      const int n_tile = tile / p.m_tiles;
      int m = tile - n_tile * p.m_tiles;
      const int tx = m % p.tiles_x;
      m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int tb = m / p.tiles_y;
      int b = tb * p.bb + bl, y = ty * sy + yy, x = tx * sx + xx;
      bool valid = slot >= 0 && b < p.B && y < p.H && x < p.W;
      if (LIN) {                                          // entry P + 1 + 128 ty + r of the padded grid (tiles_x == 1)
        const int e = p.bw + 1 + ty * kTileM + r;
        y = e / p.bw - 1;
        x = e - (y + 1) * p.bw - 1;
        b = tb;
        valid = x >= 0 && y < p.H;
      }
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      // TMEM loads are double buffered: the load of columns c+32.. is in flight while columns c.. are stored
      float v[2][32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * NT;
      if (p.debug & 2) {            // experiment: no TMEM reads, no stores
        tc_fence_before();
        mbar_arrive(&tempty[acc]);
        continue;
      }
      tmem_ld32(taddr, v[0]);
#pragma unroll
      for (int ci = 0; ci < NT / 32; ++ci) {
        tmem_ld_wait();
        if (ci + 1 < NT / 32) tmem_ld32(taddr + (ci + 1) * 32, v[(ci + 1) & 1]);
        // t[b][plane][cout/4][H+1][W+1][4] (fp32): a lane (pixel) stores 16 B per 4-channel group, consecutive lanes are
        // consecutive pixels -> every store instruction writes whole 32 B sectors (a [..][8] layout with 32 B per pixel
        // and two half-sector stores per lane ran the epilogue at half the L2 write rate and stalled the MMA pipe)
        if (p.tma_store) {
          // 32 columns = 8 channel groups of one parity plane: staged as the dense box [group][y][x][4] and stored by ONE TMA
          // tensor store (clipped at the grid edge) — 64 STG.128 per lane and tile saturate the SM -> L2 store path (the same
          // stores to L2-resident lines are no faster), a bulk store of the staged tile is cheaper
          uint8_t* stg = smem + Cfg::kStageOff;
          if (r == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // previous box read out of the staging
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (slot >= 0) {
            const float* vv = v[ci & 1];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(stg + (q * p.rows + slot) * 16) = make_float4(vv[4 * q], vv[4 * q + 1], vv[4 * q + 2], vv[4 * q + 3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (r == 0) {
            const int c = ci * 32;
            const int plane = c / CT;
            const int g0 = (n_tile * CT + (c % CT)) >> 2;
            asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(stg)), "r"(tx * sx * 4), "r"(ty * sy), "r"(g0),
                         "r"(plane), "r"(ks * p.B + tb)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          continue;
        }
        if (valid && !(p.debug & 1)) {
          const int c = ci * 32;
          const int plane = c / CT;
          const int o0 = n_tile * CT + (c % CT);
          float* tptr = p.t_out + ((static_cast<size_t>(ks * p.B + b) * 4 + plane) * (p.cout >> 2) + (o0 >> 2)) * group_stride +
                        (static_cast<size_t>(y) * p.W + x) * 4;
          const float* vv = v[ci & 1];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(tptr + q * group_stride) = make_float4(vv[4 * q], vv[4 * q + 1], vv[4 * q + 2], vv[4 * q + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }

  if (p.tma_store && warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NT>
static int launch_scatter_nt(const ConvKernelParams& p, const CUtensorMap& tmap, int sms, cudaStream_t stream) {
  using Cfg = ScatterCfg<NT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(upconv_scatter_kernel<NT, 144>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_scatter_kernel<NT, 140>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_scatter_kernel<NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_scatter_kernel<NT, -121>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_scatter_kernel<NT, 204, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upconv_scatter_kernel<NT, 264, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("upconv_scatter: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return 1;
    }
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.ksplit;
  // TMA-store epilogue: one sample per tile (the layers of 16x16 and larger), whole 32-column chunks of one plane
  static const bool tma_off = [] { const char* e = getenv("SGR_TMA_STORE"); return e && e[0] == '0'; }();
  ConvKernelParams q = p;
  CUtensorMap tmap_out = tmap;
  q.tma_store = (!tma_off && p.bb == 1 && p.halo != 2 && NT / 4 >= 32 && !(p.debug & 3)) ? 1 : 0;   // linear tiles are no boxes
  if (q.tma_store && make_plane_tensor_map(&tmap_out, p.t_out, p.B * p.ksplit, p.cout, p.H, p.W, p.halo ? p.bw - 1 : p.bw, p.bh, 8, 1)) return 1;
  const int he = p.halo ? p.bw * p.box_rows : 0;
  cudaError_t le = cudaSuccess;
  if (p.halo == 2 && he == 204)
    le = launch_pdl(upconv_scatter_kernel<NT, 204, true>, dim3(std::min(total, sms)), dim3(256), Cfg::kSmemBytes, stream, tmap, tmap_out, q);
  else if (p.halo == 2 && he == 264)
    le = launch_pdl(upconv_scatter_kernel<NT, 264, true>, dim3(std::min(total, sms)), dim3(256), Cfg::kSmemBytes, stream, tmap, tmap_out, q);
  else if (p.halo == 2) {
    set_error("upconv_scatter: unsupported linear box of %d entries", he);
    return 1;
  } else if (he == 144)
    le = launch_pdl(upconv_scatter_kernel<NT, 144>, dim3(std::min(total, sms)), dim3(256), Cfg::kSmemBytes, stream, tmap, tmap_out, q);
  else if (he == 140)
    le = launch_pdl(upconv_scatter_kernel<NT, 140>, dim3(std::min(total, sms)), dim3(256), Cfg::kSmemBytes, stream, tmap, tmap_out, q);
  else if (he == 0 && p.rows == 121)
    le = launch_pdl(upconv_scatter_kernel<NT, -121>, dim3(std::min(total, sms)), dim3(256), Cfg::kSmemBytes, stream, tmap, tmap_out, q);
  else if (he == 0)
    le = launch_pdl(upconv_scatter_kernel<NT, 0>, dim3(std::min(total, sms)), dim3(256), Cfg::kSmemBytes, stream, tmap, tmap_out, q);
  else {
    set_error("upconv_scatter: unsupported halo box of %d entries", he);
    return 1;
  }
  count_launch();
  if (le != cudaSuccess) {
    set_error("upconv_scatter_kernel: launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return check_launch("upconv_scatter_kernel") ? 0 : 1;
}

int launch_upconv_scatter(const ConvKernelParams& p, const CUtensorMap& tmap, int nt, cudaStream_t stream) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    set_error("upconv_scatter: no CUDA device");
    return 1;
  }
  switch (nt) {
    case 256: return launch_scatter_nt<256>(p, tmap, sms, stream);
    case 128: return launch_scatter_nt<128>(p, tmap, sms, stream);
    default: set_error("upconv_scatter: unsupported column tile %d", nt); return 1;
  }
}

}  // namespace sgr

// Second pass of a split-K convolution (modconv_sm100.cu / modconv_halo_sm100.cu with p.ksplit > 1): adds the fp32
// partial accumulators of the K slices in slice order (deterministic) and applies the same fused epilogue
// (demodulation, NoiseInjection, FusedLeakyReLU, next style + hi/lo split, ToRGB partials: modconv_epilogue.cuh).
// Used on the layers whose pixel grid yields fewer output tiles than SMs (4x4 .. 16x16, or any layer at batch 1-4).
#include "sgr_internal.h"
#include "sgr_ptx.cuh"
#include "modconv_epilogue.cuh"

namespace sgr {

// FULL tiles (every row is a pixel; few slices).  Block = 128 rows x kColGroups column groups: thread (r, cg) owns the 32-column chunks cg, cg + kColGroups, ... of its
// row, so that a tile with few rows-of-work still spreads over 4x the threads; the ToRGB partial sums of a row are
// combined across the column groups in group order through shared memory (deterministic).
constexpr int kColGroups = 4;
__global__ void __launch_bounds__(128 * kColGroups) splitk_finish_rows_kernel(const ConvKernelParams p) {
  __shared__ float rgb_red[kColGroups][128][3];
  pdl_launch_dependents();
  pdl_wait();                                        // the partial sums come from the GEMM launch right before
  const int r = threadIdx.x & 127;
  const int cg = threadIdx.x >> 7;
  const int sub_tiles = p.halo_mt > 0 ? p.halo_mt : 1;
  const int msub = blockIdx.x;                       // m_tile * sub_tiles + sub
  const int m_tile = msub / sub_tiles, sub = msub - m_tile * sub_tiles;
  const int n_tile = blockIdx.y;
  int m = m_tile;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  const int tb = m / p.tiles_y;
  int b, y, x;
  bool valid;
  if (p.halo_mt > 0) {                               // 8x16-pixel sub-tiles, one sample per tile
    b = tb;
    y = ty * 16 + (r >> 3);
    x = tx * (8 * p.halo_mt) + sub * 8 + (r & 7);
    valid = y < p.H && x < p.W;
  } else {                                           // dense (bw, bh, bb) box
    const int xx = r % p.bw, yy = (r / p.bw) % p.bh, bl = r / (p.bw * p.bh);
    b = tb * p.bb + bl;
    y = ty * p.bh + yy;
    x = tx * p.bw + xx;
    valid = r < p.rows && b < p.B && y < p.H && x < p.W;
  }
  const int nt = p.nt;
  const size_t slice_stride = static_cast<size_t>(p.n_tiles) * p.m_tiles * sub_tiles * kTileM * nt;
  const float* src = p.kpart + ((static_cast<size_t>(n_tile) * p.m_tiles * sub_tiles + msub) * kTileM + r) * nt;
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  const size_t plane_stride = static_cast<size_t>(p.B) * p.cout * p.Hout * p.Wout;
  float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
  if (valid) {
    for (int c = cg * 32; c < nt; c += 32 * kColGroups) {
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = 0.f;
      // slices are added in order (deterministic); four slices' loads are issued together so that a thread does not pay
      // one memory latency per slice
      for (int ks = 0; ks < p.ksplit; ks += 4) {
        float4 t[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool on = ks + u < p.ksplit;
          const float* sp = src + static_cast<size_t>(on ? ks + u : ks) * slice_stride + c;
#pragma unroll
          for (int q = 0; q < 8; ++q) t[u][q] = on ? __ldcs(reinterpret_cast<const float4*>(sp + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[4 * q] += t[u][q].x; v[4 * q + 1] += t[u][q].y; v[4 * q + 2] += t[u][q].z; v[4 * q + 3] += t[u][q].w;
          }
      }
      epilogue_32cols(p, v, n_tile * nt + c, b, y, x, nw, plane_stride, rgb0, rgb1, rgb2);
    }
  }
  if (p.rgb_coef) {                                  // uniform branch
    rgb_red[cg][r][0] = rgb0;
    rgb_red[cg][r][1] = rgb1;
    rgb_red[cg][r][2] = rgb2;
    __syncthreads();
    if (cg == 0 && valid) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int g = 0; g < kColGroups; ++g) {
        s0 += rgb_red[g][r][0];
        s1 += rgb_red[g][r][1];
        s2 += rgb_red[g][r][2];
      }
      rgb_store(p, n_tile, b, y, x, s0, s1, s2);
    }
  }
}

// PARTLY FILLED tiles (4x4 .. 8x8 grids at small batches: 16 .. 64 of the 128 rows are pixels, up to 16 slices).
// Block = one 128-row tile (sub-tile) of one column tile, 512 threads, in passes of R = 128 rows (column tiles <= 128) or 64
// rows (column tile 256), i.e. up to 512 (row, 32-column chunk) pairs per pass:
//   phase 1: all threads add the K slices of the chunk, one float4 column group per thread and pass — consecutive threads read
//            consecutive columns (whole 128 B lines), every slice of an element is requested before the first add (one memory
//            latency per eight slices), slices are added in slice order (deterministic) — into a shared-memory tile;
//   phase 2: thread (row, 32-column chunk) runs the fused epilogue on its 32 columns; the ToRGB partial sums of a row are
//            combined across the chunks in chunk order through shared memory (deterministic).
// (One thread per (row, 32 columns) doing its own slice loop — the first version — left a 4x4 layer at batch 1 with 32 active
// threads issuing 128 dependent-latency-bound loads each: 20 us for 64 KB.)
constexpr int kFinThreads = 512;
__host__ __device__ constexpr int fin_rows(int nt) { return nt <= 128 ? 128 : 64; }
__global__ void __launch_bounds__(kFinThreads) splitk_finish_kernel(const ConvKernelParams p) {
  extern __shared__ float4 tile4[];                  // [R][nt / 4 + 1]
  __shared__ float rgb_red[kFinThreads][3];          // [chunk][R]
  pdl_launch_dependents();
  pdl_wait();                                        // the partial sums come from the GEMM launch right before
  const int sub_tiles = p.halo_mt > 0 ? p.halo_mt : 1;
  const int msub = blockIdx.x;                       // m_tile * sub_tiles + sub
  const int m_tile = msub / sub_tiles, sub = msub - m_tile * sub_tiles;
  const int n_tile = blockIdx.y;
  int m = m_tile;
  const int tx = m % p.tiles_x;
  m /= p.tiles_x;
  const int ty = m % p.tiles_y;
  const int tb = m / p.tiles_y;
  // tile row -> pixel of the GEMM grid (modconv_halo_kernel sub-tiles / modconv_kernel dense boxes)
  auto decode = [&](int r, int& b, int& y, int& x) -> bool {
    if (p.halo_mt > 0) {                             // 8x16-pixel sub-tiles, one sample per tile
      b = tb;
      y = ty * 16 + (r >> 3);
      x = tx * (8 * p.halo_mt) + sub * 8 + (r & 7);
      return y < p.H && x < p.W;
    }
    const int xx = r % p.bw, yy = (r / p.bw) % p.bh, bl = r / (p.bw * p.bh);      // dense (bw, bh, bb) box
    b = tb * p.bb + bl;
    y = ty * p.bh + yy;
    x = tx * p.bw + xx;
    return r < p.rows && b < p.B && y < p.H && x < p.W;
  };
  const int nt = p.nt, nt4 = nt >> 2, pitch4 = nt4 + 1;
  const size_t slice_stride = static_cast<size_t>(p.n_tiles) * p.m_tiles * sub_tiles * kTileM * nt;
  const float* src = p.kpart + (static_cast<size_t>(n_tile) * p.m_tiles * sub_tiles + msub) * kTileM * nt;
  const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
  const size_t plane_stride = static_cast<size_t>(p.B) * p.cout * p.Hout * p.Wout;
  const int R = fin_rows(nt);
  const int rl = threadIdx.x % R, ch = threadIdx.x / R;        // phase 2: row of the pass, 32-column chunk
#pragma unroll 1
  for (int r0 = 0; r0 < kTileM; r0 += R) {
    int b, y, x;
    const bool valid = decode(r0 + rl, b, y, x);
    if (!__syncthreads_or(valid)) continue;           // no pixel behind these rows (also: the previous pass's reads are done)
    for (int f = threadIdx.x; f < R * nt4; f += kFinThreads) {
      const int fr = f / nt4, c4 = f - fr * nt4;
      int fb, fy, fx;
      if (!decode(r0 + fr, fb, fy, fx)) continue;
      const float4* sp = reinterpret_cast<const float4*>(src + static_cast<size_t>(r0 + fr) * nt) + c4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ks = 0; ks < p.ksplit; ks += 8) {
        float4 t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          t[u] = ks + u < p.ksplit ? __ldcs(sp + static_cast<size_t>(ks + u) * (slice_stride >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w;
        }
      }
      tile4[fr * pitch4 + c4] = acc;
    }
    __syncthreads();
    float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
    if (valid && ch * 32 < nt) {
      float v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = tile4[rl * pitch4 + ch * 8 + q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
      epilogue_32cols(p, v, n_tile * nt + ch * 32, b, y, x, nw, plane_stride, rgb0, rgb1, rgb2);
    }
    if (p.rgb_coef) {                                  // uniform branch
      if (ch * 32 < nt) {
        rgb_red[ch * R + rl][0] = rgb0;
        rgb_red[ch * R + rl][1] = rgb1;
        rgb_red[ch * R + rl][2] = rgb2;
      }
      __syncthreads();
      if (ch == 0 && valid) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int g = 0; g * 32 < nt; ++g) {
          s0 += rgb_red[g * R + rl][0];
          s1 += rgb_red[g * R + rl][1];
          s2 += rgb_red[g * R + rl][2];
        }
        rgb_store(p, n_tile, b, y, x, s0, s1, s2);
      }
    }
  }
}

int splitk_finish_launch(const ConvKernelParams& p, cudaStream_t stream) {
  const int sub_tiles = p.halo_mt > 0 ? p.halo_mt : 1;
  dim3 grid(p.m_tiles * sub_tiles, p.n_tiles);
  if (p.nt > 256 || p.nt % 32 != 0) {
    set_error("splitk_finish: unsupported column tile %d", p.nt);
    return 1;
  }
  static bool configured = false;
  const int smem_max = fin_rows(128) * (128 / 4 + 1) * 16;         // the largest tile: 128 rows x 128 columns (+ pitch)
  if (!configured) {
    if (cudaFuncSetAttribute(splitk_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) != cudaSuccess) {
      set_error("splitk_finish: cudaFuncSetAttribute(smem=%d) failed", smem_max);
      return 1;
    }
    configured = true;
  }
  // rows of a tile that are pixels: the halo kernel's sub-tiles are full from 16x16 on; a dense box holds bw x bh x bb pixels
  const long long pixels = static_cast<long long>(p.B) * p.H * p.W;
  const bool full = p.halo_mt > 0 ? (p.H % 16 == 0 && p.W % (8 * p.halo_mt) == 0) : (p.rows == kTileM && pixels % kTileM == 0);
  if (full) {          // measured at batch 1 (ncu): 9.7 / 12.7 us against 12.8 / 16.2 us for the cooperative kernel below
    launch_pdl(splitk_finish_rows_kernel, grid, dim3(128 * kColGroups), 0, stream, p);
  } else {             // 4x4 / 8x8 at batch 1: 12.4 / 14.8 us against 19.7 / 22.2 us for the row kernel
    const int smem = fin_rows(p.nt) * (p.nt / 4 + 1) * 16;
    launch_pdl(splitk_finish_kernel, grid, dim3(kFinThreads), smem, stream, p);
  }
  count_launch();
  return check_launch("splitk_finish_kernel") ? 0 : 1;
}

}  // namespace sgr

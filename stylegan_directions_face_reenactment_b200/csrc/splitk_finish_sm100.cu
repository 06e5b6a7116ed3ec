// Second pass of a split-K convolution (modconv_sm100.cu / modconv_halo_sm100.cu with p.ksplit > 1): adds the fp32
// partial accumulators of the K slices in slice order (deterministic) and applies the same fused epilogue
// (demodulation, NoiseInjection, FusedLeakyReLU, next style + hi/lo split, ToRGB partials: modconv_epilogue.cuh).
// Used on the layers whose pixel grid yields fewer output tiles than SMs (4x4 .. 16x16, or any layer at batch 1-4).
#include "sgr_internal.h"
#include "sgr_ptx.cuh"
#include "modconv_epilogue.cuh"

namespace sgr {

// Block = 128 rows x kColGroups column groups: thread (r, cg) owns the 32-column chunks cg, cg + kColGroups, ... of its
// row, so that a tile with few rows-of-work still spreads over 4x the threads; the ToRGB partial sums of a row are
// combined across the column groups in group order through shared memory (deterministic).
constexpr int kColGroups = 4;
__global__ void __launch_bounds__(128 * kColGroups) splitk_finish_kernel(const ConvKernelParams p) {
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

int splitk_finish_launch(const ConvKernelParams& p, cudaStream_t stream) {
  const int sub_tiles = p.halo_mt > 0 ? p.halo_mt : 1;
  dim3 grid(p.m_tiles * sub_tiles, p.n_tiles);
  launch_pdl(splitk_finish_kernel, grid, dim3(128 * kColGroups), 0, stream, p);
  count_launch();
  return check_launch("splitk_finish_kernel") ? 0 : 1;
}

}  // namespace sgr

// tcgen05.mma issue-rate microbenchmark (sm_100a): how long does ONE SS-mode kind::f16 MMA (M=128, K=16) take as a function
// of N, of the alignment / row pitch of the A operand (no-swizzle K-major core matrices read in place out of a halo tile,
// as modconv_halo_sm100.cu / modconv_scatter_sm100.cu do), and of a concurrent shared-memory load (the fused FIR producers)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/microbench/mma_rate.cu && ./mma_rate
// Every SM runs the same loop (148 CTAs) so the numbers include the chip-level power behaviour.  Output: ns and SM cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../stylegan_directions_face_reenactment_b200/csrc/sgr_ptx.cuh"

using namespace sgr;

struct Case {
  int n;          // MMA N
  int a_off;      // byte offset of the A descriptor start inside its buffer (multiple of 16)
  int a_sbo;      // bytes between 8-row groups of A (128 = dense, 288 = 18-pixel halo rows, ...)
  int n2;         // second MMA shape interleaved (0 = none): alternating N / N2 on different columns
  int lsu;        // 1: warps 4..7 stream LDS.128 over 64 KiB concurrently (shared-memory pressure of producer warps)
  int iters;
};

__global__ void __launch_bounds__(256, 1) mma_rate_kernel(Case c, unsigned long long* out_cycles, unsigned long long* out_ns,
                                                          float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 32);
  uint8_t* a_base = smem + 1024;                 // 64 KiB region for A (covers any pitch)
  uint8_t* b_base = a_base + 64 * 1024;          // 32 KiB for B (N=256: 256 rows x 16 B x 2 k-slices = 8 KiB)
  uint8_t* l_base = b_base + 32 * 1024;          // 64 KiB streamed by the LSU warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_base)[i] = 0x3c003c00u;  // bf16 pairs
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    *stop = 0;
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc(kFmtBF16, 128, c.n);
    const uint32_t idesc2 = umma_idesc(kFmtBF16, 128, c.n2 ? c.n2 : c.n);
    const uint64_t adesc = umma_desc(smem_u32(a_base) + c.a_off, 16 * 1024, c.a_sbo);      // LBO: second k-slice 16 KiB away
    const uint64_t bdesc = umma_desc(smem_u32(b_base), 4096, 128);
    unsigned long long t0, t1, n0, n1;
    // warm-up
    for (int i = 0; i < 64; ++i) umma_bf16(tmem, adesc, bdesc, idesc, i != 0);
    umma_commit(bar);
    mbar_wait(bar, 0);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
    t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      umma_bf16(tmem, adesc, bdesc, idesc, 1);
      if (c.n2) umma_bf16(tmem + 256, adesc, bdesc, idesc2, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 1);
    t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
    *stop = 1;
    out_cycles[blockIdx.x] = t1 - t0;
    out_ns[blockIdx.x] = n1 - n0;
  } else if (warp >= 4 && c.lsu) {
    float acc = 0.f;
    int it = 0;
    while (!*stop) {
#pragma unroll 8
      for (int k = 0; k < 8; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(l_base + (((it * 8 + k) * 128 + (warp - 4) * 32 + lane) * 16) % (64 * 1024));
        acc += v.x + v.y + v.z + v.w;
      }
      ++it;
    }
    if (acc == 12345.f) sink[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  const int smem = 1024 + 160 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long *cyc, *ns;
  float* sink;
  cudaMalloc(&cyc, sms * 8);
  cudaMalloc(&ns, sms * 8);
  cudaMalloc(&sink, 4);
  std::vector<Case> cases;
  const int iters = 20000;
  for (int lsu = 0; lsu < 2; ++lsu) {
    for (int n : {32, 64, 96, 128, 192, 256}) {
      cases.push_back({n, 0, 128, 0, lsu, iters});          // dense, aligned
      cases.push_back({n, 16, 128, 0, lsu, iters});         // dense, start + 16 B (x-shift in a dense box)
      cases.push_back({n, 0, 288, 0, lsu, iters});          // 18-pixel halo rows (modconv_halo MT=2), aligned start
      cases.push_back({n, 304, 288, 0, lsu, iters});        // tap (1,1) of the halo tile
      cases.push_back({n, 0, 256, 0, lsu, iters});          // 16-pixel rows
      cases.push_back({n, 0, 160, 0, lsu, iters});          // 10-pixel halo rows (MT=1)
    }
    cases.push_back({128, 0, 288, 64, lsu, iters});         // concat pair: N=128 + N=64 alternating (modconv_halo NT=64)
    cases.push_back({256, 0, 128, 128, lsu, iters});
    cases.push_back({256, 0, 128, 64, lsu, iters});
  }
  printf("%-5s %-6s %-6s %-4s %-4s | %10s %10s %10s\n", "N", "a_off", "a_sbo", "N2", "lsu", "ns/iter", "cyc/iter", "TF/s chip");
  for (const Case& c : cases) {
    mma_rate_kernel<<<sms, 256, smem>>>(c, cyc, ns, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("case failed: %s\n", cudaGetErrorString(e));
      return 1;
    }
    std::vector<unsigned long long> hc(sms), hn(sms);
    cudaMemcpy(hc.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hn.data(), ns, sms * 8, cudaMemcpyDeviceToHost);
    double mc = 0, mn = 0;
    for (int i = 0; i < sms; ++i) {
      mc += hc[i];
      mn += hn[i];
    }
    mc /= sms * (double)c.iters;
    mn /= sms * (double)c.iters;
    const double flop = 2.0 * 128 * 16 * (c.n + c.n2);
    printf("%-5d %-6d %-6d %-4d %-4d | %10.2f %10.2f %10.1f\n", c.n, c.a_off, c.a_sbo, c.n2, c.lsu, mn, mc, flop / mn * sms / 1e3);
  }
  return 0;
}

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
  int a_rot;      // distinct A operands in rotation (1: every MMA re-reads the same A; 4: as a real K loop / tap loop does)
  int b_rot;      // distinct B operands in rotation
  int acc_rot;    // accumulators (TMEM column ranges) in rotation: 1 = one dependent accumulation chain
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
    uint64_t ad[4], bd[4];
    uint32_t tm[4];
    for (int k = 0; k < 4; ++k) {
      ad[k] = adesc + (((k % c.a_rot) * 6144) >> 4);           // distinct A tiles 6 KiB apart (inside the 16 KiB k-slice)
      bd[k] = bdesc + (((k % c.b_rot) * 8192) >> 4);           // distinct B slabs
      tm[k] = tmem + (k % c.acc_rot) * (512 / c.acc_rot);
    }
    for (int i = 0; i < c.iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_bf16(tm[k], ad[k], bd[k], idesc, 1);
        if (c.n2) umma_bf16(tm[k] + 256, ad[(k + 1) & 3], bd[k], idesc2, 1);
      }
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

// ---- v3/v4: repeating PATTERNS of up to 8 MMAs (shape N, TMEM column offset, B slab index + row offset, A tile index), fully
// unrolled at compile time (a runtime-length pattern loop is issue-bound at ~330 cycles per iteration).  Operands rotate
// through 4 distinct A tiles and 4 distinct B slabs per repetition, as a real K / tap loop does.
struct Pattern {
  int np;
  int n[8], col[8], bslab[8], brow[8], aidx[8];
  int iters;
  int a_lbo, a_sbo, b_lbo, rot, a_off;      // operand layouts (0 = defaults 8192 / 256 / 8192 / 4 / 0)
  int rand;                                 // 1: operands = pseudo-random bf16 in (-2, 2) (data-dependent switching power), 0: all 2^-7
};

template <int NP>
__global__ void __launch_bounds__(256, 1) mma_pattern_kernel(Pattern c, unsigned long long* out_cycles, unsigned long long* out_ns) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* a_base = smem + 1024;                 // 4 A tiles x (2 k-slices x 8 KiB) = 64 KiB
  uint8_t* b_base = a_base + 64 * 1024;          // 4 B slabs x (2 k-slices x 8 KiB) = 64 KiB (N <= 512 rows per slice)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 1024) / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (c.rand) {                              // two bf16 with random sign / mantissa, exponents 2^-2 .. 2^0
      uint32_t h = (i + 1u) * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x80ff80ffu) | 0x3e003e00u | ((h >> 3) & 0x01000100u);
    }
    reinterpret_cast<uint32_t*>(a_base)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
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
    // rep r of the pattern uses A tile (aidx + r) % 4 and B slab (bslab + r) % 4
    uint64_t ad[4][NP], bd[4][NP];
    uint32_t id[NP], tm[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      id[k] = umma_idesc(kFmtBF16, 128, c.n[k]);
      tm[k] = tmem + c.col[k];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rr = c.rot == 1 ? 0 : r;
        ad[r][k] = umma_desc(smem_u32(a_base) + ((c.aidx[k] + rr) & 3) * 16384 + c.a_off, c.a_lbo ? c.a_lbo : 8192, c.a_sbo ? c.a_sbo : 256);
        bd[r][k] = umma_desc(smem_u32(b_base) + ((c.bslab[k] + rr) & 3) * 16384 + c.brow[k] * 16, c.b_lbo ? c.b_lbo : 8192, 128);
      }
    }
    unsigned long long t0, t1, n0, n1;
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < NP; ++k) umma_bf16(tm[k], ad[i & 3][k], bd[i & 3][k], id[k], i != 0);
    umma_commit(bar);
    mbar_wait(bar, 0);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
    t0 = clock64();
    for (int i = 0; i < c.iters; i += 4) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < NP; ++k) umma_bf16(tm[k], ad[r][k], bd[r][k], id[k], 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 1);
    t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
    out_cycles[blockIdx.x] = t1 - t0;
    out_ns[blockIdx.x] = n1 - n0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int NP>
static void launch_pattern(const Pattern& p, int sms, unsigned long long* cyc, unsigned long long* ns, int smem) {
  cudaFuncSetAttribute(mma_pattern_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_pattern_kernel<NP><<<sms, 256, smem>>>(p, cyc, ns);
}

static double time_pattern(const Pattern& p, int sms, unsigned long long* cyc, unsigned long long* ns, int smem) {
  switch (p.np) {
    case 1: launch_pattern<1>(p, sms, cyc, ns, smem); break;
    case 2: launch_pattern<2>(p, sms, cyc, ns, smem); break;
    case 3: launch_pattern<3>(p, sms, cyc, ns, smem); break;
    default: return -1;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::vector<unsigned long long> hc(sms);
  cudaMemcpy(hc.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
  double mc = 0;
  for (int i = 0; i < sms; ++i) mc += hc[i];
  return mc / (sms * (double)p.iters);
}

// which operand layout property costs the N=256 MMA 171 instead of 128 cycles?
static void run_layout_sweep(int sms, unsigned long long* cyc, unsigned long long* ns, int smem) {
  printf("\n%-5s %-7s %-6s %-6s %-6s %-4s | %10s\n", "N", "a_lbo", "a_sbo", "a_off", "b_lbo", "rot", "cyc/MMA");
  for (int n : {256, 128, 64}) {
    for (int rot : {4, 1}) {
      for (int b_lbo : {n * 16, 8192}) {
        for (int a_lbo : {8192, 16384 - 8192 + 2880, 2880, 5184, 2048}) {
          for (int a_sbo : {256, 128, 160, 288}) {
            if (a_lbo == 2048 && a_sbo != 128) continue;                 // dense 128-row tile: k-slices back to back
            if ((a_lbo == 2880 && a_sbo != 160) || (a_lbo == 5184 && a_sbo != 288)) continue;   // the halo tiles of modconv_halo
            for (int a_off : {0, 176}) {
              if (a_off && a_sbo != 160) continue;
              Pattern p = {1, {n}, {0}, {0}, {0}, {0}, 8000, a_lbo, a_sbo, b_lbo, rot, a_off};
              const double c = time_pattern(p, sms, cyc, ns, smem);
              printf("%-5d %-7d %-6d %-6d %-6d %-4d | %10.2f\n", n, a_lbo, a_sbo, a_off, b_lbo, rot, c);
            }
          }
        }
      }
    }
  }
}

// Is the N=256 stream POWER-bound with realistic operands?  Same MMA stream, constant vs random data, short vs long runs.
static void run_power(int sms, unsigned long long* cyc, unsigned long long* ns, int smem) {
  printf("\n%-6s %-6s %-9s | %10s %10s %10s\n", "N", "data", "MMAs", "ns/MMA", "cyc/MMA", "TF/s chip");
  for (int n : {256, 128}) {
    for (int rnd : {0, 1}) {
      for (int it : {8000, 400000}) {
        Pattern p = {1, {n}, {0}, {0}, {0}, {0}, it, 0, 0, 0, 4, 0, rnd};
        launch_pattern<1>(p, sms, cyc, ns, smem);
        if (cudaDeviceSynchronize() != cudaSuccess) return;
        std::vector<unsigned long long> hc(sms), hn(sms);
        cudaMemcpy(hc.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hn.data(), ns, sms * 8, cudaMemcpyDeviceToHost);
        double mc = 0, mn = 0;
        for (int i = 0; i < sms; ++i) {
          mc += hc[i];
          mn += hn[i];
        }
        mc /= sms * (double)it;
        mn /= sms * (double)it;
        printf("%-6d %-6s %-9d | %10.2f %10.2f %10.1f\n", n, rnd ? "random" : "const", it, mn, mc, 2.0 * 128 * 16 * n / mn * sms / 1e3);
      }
    }
  }
}

static void run_patterns(int sms, unsigned long long* cyc, unsigned long long* ns, int smem) {
  struct Named { const char* name; Pattern p; };
  const int it = 8000;
  // fields: np, n[], col[], bslab[], brow[], aidx[]
  std::vector<Named> ps = {
      {"N=256 stream, one accumulator", {1, {256}, {0}, {0}, {0}, {0}, it}},
      {"N=256 x2: same A, two B slabs, two accumulators", {2, {256, 256}, {0, 256}, {0, 1}, {0, 0}, {0, 0}, it}},
      {"N=256 x2: two A, same B, two accumulators (MT=2)", {2, {256, 256}, {0, 256}, {0, 0}, {0, 0}, {0, 1}, it}},
      {"N=256 x3 lh hl hh (current halo<256,1> order)", {3, {256, 256, 256}, {0, 0, 0}, {0, 1, 0}, {0, 0, 0}, {1, 0, 0}, it}},
      {"N=256 x3 hl hh lh (A shared, then B shared)", {3, {256, 256, 256}, {0, 0, 0}, {1, 0, 0}, {0, 0, 0}, {0, 0, 1}, it}},
      {"N=128 stream, one accumulator", {1, {128}, {0}, {0}, {0}, {0}, it}},
      {"N=128 x2: same A, B halves of one slab, acc halves (N=256 split)", {2, {128, 128}, {0, 128}, {0, 0}, {0, 128}, {0, 0}, it}},
      {"N=128 x2: same A, same acc, two B (K-like)", {2, {128, 128}, {0, 0}, {0, 1}, {0, 0}, {0, 0}, it}},
      {"N=128 x2: two A, same B, two acc (MT=2 interleave)", {2, {128, 128}, {0, 128}, {0, 0}, {0, 0}, {0, 1}, it}},
      {"N=128 x2: two A, two B, two acc", {2, {128, 128}, {0, 128}, {0, 1}, {0, 0}, {0, 1}, it}},
      {"N=128 x2: same A, same B, two acc", {2, {128, 128}, {0, 128}, {0, 0}, {0, 0}, {0, 0}, it}},
      {"3 products as N=128 halves: lh lh' hl hl' hh hh'", {6, {128, 128, 128, 128, 128, 128}, {0, 128, 0, 128, 0, 128}, {0, 0, 1, 1, 0, 0}, {0, 128, 0, 128, 0, 128}, {1, 1, 0, 0, 0, 0}, it}},
      {"halo<128,2> new order: lh0 lh1 hl0 hl1 hh0 hh1", {6, {128, 128, 128, 128, 128, 128}, {0, 128, 0, 128, 0, 128}, {0, 0, 1, 1, 0, 0}, {0, 0, 0, 0, 0, 0}, {1, 3, 0, 2, 0, 2}, it}},
      {"halo<128,2> old order: lh0 hl0 hh0 lh1 hl1 hh1", {6, {128, 128, 128, 128, 128, 128}, {0, 0, 0, 128, 128, 128}, {0, 1, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 0}, {1, 0, 0, 3, 2, 2}, it}},
      {"scatter 256,128,128,64 (4 A tiles, one acc)", {4, {256, 128, 128, 64}, {0, 0, 64, 64}, {0, 1, 2, 3}, {0, 0, 0, 0}, {0, 1, 2, 3}, it}},
      {"scatter as N<=128 pieces alternating acc halves", {6, {128, 128, 128, 64, 64, 64}, {0, 128, 0, 128, 64, 64}, {0, 0, 1, 2, 2, 3}, {0, 128, 0, 0, 64, 0}, {0, 0, 1, 2, 2, 3}, it}},
      {"concat NT=64 x2 sub-tiles: 128 128 64 64", {4, {128, 128, 64, 64}, {0, 128, 0, 128}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 2, 1, 3}, it}},
      {"N=64 stream", {1, {64}, {0}, {0}, {0}, {0}, it}},
      {"N=64 x2: two A, same B, two acc", {2, {64, 64}, {0, 64}, {0, 0}, {0, 0}, {0, 1}, it}},
      {"N=64 x4: four A, same B, four acc", {4, {64, 64, 64, 64}, {0, 64, 128, 192}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 1, 2, 3}, it}},
      {"N=192 stream (3 dx taps x 64)", {1, {192}, {0}, {0}, {0}, {0}, it}},
      {"N=192 x2: two A, same B, two acc", {2, {192, 192}, {0, 256}, {0, 0}, {0, 0}, {0, 1}, it}},
  };
  printf("\n%-66s | %10s %11s %12s %10s\n", "pattern (operands rotate over 4 A tiles / 4 B slabs)", "ns/pattern", "cyc/pattern", "cyc/256cols", "TF/s chip");
  for (auto& nm : ps) {
    switch (nm.p.np) {
      case 1: launch_pattern<1>(nm.p, sms, cyc, ns, smem); break;
      case 2: launch_pattern<2>(nm.p, sms, cyc, ns, smem); break;
      case 3: launch_pattern<3>(nm.p, sms, cyc, ns, smem); break;
      case 4: launch_pattern<4>(nm.p, sms, cyc, ns, smem); break;
      case 6: launch_pattern<6>(nm.p, sms, cyc, ns, smem); break;
      default: continue;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s failed: %s\n", nm.name, cudaGetErrorString(e));
      return;
    }
    std::vector<unsigned long long> hc(sms), hn(sms);
    cudaMemcpy(hc.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hn.data(), ns, sms * 8, cudaMemcpyDeviceToHost);
    double mc = 0, mn = 0;
    for (int i = 0; i < sms; ++i) {
      mc += hc[i];
      mn += hn[i];
    }
    mc /= sms * (double)nm.p.iters;
    mn /= sms * (double)nm.p.iters;
    int cols = 0;
    for (int k = 0; k < nm.p.np; ++k) cols += nm.p.n[k];
    printf("%-66s | %10.2f %11.2f %12.2f %10.1f\n", nm.name, mn, mc, mc * 256.0 / cols, 2.0 * 128 * 16 * cols / mn * sms / 1e3);
  }
}

int main(int argc, char** argv) {
  const int smem = 1024 + 160 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long *cyc, *ns;
  float* sink;
  cudaMalloc(&cyc, sms * 8);
  cudaMalloc(&ns, sms * 8);
  cudaMalloc(&sink, 4);
  if (argc > 1 && argv[1][0] == 'p') {
    run_patterns(sms, cyc, ns, smem);
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'w') {
    run_power(sms, cyc, ns, smem);
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'l') {
    run_layout_sweep(sms, cyc, ns, smem);
    return 0;
  }
  std::vector<Case> cases;
  const int iters = 20000;
  for (int n : {64, 128, 192, 256}) {
    for (int acc : {1, 2, 4}) {
      if (acc * n > 512) continue;
      for (int arot : {1, 4}) {
        for (int brot : {1, 4}) {
          cases.push_back({n, 0, 288, 0, 0, iters, arot, brot, acc});
        }
      }
    }
  }
  // the concat pair of modconv_halo NT=64 (N=128 then N=64 on two sub-tiles), distinct operands
  cases.push_back({128, 0, 288, 64, 0, iters, 4, 1, 1});
  // shared-memory pressure on the distinct-operand cases
  for (int n : {64, 128, 256}) cases.push_back({n, 0, 288, 0, 1, iters, 4, 1, n == 256 ? 2 : 4});
  printf("%-5s %-5s %-5s %-5s %-4s %-4s | %10s %10s %10s\n", "N", "N2", "arot", "brot", "acc", "lsu", "ns/iter", "cyc/iter", "TF/s chip");
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
    printf("%-5d %-5d %-5d %-5d %-4d %-4d | %10.2f %10.2f %10.1f\n", c.n, c.n2, c.a_rot, c.b_rot, c.acc_rot, c.lsu, mn, mc, flop / mn * sms / 1e3);
  }
  return 0;
}

// Micro-benchmark 4: does the B-operand shared-memory layout change the rate of tcgen05.mma.cta_group::2 (M = 256)?
// A stays K-major no-swizzle (the conv kernel's shift trick needs it); B is either no-swizzle ([k8][rows][8], as
// packed today) or K-major SWIZZLE_128B (rows of 64 bf16 = 128 B, 8-row atoms of 1 KB).  Timing only: operands are zero.
#include <cstdio>
#include <cstdlib>
#include "../flope_b200/csrc/common.cuh"
using namespace flope;

template <int N, int NACC, int BSWZ, int CG>
__global__ void __launch_bounds__(128, 1) k(uint32_t a_lbo, int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    if (CG == 2) { tmem_alloc2(&tmem_ptr, 512); tmem_relinquish2(); } else { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  constexpr uint32_t IDESC = umma_idesc_bf16(CG == 2 ? 256 : 128, N);
  constexpr int NB = CG == 2 ? N / 2 : N;                       // B rows held by this CTA
  if (threadIdx.x < 32) {
    const uint32_t a_addr = smem_u32(base);
    const uint32_t b_addr = smem_u32(base) + 96 * 1024;          // 1 KB aligned
    const uint32_t desc_hi_a = (128u >> 4) | (1u << 14);         // SBO 128 B, version 1, no swizzle
    // swizzled B: SBO = 1024 B between 8-row atoms, layout type 2 (SWIZZLE_128B) in bits 61..63 -> bits 29..31 of the high word
    const uint32_t desc_hi_b = BSWZ ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : desc_hi_a;
    const uint32_t a_lo0 = ((a_lbo >> 4) << 16) + (a_addr >> 4);
    const uint32_t b_lo0 = BSWZ ? ((1u << 16) + (b_addr >> 4)) : ((((uint32_t)NB * 16u >> 4) << 16) + (b_addr >> 4));
    const uint32_t a_kstep = 2u * (a_lbo >> 4);
    constexpr uint32_t b_kstep = BSWZ ? 2u : 2u * NB;            // swizzled: 32 B along the 128-byte row per K = 16
    long long t0 = 0, t1 = 0;
    const uint32_t leader = (elect_one() && rank == 0) ? 1u : 0u;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          if (CG == 2) umma2_bf16_if(leader, tmem + a * N, a_lo0 + kk * a_kstep + a * 128u, desc_hi_a, b_lo0 + kk * b_kstep, desc_hi_b, IDESC, 1u);
          else umma_bf16_if(leader, tmem + a * N, a_lo0 + kk * a_kstep + a * 128u, desc_hi_a, b_lo0 + kk * b_kstep, desc_hi_b, IDESC, 1u);
        }
    }
    if (CG == 2) tc_commit2_if(leader, &bar); else tc_commit_if(leader, &bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) out_cycles[blockIdx.x] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (threadIdx.x < 32) { if (CG == 2) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int N, int NACC, int BSWZ, int CG>
void run(uint32_t a_lbo) {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k<N, NACC, BSWZ, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k<N, NACC, BSWZ, CG>, a_lbo, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 4.0 * NACC);
  printf("cta_group::%d M=%d N=%3d n_acc=%d B %-12s: %6.1f cycles/MMA -> %5.1f%% of the nominal rate\n", CG, CG * 128, N, NACC,
         BSWZ ? "SWIZZLE_128B" : "no-swizzle", per, 100.0 * (N / 2.0) / per);
  cudaFree(d);
}

int main() {
  run<256, 1, 0, 2>(2304); run<256, 1, 1, 2>(2304);
  run<128, 2, 0, 2>(5056); run<128, 2, 1, 2>(5056);
  run<64, 4, 0, 2>(10048); run<64, 4, 1, 2>(10048);
  run<256, 1, 0, 1>(2304); run<256, 1, 1, 1>(2304);
  run<64, 4, 0, 1>(10048); run<64, 4, 1, 1>(10048);
  return 0;
}

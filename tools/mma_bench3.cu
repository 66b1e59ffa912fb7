// Micro-benchmark 3: hardware rate of tcgen05.mma.cta_group::2 (M = 256 across a CTA pair), precomputed descriptors.
// Each CTA holds its 128 A rows and N/2 B rows (K-major, no swizzle), like the PAIR conv kernels.
#include <cstdio>
#include <cstdlib>
#include "../flope_b200/csrc/common.cuh"
using namespace flope;

template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) k(uint32_t a_lbo, int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(&tmem_ptr, 512); tmem_relinquish2(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  constexpr uint32_t IDESC = umma_idesc_bf16(256, N);
  constexpr int NB = N / 2;
  if (threadIdx.x < 32) {
    const uint32_t a_addr = smem_u32(base);
    const uint32_t b_addr = smem_u32(base) + 96 * 1024;
    // descriptors as in the conv kernel: constant high word, low word = base + compile-time multiples (uniform datapath)
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((a_lbo >> 4) << 16) + (a_addr >> 4);
    const uint32_t b_lo0 = (((uint32_t)NB * 16u >> 4) << 16) + (b_addr >> 4);
    const uint32_t a_kstep = 2u * (a_lbo >> 4);
    constexpr uint32_t b_kstep = 2u * NB;
    long long t0 = 0, t1 = 0;
    const uint32_t leader = (elect_one() && rank == 0) ? 1u : 0u;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          umma2_bf16_if(leader, tmem + a * N, a_lo0 + kk * a_kstep + a * 128u, desc_hi, b_lo0 + kk * b_kstep, desc_hi, IDESC, 1u);
    }
    tc_commit2_if(leader, &bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) out_cycles[blockIdx.x] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tmem, 512);
}

template <int N, int NACC>
void run(uint32_t a_lbo) {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k<N, NACC>, a_lbo, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 4.0 * NACC);
  printf("cta_group::2 M=256 N=%3d n_acc=%d: %6.1f cycles/MMA -> %5.1f%% of tensor peak (per SM: A 4 KB + B %.1f KB)\n", N, NACC, per,
         100.0 * (N / 2.0) / per, N / 2 * 32 / 1024.0);
  cudaFree(d);
}

int main() {
  run<64, 1>(10048); run<64, 4>(10048); run<64, 8>(10048);
  run<128, 1>(5056); run<128, 2>(5056); run<128, 4>(5056);
  run<256, 1>(2304); run<256, 2>(2304);
  run<32, 8>(10048);
  return 0;
}

// Micro-benchmark 2: hardware limit of tcgen05.mma issue with loop-invariant, precomputed descriptors.
#include <cstdio>
#include <cstdlib>
#include "../flope_b200/csrc/common.cuh"
using namespace flope;

template <int N, int NACC, int ORDER>   // ORDER 0: k outer, acc inner; 1: acc outer, k inner
__global__ void __launch_bounds__(128, 1) k(uint32_t a_lbo, int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, N);
  if (threadIdx.x < 32) {
    const uint32_t a_addr = smem_u32(base);
    const uint32_t b_addr = smem_u32(base) + 96 * 1024;
    uint64_t ad[4][NACC], bd[4];
    for (int kk = 0; kk < 4; ++kk) {
      bd[kk] = umma_desc(b_addr + 2 * kk * N * 16, N * 16, 128);
      for (int a = 0; a < NACC; ++a) ad[kk][a] = umma_desc(a_addr + a * 2048 + 2 * kk * a_lbo, a_lbo, 128);
    }
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (ORDER == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int a = 0; a < NACC; ++a) umma_bf16(tmem + a * N, ad[kk][a], bd[kk], IDESC, 1u);
        } else {
#pragma unroll
          for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem + a * N, ad[kk][a], bd[kk], IDESC, 1u);
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) out_cycles[blockIdx.x] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int ORDER>
void run(uint32_t a_lbo) {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k<N, NACC, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<N, NACC, ORDER><<<148, 128, 200 * 1024>>>(a_lbo, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 4.0 * NACC);
  printf("N=%3d n_acc=%d order=%s: %6.1f cycles/MMA -> %5.1f%% of tensor peak\n", N, NACC, ORDER ? "acc-outer" : "k-outer  ", per,
         100.0 * (N / 2.0) / per);
  cudaFree(d);
}

int main() {
  run<64, 1, 0>(10048); run<64, 2, 0>(10048); run<64, 4, 0>(10048); run<64, 4, 1>(10048); run<64, 8, 0>(10048); run<64, 8, 1>(10048);
  run<128, 1, 0>(5056); run<128, 2, 0>(5056); run<128, 2, 1>(5056); run<128, 4, 0>(5056); run<128, 4, 1>(5056);
  run<256, 1, 0>(2304); run<256, 2, 0>(2304); run<256, 2, 1>(2304);
  run<32, 8, 0>(10048); run<16, 8, 0>(10048);
  return 0;
}

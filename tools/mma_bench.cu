// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128) for several operand layouts.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../flope_b200/csrc/common.cuh"
using namespace flope;

__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// mode 0: no-swizzle, A LBO = a_lbo (bytes), SBO 128; B LBO = N*16, SBO 128
// mode 1: SW128 K-major: rows of 128 B, SBO = 1024, k-step = 32 B inside the row
template <int N, int M>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int mode, uint32_t a_lbo, int iters, int n_acc, int vary_a,
                                                           long long* out_cycles) {
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
  constexpr uint32_t IDESC = umma_idesc_bf16(M, N);
  if (threadIdx.x < 32) {
    const uint32_t a_addr = smem_u32(base);
    const uint32_t b_addr = smem_u32(base) + 96 * 1024;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (vary_a == 2) {          // accumulator outer, k inner
          for (int a = 0; a < n_acc; ++a) {
#pragma unroll 4
            for (int k = 0; k < 4; ++k) {
              uint64_t ad = desc_generic(a_addr + a * 2048 + 2 * k * a_lbo, a_lbo, 128, 0);
              uint64_t bd = desc_generic(b_addr + 2 * k * N * 16, N * 16, 128, 0);
              umma_bf16(tmem + a * N, ad, bd, IDESC, 1u);
            }
          }
          continue;
        }
#pragma unroll 4
        for (int k = 0; k < 4; ++k) {
          for (int a = 0; a < n_acc; ++a) {
            uint64_t ad, bd;
            const uint32_t shift = vary_a ? (uint32_t)((it % 9) * 16) : 0u;
            if (mode == 0) {
              ad = desc_generic(a_addr + shift + a * 2048 + 2 * k * a_lbo, a_lbo, 128, 0);
              bd = desc_generic(b_addr + 2 * k * N * 16, N * 16, 128, 0);
            } else if (mode == 1) {
              ad = desc_generic(a_addr + a * 16384 + k * 32, 16, 1024, 2);
              bd = desc_generic(b_addr + k * 32, 16, 1024, 2);
            } else {                 // mode 2: A fixed (weights, M rows), B varies per accumulator (pixels as N)
              ad = desc_generic(a_addr + 2 * k * 2048, 2048, 128, 0);
              bd = desc_generic(b_addr + a * (N * 16) + 2 * k * a_lbo, a_lbo, 128, 0);
            }
            umma_bf16(tmem + a * N, ad, bd, IDESC, 1u);
          }
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

template <int N, int M = 128>
void run(int mode, uint32_t a_lbo, int n_acc, int vary_a, const char* label) {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(mma_bench_kernel<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {148}) {
    mma_bench_kernel<N, M><<<grid, 128, 200 * 1024>>>(mode, a_lbo, iters, n_acc, vary_a, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", label, cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / (iters * 4.0 * n_acc);
    printf("%-40s M=%3d N=%3d grid=%3d n_acc=%d: %6.1f cycles/MMA -> %5.1f%% of tensor peak\n", label, M, N, grid, n_acc, per,
           100.0 * (M * N / 256.0) / per);
  }
  cudaFree(d);
}

int main() {
  run<64>(0, 10048, 1, 0, "pixM noswz k-outer");
  run<64>(0, 10048, 2, 0, "pixM noswz k-outer");
  run<64>(0, 10048, 4, 0, "pixM noswz k-outer");
  run<64>(0, 10048, 8, 0, "pixM noswz k-outer");
  run<64>(0, 10048, 4, 2, "pixM noswz acc-outer");
  run<128>(0, 5056, 1, 0, "pixM noswz k-outer");
  run<128>(0, 5056, 2, 0, "pixM noswz k-outer");
  run<128>(0, 5056, 4, 0, "pixM noswz k-outer");
  run<128>(0, 5056, 2, 2, "pixM noswz acc-outer");
  run<256>(0, 2304, 2, 0, "pixM noswz k-outer");
  run<256, 64>(2, 10048, 1, 0, "coutM=64 pixN=256 (weights as A)");
  run<256, 64>(2, 10048, 2, 0, "coutM=64 pixN=256 (weights as A)");
  run<128, 64>(2, 10048, 2, 0, "coutM=64 pixN=128 (weights as A)");
  run<128, 64>(2, 10048, 4, 0, "coutM=64 pixN=128 (weights as A)");
  run<256, 128>(2, 5056, 1, 0, "coutM=128 pixN=256 (weights as A)");
  run<256, 128>(2, 5056, 2, 0, "coutM=128 pixN=256 (weights as A)");
  run<192, 128>(2, 5056, 2, 0, "coutM=128 pixN=192 (weights as A)");
  run<128, 128>(2, 5056, 4, 0, "coutM=128 pixN=128 (weights as A)");
  return 0;
}

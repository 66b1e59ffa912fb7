/* Plain-C client of the flope_b200 C ABI (include/flope_b200.h): no Python, no torch, no C++.
 *
 *   c_abi_smoke <weights.bin> <crops.bin> <out_r9.bin>
 *
 * weights.bin : int32 n_tensors, then per tensor: int32 name_len, name bytes, int32 ndim, int64 shape[ndim], float32 data
 * crops.bin   : int32 n, int32 S, then n*3*S*S float32 (NCHW, [0,1])
 * out_r9.bin  : n*9 float32 raw pose vectors followed by n*9 float64 yaw-nullified rotations
 * tests/test_gpu_c_abi.py writes the inputs, runs this program and compares with the ctypes path bit for bit. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "flope_b200.h"

#define CHECK_FLOPE(x) do { int rc_ = (x); if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, flope_last_error()); return 2; } } while (0)
#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s weights.bin crops.bin out.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int32_t nt = 0;
  if (fread(&nt, 4, 1, f) != 1 || nt <= 0 || nt > 4096) return 1;
  flope_tensor_desc* td = (flope_tensor_desc*)calloc((size_t)nt, sizeof(*td));
  for (int i = 0; i < nt; ++i) {
    int32_t len = 0, nd = 0;
    if (fread(&len, 4, 1, f) != 1 || len <= 0 || len > 255) return 1;
    char* name = (char*)calloc((size_t)len + 1, 1);
    if (fread(name, 1, (size_t)len, f) != (size_t)len || fread(&nd, 4, 1, f) != 1 || nd < 0 || nd > 4) return 1;
    int64_t numel = 1;
    for (int d = 0; d < nd; ++d) {
      if (fread(&td[i].shape[d], 8, 1, f) != 1) return 1;
      numel *= td[i].shape[d];
    }
    float* data = (float*)malloc((size_t)numel * sizeof(float));
    if (fread(data, sizeof(float), (size_t)numel, f) != (size_t)numel) return 1;
    td[i].name = name; td[i].data = data; td[i].ndim = nd;
  }
  fclose(f);

  f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 1; }
  int32_t n = 0, S = 0;
  if (fread(&n, 4, 1, f) != 1 || fread(&S, 4, 1, f) != 1 || n <= 0 || S <= 0) return 1;
  const size_t in_elems = (size_t)n * 3 * S * S;
  float* h_in = (float*)malloc(in_elems * sizeof(float));
  if (fread(h_in, sizeof(float), in_elems, f) != in_elems) return 1;
  fclose(f);

  printf("flope_version %d, %d tensors, %d crops of %dx%d\n", flope_version(), nt, n, S, S);
  flope_engine* eng = NULL;
  CHECK_FLOPE(flope_engine_create(&eng, 0, n, S));
  CHECK_FLOPE(flope_engine_load_weights(eng, td, nt));
  float *d_in = NULL, *d_r9 = NULL;
  double* d_yaw = NULL;
  cudaStream_t st;
  CHECK_CUDA(cudaStreamCreate(&st));
  CHECK_CUDA(cudaMalloc((void**)&d_in, in_elems * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&d_r9, (size_t)n * 9 * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&d_yaw, (size_t)n * 9 * sizeof(double)));
  CHECK_CUDA(cudaMemcpyAsync(d_in, h_in, in_elems * sizeof(float), cudaMemcpyHostToDevice, st));
  CHECK_FLOPE(flope_posenet_forward(eng, d_in, n, d_r9, st));
  CHECK_FLOPE(flope_pose_head(eng, d_r9, n, NULL, d_yaw, st));
  float* h_r9 = (float*)malloc((size_t)n * 9 * sizeof(float));
  double* h_yaw = (double*)malloc((size_t)n * 9 * sizeof(double));
  CHECK_CUDA(cudaMemcpyAsync(h_r9, d_r9, (size_t)n * 9 * sizeof(float), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaMemcpyAsync(h_yaw, d_yaw, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaStreamSynchronize(st));                       /* the ABI never synchronises for the caller */
  /* error path: a NULL output is a code + message, not a crash */
  if (flope_posenet_forward(eng, d_in, n, NULL, st) != FLOPE_EINVAL) { fprintf(stderr, "expected FLOPE_EINVAL\n"); return 4; }
  f = fopen(argv[3], "wb");
  if (!f) { perror(argv[3]); return 1; }
  fwrite(h_r9, sizeof(float), (size_t)n * 9, f);
  fwrite(h_yaw, sizeof(double), (size_t)n * 9, f);
  fclose(f);
  printf("launches in the last call: %d; r9[0] = %g %g %g\n", flope_engine_last_launches(eng), h_r9[0], h_r9[1], h_r9[2]);
  flope_engine_destroy(eng);
  cudaFree(d_in); cudaFree(d_r9); cudaFree(d_yaw);
  cudaStreamDestroy(st);
  return 0;
}

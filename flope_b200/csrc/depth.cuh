// flope_b200: the depth branch of get_flower_poses on the device (sm_100a; HBM-bound byte/float work).
//
// Replaces get_depth_value (sunflower/utils/image_manipulation.py:39-96), called at
// sunflower/predictor/pose_predictor.py:118-121 and fast_pose_predictor.py:90-93:
//   good = (depth > near) & (depth < far);  seg = (mask > 128) & good;  seg = erode(seg, ellipse 10x10)
//   per box: values = depth[hmin:hmax, wmin:wmax][seg[...]] * 1000 (mm);  mean, count (>= 50 px = reliable)
// Two kernels: the validity mask + elliptical erosion for the whole frame (shared-memory tile with per-row
// prefix counts, so each structuring-element row is two look-ups), then one CTA per box for the masked sum.
#pragma once
#include "common.cuh"

namespace flope {

constexpr int kErodeMaxK = 31;
constexpr int kErodeTileW = 64, kErodeTileH = 16;

struct DepthParams {
  const void* depth;          // (H,W) float32 metres (dtype 0) or uint16 raw (dtype 1, metres = raw / div)
  int dtype;
  float div;
  const uint8_t* mask;        // (H,W) u8, 0 / 255
  int H, W;
  float near_plane, far_plane;
  int k;                      // structuring element side (the reference uses 10), anchor (k/2, k/2)
  int j1[kErodeMaxK], j2[kErodeMaxK];   // per element row: half-open span of ones (cv2 MORPH_ELLIPSE)
  uint8_t* eroded;            // (H,W) out: 1 where the eroded mask is set
};

__device__ __forceinline__ float depth_metres(const DepthParams& p, long long i) {
  if (p.dtype == 1) return __fdiv_rn((float)reinterpret_cast<const uint16_t*>(p.depth)[i], p.div);   // astype(float32) / div
  return reinterpret_cast<const float*>(p.depth)[i];
}

// Output tile 64x16; the (64+k) x (16+k) input region becomes inclusive per-row prefix counts in shared memory.
// Pixels outside the image never erode (cv2's default border value for erosion is +inf).
__global__ void __launch_bounds__(256) erode_valid_kernel(const __grid_constant__ DepthParams p) {
  __shared__ uint8_t s_pre[kErodeTileH + kErodeMaxK][kErodeTileW + kErodeMaxK + 1];
  const int a = p.k >> 1;
  const int x0 = blockIdx.x * kErodeTileW - a, y0 = blockIdx.y * kErodeTileH - a;
  const int RW = kErodeTileW + p.k, RH = kErodeTileH + p.k;
  for (int i = threadIdx.x; i < RW * RH; i += blockDim.x) {
    const int ry = i / RW, rx = i - ry * RW;
    const int gy = y0 + ry, gx = x0 + rx;
    uint8_t v = 1;
    if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
      const long long g = (long long)gy * p.W + gx;
      const float d = depth_metres(p, g);
      v = (p.mask[g] > 128 && d > p.near_plane && d < p.far_plane) ? 1 : 0;
    }
    s_pre[ry][rx + 1] = v;
  }
  __syncthreads();
  for (int ry = threadIdx.x; ry < RH; ry += blockDim.x) {
    uint8_t acc = 0;
    s_pre[ry][0] = 0;
    for (int rx = 1; rx <= RW; ++rx) { acc += s_pre[ry][rx]; s_pre[ry][rx] = acc; }
  }
  __syncthreads();
  const int tx = threadIdx.x & (kErodeTileW - 1);
  for (int ty = threadIdx.x / kErodeTileW; ty < kErodeTileH; ty += blockDim.x / kErodeTileW) {
    const int gx = blockIdx.x * kErodeTileW + tx, gy = blockIdx.y * kErodeTileH + ty;
    if (gx >= p.W || gy >= p.H) continue;
    bool all = true;
    for (int i = 0; i < p.k; ++i) {
      const int j1 = p.j1[i], j2 = p.j2[i];
      all = all && ((int)(uint8_t)(s_pre[ty + i][tx + j2] - s_pre[ty + i][tx + j1]) == j2 - j1);
    }
    p.eroded[(long long)gy * p.W + gx] = all ? 1 : 0;
  }
}

// kBoxSplit CTAs per box (row stripes): masked sum of depth in millimetres (float32 multiply like the reference) and
// the pixel count.  The sum is accumulated as a 64-bit fixed-point integer in units of 2^-20 mm (exact for every float32
// value of 8 mm and more, rounded to 1e-6 mm below; 2 M pixels x 2^32 stay far below 2^63), so the atomics across
// CTAs are order-independent and the result is deterministic; box_depth_finish_kernel turns (sum, count) into metres.
constexpr int kBoxSplit = 16;
constexpr int kDepthFixShift = 20;

__global__ void __launch_bounds__(256) box_depth_kernel(const __grid_constant__ DepthParams p, const int32_t* __restrict__ boxes,
                                                        unsigned long long* __restrict__ acc /* n x 2: sum, count */) {
  __shared__ unsigned long long s_sum[8];
  __shared__ int s_cnt[8];
  const int b = blockIdx.x;
  const int wmin = max(boxes[4 * b], 0), hmin = max(boxes[4 * b + 1], 0);
  const int wmax = min(boxes[4 * b + 2], p.W), hmax = min(boxes[4 * b + 3], p.H);
  const int bw = max(wmax - wmin, 0), bh = max(hmax - hmin, 0);
  const int rows = (bh + kBoxSplit - 1) / kBoxSplit;                 // this CTA's stripe of box rows
  const int r0 = blockIdx.y * rows, r1 = min(bh, r0 + rows);
  unsigned long long sum = 0;
  int cnt = 0;
  for (int i = threadIdx.x; i < bw * max(r1 - r0, 0); i += blockDim.x) {
    const int y = r0 + i / bw, x = i % bw;
    const long long g = (long long)(hmin + y) * p.W + wmin + x;
    if (p.eroded[g]) {
      const float mm = __fmul_rn(depth_metres(p, g), 1000.0f);
      sum += (unsigned long long)__double2ll_rn((double)mm * (double)(1 << kDepthFixShift));
      ++cnt;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long ts = 0;
    int tc = 0;
    for (int w = 0; w < 8; ++w) { ts += s_sum[w]; tc += s_cnt[w]; }
    if (tc) {
      atomicAdd(&acc[2 * b], ts);
      atomicAdd(&acc[2 * b + 1], (unsigned long long)tc);
    }
  }
}

// val = float32(mean in mm) / 1000 in metres (0 when no pixel), count decides reliability (>= 50 in the reference)
__global__ void box_depth_finish_kernel(const unsigned long long* __restrict__ acc, int n, double* __restrict__ val,
                                        int32_t* __restrict__ count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const unsigned long long s = acc[2 * b], c = acc[2 * b + 1];
  count[b] = (int32_t)c;
  const double mean_mm = c ? ((double)s / (double)(1 << kDepthFixShift)) / (double)c : 0.0;
  val[b] = c ? (double)__fdiv_rn((float)mean_mm, 1000.0f) : 0.0;
}

}  // namespace flope

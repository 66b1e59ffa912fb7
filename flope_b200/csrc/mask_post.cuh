// flope_b200: YOLO-seg post-processing on the device.
//
// Replaces the tensor -> numpy -> cv2 round trip of FastPosePredictor.get_bbox_mask
// (sunflower/predictor/fast_pose_predictor.py:44-57, dup scripts/generate_metrics_utils.py:114-127):
//   mask = clip(sum(masks, axis=0), 0, 1) * 255 -> uint8;  mask = cv2.resize(mask, (W, H))   # INTER_LINEAR
// so the full-resolution mask never leaves the GPU between the detector and the ROI / depth kernels.
// The resize reproduces cv2's uint8 INTER_LINEAR fixed-point arithmetic bit for bit (same coefficient and
// rounding rules as roi_crop.cuh's linear_coefs; one channel, arbitrary source and destination sizes).
#pragma once
#include "common.cuh"
#include "roi_crop.cuh"

namespace flope {

// (n,h,w) float instance masks -> (h,w) uint8 union: uint8(clip(sum, 0, 1) * 255), fp32 sum in instance order
__global__ void merge_instance_masks_kernel(const float* __restrict__ masks, int n, long long hw, uint8_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < n; ++k) s = __fadd_rn(s, masks[(long long)k * hw + i]);
    s = fminf(fmaxf(s, 0.f), 1.f);
    out[i] = (uint8_t)(int)__fmul_rn(s, 255.f);
  }
}

// per destination index: source index and the two 11-bit coefficients (cv2's HResizeLinear / VResizeLinear tables)
__global__ void linear_table_kernel(int dst, int src, int vertical, int* __restrict__ ofs, short* __restrict__ coef) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dst) return;
  const double scale = 1.0 / ((double)dst / (double)src);
  short ic[2];
  int s;
  linear_coefs(d, scale, src, vertical != 0, s, ic);
  ofs[d] = s;
  coef[2 * d] = ic[0];
  coef[2 * d + 1] = ic[1];
}

// single-channel uint8 bilinear resize, one thread per destination pixel
__global__ void resize_linear_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh, int dw,
                                        const int* __restrict__ xofs, const short* __restrict__ xcoef,
                                        const int* __restrict__ yofs, const short* __restrict__ ycoef) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const int sx = xofs[x], sy = yofs[y];
  const int xa = min(max(sx, 0), sw - 1), xb = min(max(sx + 1, 0), sw - 1);
  const int ya = min(max(sy, 0), sh - 1), yb = min(max(sy + 1, 0), sh - 1);
  const int cx0 = xcoef[2 * x], cx1 = xcoef[2 * x + 1];
  const int b0 = ycoef[2 * y], b1 = ycoef[2 * y + 1];
  const uint8_t* ra = src + (long long)ya * sw;
  const uint8_t* rb = src + (long long)yb * sw;
  const int h0 = ((int)ra[xa] * cx0 + (int)ra[xb] * cx1) >> 4;
  const int h1 = ((int)rb[xa] * cx0 + (int)rb[xb] * cx1) >> 4;
  const int v = (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2;
  dst[(long long)y * dw + x] = (uint8_t)min(max(v, 0), 255);
}

}  // namespace flope

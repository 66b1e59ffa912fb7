// flope_b200: fused ROI crop / resize / mask / normalise kernel over uint8 frames (sm_100a).
//
// Replaces the per-box Python loop of the reference
// (sunflower/predictor/pose_predictor.py:138-153, fast_pose_predictor.py:108-123):
//   crop = frame[ymin:ymax, xmin:xmax]; cv2.resize(crop,(S,S),interp) for image and mask;
//   out  = float32( crop * (mask/255.0) / 255.0 )
// The resize reproduces cv2's uint8 fixed-point arithmetic bit for bit (11-bit coefficients,
// int32 accumulation, the same rounding and saturation; oracle/resize.py is the specification):
//   interp 1: INTER_LANCZOS4 (8x8 taps, replicate border of the crop)  - the reference's mode
//   interp 0: INTER_LINEAR   (2x2 taps)                                 - the benchmark mode
// Output formats:
//   0: float32 NCHW (B,3,S,S)          - the reference's tensor layout
//   1: bf16 space-to-depth blocked-pixel - what the tcgen05 stem reads (pointwise.cuh ingest layout)
//
// Thread = one output column; it marches down a strip of output rows keeping a sliding
// window of horizontally-filtered source rows in registers (separable filter, each source
// row is filtered once per strip), so the kernel is bound by the output write.
#pragma once
#include "common.cuh"

namespace flope {

struct RoiParams {
  const uint8_t* frames;      // (n_frames, H, W, 3) u8
  long long frame_stride;     // bytes between frames
  const uint8_t* masks;       // (n_frames, H, W) u8 or nullptr
  long long mask_stride;
  int H, W;
  const int32_t* boxes;       // (n, 5): frame, xmin, ymin, xmax, ymax (already squarified + in-frame)
  int n;
  int S;                      // output side
  int out_fmt;
  void* out;
  Geom g;                     // fmt 1 geometry (S/2 grid)
  int rows_per_strip;
};

constexpr int kRoiMaxStripRows = 128;

// cv2 interpolateLanczos4 + fixed-point conversion, for destination index d (see oracle/resize.py)
__device__ inline void lanczos4_coefs(int d, double scale, int& s_out, short (&ic)[8]) {
  const double CV_PI_ = 3.1415926535897932384626433832795;
  float fx = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  const int s = (int)floorf(fx);
  fx -= (float)s;
  const double s45 = 0.70710678118654752440084436210485;
  const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  float c[8];
  float sum = 0.f;
  const float xb = __fadd_rn(fx, 3.0f);
  const double y0 = __dmul_rn(__dmul_rn(-(double)xb, CV_PI_), 0.25);
  const double s0 = sin(y0), c0 = cos(y0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float t = __fsub_rn(xb, (float)i);
    if (fabsf(t) >= 1e-6f) {
      const double y = __dmul_rn(__dmul_rn(-(double)t, CV_PI_), 0.25);
      c[i] = (float)__ddiv_rn(__dadd_rn(__dmul_rn(cs[i][0], s0), __dmul_rn(cs[i][1], c0)), __dmul_rn(y, y));
    } else {
      c[i] = 1e30f;
    }
    sum = __fadd_rn(sum, c[i]);
  }
  const float inv = __fdiv_rn(1.f, sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = __float2int_rn(__fmul_rn(__fmul_rn(c[i], inv), 2048.f));
    ic[i] = (short)max(-32768, min(32767, v));
  }
  s_out = s;
}

// cv2 INTER_LINEAR coefficients; horizontal axis clamps the fraction at the border, vertical does not
__device__ inline void linear_coefs(int d, double scale, int src, bool vertical, int& s_out, short (&ic)[2]) {
  float fx = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  int s = (int)floorf(fx);
  fx -= (float)s;
  if (!vertical) {
    if (s < 0) { fx = 0.f; s = 0; }
    if (s >= src - 1) { fx = 0.f; s = src - 1; }
  }
  const int v0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f));
  const int v1 = __float2int_rn(__fmul_rn(fx, 2048.f));
  ic[0] = (short)max(-32768, min(32767, v0));
  ic[1] = (short)max(-32768, min(32767, v1));
  s_out = s;
}

template <int TAPS, bool HAS_MASK>
__global__ void __launch_bounds__(256) roi_crop_kernel(const __grid_constant__ RoiParams p) {
  constexpr int NCH = HAS_MASK ? 4 : 3;
  constexpr int ORG = TAPS == 8 ? 3 : 0;          // window origin relative to floor(src coord)
  __shared__ int s_sy[kRoiMaxStripRows];
  __shared__ short s_icy[kRoiMaxStripRows][TAPS];

  const int crop = blockIdx.z;
  const int32_t* bx = p.boxes + (size_t)crop * 5;
  const int frame = bx[0], xmin = bx[1], ymin = bx[2];
  const int sw = bx[3] - bx[1], sh = bx[4] - bx[2];
  const int S = p.S;
  const int y_begin = blockIdx.y * p.rows_per_strip;
  const int y_end = min(S, y_begin + p.rows_per_strip);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (sw <= 0 || sh <= 0) return;                 // guarded by the host; never dereference an empty box

  const double scale_x = 1.0 / ((double)S / (double)sw);
  const double scale_y = 1.0 / ((double)S / (double)sh);
  for (int r = threadIdx.x; r < y_end - y_begin; r += blockDim.x) {
    short ic[TAPS];
    int sy;
    if constexpr (TAPS == 8) lanczos4_coefs(y_begin + r, scale_y, sy, ic);
    else linear_coefs(y_begin + r, scale_y, sh, true, sy, ic);
    s_sy[r] = sy;
#pragma unroll
    for (int j = 0; j < TAPS; ++j) s_icy[r][j] = ic[j];
  }
  __syncthreads();
  if (x >= S) return;

  // horizontal taps of this column: byte offsets into a source row, and coefficients
  int xo[TAPS];
  int icx[TAPS];
  {
    short ic[TAPS];
    int sx;
    if constexpr (TAPS == 8) lanczos4_coefs(x, scale_x, sx, ic);
    else linear_coefs(x, scale_x, sw, false, sx, ic);
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
      xo[j] = min(max(sx - ORG + j, 0), sw - 1);
      icx[j] = ic[j];
    }
  }
  const uint8_t* img = p.frames + (long long)frame * p.frame_stride + ((long long)ymin * p.W + xmin) * 3;
  const uint8_t* msk = HAS_MASK ? p.masks + (long long)frame * p.mask_stride + (long long)ymin * p.W + xmin : nullptr;

  int win[TAPS][NCH];     // horizontally filtered rows u .. u+TAPS-1 (u = unclamped source row index)
  int u = 0;
  bool primed = false;

  auto hrow = [&](int urow, int (&dst)[NCH]) {
    const int r = min(max(urow, 0), sh - 1);
    const uint8_t* row = img + (long long)r * p.W * 3;
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
      const uint8_t* px = row + xo[j] * 3;
      a0 += (int)__ldg(px) * icx[j];
      a1 += (int)__ldg(px + 1) * icx[j];
      a2 += (int)__ldg(px + 2) * icx[j];
      if (HAS_MASK) a3 += (int)__ldg(msk + (long long)r * p.W + xo[j]) * icx[j];
    }
    dst[0] = a0; dst[1] = a1; dst[2] = a2;
    if (HAS_MASK) dst[3] = a3;
  };

  for (int y = y_begin; y < y_end; ++y) {
    const int u_new = s_sy[y - y_begin] - ORG;
    if (!primed || u_new - u >= TAPS || u_new < u) {
#pragma unroll
      for (int j = 0; j < TAPS; ++j) hrow(u_new + j, win[j]);
      primed = true;
    } else {
      for (int s = u; s < u_new; ++s) {
#pragma unroll
        for (int j = 0; j < TAPS - 1; ++j) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) win[j][c] = win[j + 1][c];
        }
        hrow(s + TAPS, win[TAPS - 1]);
      }
    }
    u = u_new;

    int v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int r;
      if (TAPS == 8) {
        int acc = 0;
#pragma unroll
        for (int j = 0; j < TAPS; ++j) acc += win[j][c] * (int)s_icy[y - y_begin][j];
        r = (acc + (1 << 21)) >> 22;
      } else {
        const int b0 = s_icy[y - y_begin][0], b1 = s_icy[y - y_begin][1];
        r = (((b0 * (win[0][c] >> 4)) >> 16) + ((b1 * (win[1][c] >> 4)) >> 16) + 2) >> 2;
      }
      v[c] = min(max(r, 0), 255);
    }
    // float32((double(img) * (double(mask)/255.0)) / 255.0) == fp32-rounded (img*mask)/65025 for all
    // 65536 (img,mask) pairs (checked exhaustively in tests/test_oracle_resize.py and on the device)
    const int m = HAS_MASK ? v[3] : 255;
    const float f0 = __fdiv_rn((float)(v[0] * m), 65025.f);
    const float f1 = __fdiv_rn((float)(v[1] * m), 65025.f);
    const float f2 = __fdiv_rn((float)(v[2] * m), 65025.f);
    if (p.out_fmt == 0) {
      float* o = reinterpret_cast<float*>(p.out) + ((long long)crop * 3 * S + y) * S + x;
      o[0] = f0;
      o[(long long)S * S] = f1;
      o[2LL * S * S] = f2;
    } else {
      uint2 o;
      o.x = pack_bf16x2(f0, f1);
      o.y = pack_bf16x2(f2, 0.f);
      const long long pos = p.g.base + geom_pos(p.g, crop, y >> 1, x >> 1);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)(y & 1) * p.g.plane + pos) * 8 + (x & 1) * 4;
      *reinterpret_cast<uint2*>(dst) = o;
    }
  }
}

}  // namespace flope

// flope_b200: fused ROI crop / resize / mask / normalise kernels over uint8 frames (sm_100a).
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
// This header holds the arithmetic shared by every ROI kernel (coefficient tables, the exact normalise) and the
// generic fallback kernel roi_crop_kernel<TAPS,HAS_MASK>: one thread per output column reading its taps straight from
// global memory, used for frame widths that are not a multiple of 16 and output sides the streaming kernels do not
// cover.  The production kernels are in roi_stream.cuh.
#pragma once
#include "common.cuh"

#include <math.h>

#if defined(__CUDACC__)
#define FLOPE_HD __host__ __device__ __forceinline__
#else
#define FLOPE_HD inline
#endif

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

// ---------------------------------------------------------------------------------------------
// Arithmetic shared by the device kernels and the host emulation.  cv2 computes its coefficient tables with
// individually rounded float / double operations; the device versions use the _rn intrinsics so that ptxas
// cannot contract them into FMAs.
// ---------------------------------------------------------------------------------------------
FLOPE_HD double rn_dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
FLOPE_HD double rn_dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
FLOPE_HD double rn_dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
FLOPE_HD double rn_ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
FLOPE_HD float rn_fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
FLOPE_HD float rn_fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
FLOPE_HD float rn_fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
FLOPE_HD float rn_fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
FLOPE_HD int rint_to_int(float v) {
#ifdef __CUDA_ARCH__
  return __float2int_rn(v);
#else
  return (int)nearbyintf(v);      // default rounding mode: to nearest even, like cvRound
#endif
}
FLOPE_HD int imin(int a, int b) { return a < b ? a : b; }
FLOPE_HD int imax(int a, int b) { return a > b ? a : b; }
FLOPE_HD int iclamp(int v, int lo, int hi) { return imin(imax(v, lo), hi); }

// source coordinate of destination index d: (float)((d + 0.5) * scale - 0.5), floor and fraction (cv::resize)
FLOPE_HD int src_coord(int d, double scale, float& frac) {
  float fx = (float)rn_dsub(rn_dmul((double)d + 0.5, scale), 0.5);
  const int s = (int)floorf(fx);
  frac = rn_fsub(fx, (float)s);
  return s;
}
FLOPE_HD double axis_scale(int src, int dst) { return 1.0 / ((double)dst / (double)src); }

// cv2 interpolateLanczos4 + fixed-point conversion, for destination index d (see oracle/resize.py)
FLOPE_HD void lanczos4_coefs(int d, double scale, int& s_out, short (&ic)[8]) {
  const double CV_PI_ = 3.1415926535897932384626433832795;
  float fx;
  const int s = src_coord(d, scale, fx);
  const double s45 = 0.70710678118654752440084436210485;
  const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  float c[8];
  float sum = 0.f;
  const float xb = rn_fadd(fx, 3.0f);
  const double y0 = rn_dmul(rn_dmul(-(double)xb, CV_PI_), 0.25);
  const double s0 = sin(y0), c0 = cos(y0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float t = rn_fsub(xb, (float)i);
    if (fabsf(t) >= 1e-6f) {
      const double y = rn_dmul(rn_dmul(-(double)t, CV_PI_), 0.25);
      c[i] = (float)rn_ddiv(rn_dadd(rn_dmul(cs[i][0], s0), rn_dmul(cs[i][1], c0)), rn_dmul(y, y));
    } else {
      c[i] = 1e30f;
    }
    sum = rn_fadd(sum, c[i]);
  }
  const float inv = rn_fdiv(1.f, sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int v = rint_to_int(rn_fmul(rn_fmul(c[i], inv), 2048.f));
    ic[i] = (short)imax(-32768, imin(32767, v));
  }
  s_out = s;
}

// cv2 INTER_LINEAR coefficients; horizontal axis clamps the fraction at the border, vertical does not
FLOPE_HD void linear_coefs(int d, double scale, int src, bool vertical, int& s_out, short (&ic)[2]) {
  float fx;
  int s = src_coord(d, scale, fx);
  if (!vertical) {
    if (s < 0) { fx = 0.f; s = 0; }
    if (s >= src - 1) { fx = 0.f; s = src - 1; }
  }
  const int v0 = rint_to_int(rn_fmul(rn_fsub(1.f, fx), 2048.f));
  const int v1 = rint_to_int(rn_fmul(fx, 2048.f));
  ic[0] = (short)imax(-32768, imin(32767, v0));
  ic[1] = (short)imax(-32768, imin(32767, v1));
  s_out = s;
}

// float32((double(img) * (double(mask)/255.0)) / 255.0) == correctly rounded fp32 (img*mask)/65025 for all
// 65536 (img,mask) pairs (tests/test_oracle_resize.py).  One reciprocal multiply plus one exact-residual
// correction step gives that correctly rounded quotient for every integer numerator 0..65025; the device
// result is checked exhaustively against the oracle table in tests/test_gpu_roi.py.
FLOPE_HD float normalise_u8(int img, int mask) {
  const float x = (float)(img * mask);
  const float r = 1.0f / 65025.0f;
  const float q = x * r;
  const float e = fmaf(-q, 65025.0f, x);
  return fmaf(e, r, q);
}

FLOPE_HD uint32_t f32_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
// bf16 bit pattern of a finite non-negative float, round to nearest even (== cvt.rn.bf16.f32)
FLOPE_HD uint32_t bf16_bits_rn(float f) {
  const uint32_t u = f32_bits(f);
  return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}

struct U32x2 { uint32_t x, y; };
struct alignas(16) U32x4 { uint32_t x, y, z, w; };

#ifdef __CUDACC__
// =============================================================================================
// generic fallback: thread = one output column reading its taps straight from global memory
// =============================================================================================
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
  if (sw <= 0 || sh <= 0) {                       // empty box (the Python layer rejects these): the crop is defined as zeros
    if (x < S)
      for (int y = y_begin; y < y_end; ++y) {
        if (p.out_fmt == 0) {
          float* o = reinterpret_cast<float*>(p.out) + ((long long)crop * 3 * S + y) * S + x;
          o[0] = 0.f; o[(long long)S * S] = 0.f; o[2LL * S * S] = 0.f;
        } else {
          const long long pos = p.g.base + geom_pos(p.g, crop, y >> 1, x >> 1);
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)(y & 1) * p.g.plane + pos) * 8 + (x & 1) * 4;
          *reinterpret_cast<uint2*>(dst) = make_uint2(0u, 0u);
        }
      }
    return;
  }

  const double scale_x = axis_scale(sw, S);
  const double scale_y = axis_scale(sh, S);
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
    const int m = HAS_MASK ? v[3] : 255;
    const float f0 = normalise_u8(v[0], m), f1 = normalise_u8(v[1], m), f2 = normalise_u8(v[2], m);
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
#endif  // __CUDACC__

}  // namespace flope

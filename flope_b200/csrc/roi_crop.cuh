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

// float32((double(img) * (double(mask)/255.0)) / 255.0) == correctly rounded fp32 (img*mask)/65025 for all
// 65536 (img,mask) pairs (tests/test_oracle_resize.py).  One reciprocal multiply plus one exact-residual
// correction step gives that correctly rounded quotient for every integer numerator 0..65025; the device
// result is checked exhaustively against the oracle table in tests/test_gpu_roi.py.
__device__ __forceinline__ float normalise_u8(int img, int mask) {
  const float x = (float)(img * mask);
  const float r = 1.0f / 65025.0f;
  const float q = x * r;
  const float e = fmaf(-q, 65025.0f, x);
  return fmaf(e, r, q);
}

// ---------------------------------------------------------------------------------------------
// Bilinear (cv2.INTER_LINEAR, uint8-exact) specialisation: the benchmark mode.
// Thread = one output column marching down a strip.  Two horizontally filtered source rows are kept
// (pre-shifted by 4 as cv2's vertical pass wants them); the vertical pass is two IMAD.HI per channel:
//   ((b*(S>>4))>>16) == mulhi(b<<16, S>>4).   The result never leaves [0,255], so no saturation is needed.
// ---------------------------------------------------------------------------------------------
template <bool HAS_MASK>
__global__ void __launch_bounds__(256, 4) roi_bilinear_kernel(const __grid_constant__ RoiParams p) {
  // thread = two adjacent output columns (2q, 2q+1) marching down the strip: one thread produces both
  // 16-byte halves of a space-to-depth stem pixel, and the row bookkeeping / coefficient fetches are
  // shared by the two columns.
  constexpr int NCH = HAS_MASK ? 4 : 3;
  __shared__ int s_sy[kRoiMaxStripRows];
  __shared__ uint32_t s_b0[kRoiMaxStripRows];
  __shared__ uint32_t s_b1[kRoiMaxStripRows];
  __shared__ float s_lut[256];             // normalise_u8(i, 255): the value of an unmasked / fully masked-in pixel

  const int crop = blockIdx.z;
  const int32_t* bx = p.boxes + (size_t)crop * 5;
  const int frame = bx[0], xmin = bx[1], ymin = bx[2];
  const int sw = bx[3] - bx[1], sh = bx[4] - bx[2];
  const int S = p.S;
  const int y_begin = blockIdx.y * p.rows_per_strip;           // even
  const int y_end = min(S, y_begin + p.rows_per_strip);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (sw <= 0 || sh <= 0) return;

  const double scale_x = 1.0 / ((double)S / (double)sw);
  const double scale_y = 1.0 / ((double)S / (double)sh);
  for (int r = threadIdx.x; r < y_end - y_begin; r += blockDim.x) {
    short ic[2];
    int sy;
    linear_coefs(y_begin + r, scale_y, sh, true, sy, ic);
    s_sy[r] = sy;
    s_b0[r] = (uint32_t)(int)ic[0] << 16;
    s_b1[r] = (uint32_t)(int)ic[1] << 16;
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = normalise_u8(i, 255);
  __syncthreads();
  if (2 * q >= S) return;

  int cx0[2], cx1[2];
  uint32_t oa[2], ob[2];                   // source columns of the two horizontal taps of each output column
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    short ic[2];
    int sx;
    linear_coefs(2 * q + k, scale_x, sw, false, sx, ic);
    cx0[k] = ic[0]; cx1[k] = ic[1];
    oa[k] = (uint32_t)min(max(sx, 0), sw - 1);
    ob[k] = (uint32_t)min(max(sx + 1, 0), sw - 1);
  }
  const uint8_t* img = p.frames + (long long)frame * p.frame_stride + ((long long)ymin * p.W + xmin) * 3;
  const uint8_t* msk = HAS_MASK ? p.masks + (long long)frame * p.mask_stride + (long long)ymin * p.W + xmin : nullptr;
  const uint32_t row_bytes = (uint32_t)p.W * 3u;

  uint32_t h0[2][NCH], h1[2][NCH];         // (horizontally filtered row) >> 4 for source rows u and u+1, per column
  // The tap bytes of a source row are loaded (load_raw) and filtered (filt) in two steps so that the loads of the
  // row the NEXT output row needs are in flight while this output row is computed (software pipelining: the
  // kernel is latency bound, not issue bound).
  struct Raw { uint32_t a[2][NCH], b[2][NCH]; };
  auto load_raw = [&](int urow, Raw& w) {
    const uint32_t r = (uint32_t)min(max(urow, 0), sh - 1);
    const uint8_t* row = img + r * row_bytes;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint8_t* pa = row + oa[k] * 3u;
      const uint8_t* pb = row + ob[k] * 3u;
#pragma unroll
      for (int c = 0; c < 3; ++c) { w.a[k][c] = __ldg(pa + c); w.b[k][c] = __ldg(pb + c); }
    }
    if (HAS_MASK) {
      const uint8_t* mrow = msk + r * (uint32_t)p.W;
#pragma unroll
      for (int k = 0; k < 2; ++k) { w.a[k][NCH - 1] = __ldg(mrow + oa[k]); w.b[k][NCH - 1] = __ldg(mrow + ob[k]); }
    }
  };
  auto filt = [&](const Raw& w, uint32_t (&dst)[2][NCH]) {
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int c = 0; c < NCH; ++c) dst[k][c] = (uint32_t)((int)w.a[k][c] * cx0[k] + (int)w.b[k][c] * cx1[k]) >> 4;
  };

  int u = s_sy[0];
  Raw nx;
  load_raw(u, nx);
  filt(nx, h0);
  load_raw(u + 1, nx);
  filt(nx, h1);
  int nx_row = u - 1;                      // source row whose tap bytes `nx` holds unfiltered (none yet)
  const long long plane_sz = (long long)S * S;
  float* o32 = reinterpret_cast<float*>(p.out) + ((long long)crop * 3 * S + y_begin) * S + 2 * q;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(p.out) +
                       ((long long)p.g.base + geom_pos(p.g, crop, y_begin >> 1, q)) * 8;
  const long long o16_plane = p.g.plane * 8;
  const long long o16_row = (long long)p.g.Wp * 8;
  for (int y = y_begin; y < y_end; ++y) {
    const int u_new = s_sy[y - y_begin];
    if (u_new != u) {
      if (u_new == u + 1) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int c = 0; c < NCH; ++c) h0[k][c] = h1[k][c];
        if (nx_row != u_new + 1) load_raw(u_new + 1, nx);     // normally prefetched one output row ago
        filt(nx, h1);
      } else {
        load_raw(u_new, nx);
        filt(nx, h0);
        load_raw(u_new + 1, nx);
        filt(nx, h1);
      }
      u = u_new;
    }
    if (y + 1 < y_end && s_sy[y + 1 - y_begin] == u + 1) {    // the next output row slides the window by one source row
      load_raw(u + 2, nx);
      nx_row = u + 2;
    }
    const uint32_t b0 = s_b0[y - y_begin], b1 = s_b1[y - y_begin];
    float f[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t v[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c) v[c] = (__umulhi(b0, h0[k][c]) + (__umulhi(b1, h1[k][c]) + 2u)) >> 2;
      const uint32_t m = HAS_MASK ? v[3] : 255u;
      if (m == 255u) {
        f[k][0] = s_lut[v[0]]; f[k][1] = s_lut[v[1]]; f[k][2] = s_lut[v[2]];
      } else if (m == 0u) {
        f[k][0] = f[k][1] = f[k][2] = 0.f;
      } else {
        f[k][0] = normalise_u8((int)v[0], (int)m); f[k][1] = normalise_u8((int)v[1], (int)m);
        f[k][2] = normalise_u8((int)v[2], (int)m);
      }
    }
    if (p.out_fmt == 0) {
      *reinterpret_cast<float2*>(o32) = make_float2(f[0][0], f[1][0]);
      *reinterpret_cast<float2*>(o32 + plane_sz) = make_float2(f[0][1], f[1][1]);
      *reinterpret_cast<float2*>(o32 + 2 * plane_sz) = make_float2(f[0][2], f[1][2]);
      o32 += S;
    } else {
      uint4 o;
      o.x = pack_bf16x2(f[0][0], f[0][1]); o.y = pack_bf16x2(f[0][2], 0.f);
      o.z = pack_bf16x2(f[1][0], f[1][1]); o.w = pack_bf16x2(f[1][2], 0.f);
      *reinterpret_cast<uint4*>((y & 1) ? o16 + o16_plane : o16) = o;
      if (y & 1) o16 += o16_row;
    }
  }
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

}  // namespace flope

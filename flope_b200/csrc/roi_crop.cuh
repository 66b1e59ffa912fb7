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
// Two generations of kernels live here:
//  * roi_linear2_kernel / roi_lanczos2_kernel (the production path): a CTA owns a strip of output rows of one
//    crop.  The source rows the strip needs are staged into shared memory with one TMA bulk copy per row
//    (16-byte aligned superset of the row segment: every DRAM sector is fetched once, coalesced), then every
//    thread marches down the strip with the horizontally filtered source rows in registers: the window of a
//    column is read as aligned 32-bit words, re-aligned with funnel shifts, gathered into (tap, tap+1) byte
//    pairs with byte permutes and filtered with DP2A (16-bit coefficients x 8-bit pixels), so a filtered
//    value costs ~1.3 instructions instead of a byte load + IMAD per tap.  The register window is a ring
//    indexed by the source row's index modulo the tap count; the *coefficients* are rotated instead of the data.
//  * roi_crop_kernel<TAPS,HAS_MASK> (generic fallback): one thread per output column reading taps straight
//    from global memory.  Used for frame widths that are not a multiple of 4, output sides above 512 and
//    frames too wide for the staging buffer.
// The per-thread programs are plain functions of (shared-memory image, thread index); tests/emu compiles them
// for the host to check the index arithmetic bit for bit against cv2 without a GPU.
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
  // staged kernels only
  const uint8_t* frames_end;  // one past the last byte a bulk copy may read
  const uint8_t* masks_end;
  int cols_cta;               // output columns per CTA
  int n_sub;                  // sub-strips per CTA (linear kernel: threads = n_sub * cols_cta / 2)
  int data_bytes;             // shared-memory staging area
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

// ---- integer SIMD primitives (PRMT / SHF / IDP.2A / IMAD.HI on the device) ----
FLOPE_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = (sel >> (4 * i)) & 0xF;
    uint32_t byte = (uint32_t)(v >> (8 * (s & 7))) & 0xFF;
    if (s & 8) byte = (byte & 0x80) ? 0xFF : 0;
    r |= byte << (8 * i);
  }
  return r;
#endif
}
// low 32 bits of (hi:lo) >> (shift & 31)
FLOPE_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, shift);
#else
  shift &= 31;
  return shift ? (lo >> shift) | (hi << (32 - shift)) : lo;
#endif
}
// c + coef.lo16 * bytes.b0 + coef.hi16 * bytes.b1 (lo) / bytes.b2, bytes.b3 (hi); unsigned coefficients
FLOPE_HD uint32_t dp2a_lo_uu(uint32_t coef, uint32_t bytes, uint32_t c) {
#ifdef __CUDA_ARCH__
  return __dp2a_lo(coef, bytes, c);
#else
  return c + (coef & 0xFFFF) * (bytes & 0xFF) + (coef >> 16) * ((bytes >> 8) & 0xFF);
#endif
}
FLOPE_HD uint32_t dp2a_hi_uu(uint32_t coef, uint32_t bytes, uint32_t c) {
#ifdef __CUDA_ARCH__
  return __dp2a_hi(coef, bytes, c);
#else
  return c + (coef & 0xFFFF) * ((bytes >> 16) & 0xFF) + (coef >> 16) * (bytes >> 24);
#endif
}
// signed 16-bit coefficients x unsigned bytes
FLOPE_HD int dp2a_lo_su(uint32_t coef, uint32_t bytes, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef), "r"(bytes), "r"(c));
  return d;
#else
  return c + (int)(short)(coef & 0xFFFF) * (int)(bytes & 0xFF) + (int)(short)(coef >> 16) * (int)((bytes >> 8) & 0xFF);
#endif
}
FLOPE_HD int dp2a_hi_su(uint32_t coef, uint32_t bytes, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef), "r"(bytes), "r"(c));
  return d;
#else
  return c + (int)(short)(coef & 0xFFFF) * (int)((bytes >> 16) & 0xFF) + (int)(short)(coef >> 16) * (int)(bytes >> 24);
#endif
}
FLOPE_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// =============================================================================================
// Staged kernels
// =============================================================================================
constexpr int kRoi2MaxRows = 192;     // source rows per staged chunk (row table size)
constexpr int kRoi2MaxChunks = 64;
constexpr int kRoi2ZeroWord = 8;      // smem offset of a word that holds 0 (behind the mbarrier)

// Shared memory as the thread programs see it: byte offsets into the CTA's dynamic shared memory.  On the device the
// accessors index the extern array directly, so every access is an LDS/STS with the table offset folded into the
// immediate; on the host (tests/emu) they index a plain buffer.
struct U32x2 { uint32_t x, y; };
struct alignas(16) U32x4 { uint32_t x, y, z, w; };
#ifdef __CUDACC__
extern __shared__ __align__(1024) uint8_t roi_smem[];
// explicit ld.shared / st.shared on 32-bit shared-space addresses: an LDS/STS per access, no generic-window arithmetic
struct SmemDev {
  uint32_t base;
  __device__ __forceinline__ uint8_t* at(int off) const { return roi_smem + off; }
  __device__ __forceinline__ uint32_t ld8(int off) const { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + off)); return v; }
  __device__ __forceinline__ uint32_t ld32(int off) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + off)); return v; }
  __device__ __forceinline__ U32x2 ld64(int off) const { U32x2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + off)); return v; }
  __device__ __forceinline__ U32x4 ld128(int off) const {
    U32x4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + off));
    return v;
  }
  // absolute (shared-space) addresses: thread programs keep their table pointers in registers instead of re-deriving
  // the shared window base at every access
  __device__ __forceinline__ int abs(int off) const { return (int)base + off; }
  __device__ __forceinline__ uint32_t ld32a(int a) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
  __device__ __forceinline__ U32x2 ld64a(int a) const { U32x2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
  __device__ __forceinline__ U32x4 ld128a(int a) const {
    U32x4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  }
  __device__ __forceinline__ void st8(int off, uint32_t v) const { asm volatile("st.shared.u8 [%0], %1;" ::"r"(base + off), "r"(v) : "memory"); }
  __device__ __forceinline__ void st32(int off, uint32_t v) const { asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + off), "r"(v) : "memory"); }
  __device__ __forceinline__ void st64(int off, U32x2 v) const { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(base + off), "r"(v.x), "r"(v.y) : "memory"); }
  __device__ __forceinline__ void st128(int off, U32x4 v) const {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + off), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
};
#endif
struct SmemHost {
  uint8_t* base;
  FLOPE_HD uint8_t* at(int off) const { return base + off; }
  FLOPE_HD uint32_t ld8(int off) const { return base[off]; }
  FLOPE_HD uint32_t ld32(int off) const { uint32_t v; memcpy(&v, base + off, 4); return v; }
  FLOPE_HD U32x2 ld64(int off) const { U32x2 v; memcpy(&v, base + off, 8); return v; }
  FLOPE_HD U32x4 ld128(int off) const { U32x4 v; memcpy(&v, base + off, 16); return v; }
  FLOPE_HD int abs(int off) const { return off; }
  FLOPE_HD uint32_t ld32a(int a) const { return ld32(a); }
  FLOPE_HD U32x2 ld64a(int a) const { return ld64(a); }
  FLOPE_HD U32x4 ld128a(int a) const { return ld128(a); }
  FLOPE_HD void st8(int off, uint32_t v) const { base[off] = (uint8_t)v; }
  FLOPE_HD void st32(int off, uint32_t v) const { memcpy(base + off, &v, 4); }
  FLOPE_HD void st64(int off, U32x2 v) const { memcpy(base + off, &v, 8); }
  FLOPE_HD void st128(int off, U32x4 v) const { memcpy(base + off, &v, 16); }
};

// shared-memory layout (byte offsets); everything a thread program reads lives here
struct Roi2Layout {
  int plan;      // int n_chunks; int chunk_end[kRoi2MaxChunks]
  int lut;       // 256 x u32: normalise_u8(i, 255) as float bits (fmt 0) or bf16 bits (fmt 1)
  int xtab;      // per output column of the CTA
  int ytab;      // per output row of the strip
  int rowtab;    // per staged source row: {image address, mask address} (smem byte offsets)
  int data;      // staged rows
  int total;
};
constexpr int kXtabBytes2 = 8;        // linear : {int sx - x0, u32 c0 | c1 << 16}
constexpr int kYtabBytes2 = 16;       // linear : {int sy, u32 coef of the even source row << 16, same for the odd row, clamped rows r0 | r1 << 16}
constexpr int kXtabBytes8 = 32;       // lanczos: {int sx - 3 - x0, u32 c01, c23, c45, c67, 0, 0, 0}
constexpr int kYtabBytes8 = 48;       // lanczos: {int sy, 0, 0, 0, int coef[8] rotated to the register ring}

FLOPE_HD Roi2Layout roi2_layout(int taps, int cols_cta, int strip_rows, int data_bytes) {
  Roi2Layout L;
  L.plan = 16;
  L.lut = 1024;                                                         // 1 KB aligned: entry address = lut | (4 * value)
  L.xtab = L.lut + 1024;
  L.ytab = L.xtab + cols_cta * (taps == 8 ? kXtabBytes8 : kXtabBytes2);
  L.ytab = (L.ytab + 15) & ~15;
  L.rowtab = L.ytab + strip_rows * (taps == 8 ? kYtabBytes8 : kYtabBytes2);
  L.data = (L.rowtab + kRoi2MaxRows * 8 + 127) & ~127;
  L.total = L.data + data_bytes;
  return L;
}

// per-CTA constants
struct Roi2Cta {
  int crop, frame, xmin, ymin, sw, sh, S;
  int x_begin, x_end;        // output columns of this CTA
  int y_begin, y_end;        // output rows of this CTA
  int xc0, xc1;              // staged source columns (inclusive), clipped to the crop
  int pad_l, pad_r;          // 1 when the CTA's taps reach the replicated border (lanczos)
  int pitch_i, pitch_m;      // bytes per staged image / mask row
  int rows_fit;
  int a0, m0;                // (address of the first staged image / mask byte) & 3: the same for every row (W % 4 == 0)
  double scale_x, scale_y;
  const uint8_t* img0;       // global address of source pixel (row 0, xc0)
  const uint8_t* msk0;
  long long row_bytes_i, row_bytes_m;
  uint8_t* out_crop;         // fmt 0: (crop, channel 0, 0, 0); fmt 1: position (crop, 0, 0) of plane 0
  uint32_t plane_bytes;      // fmt 0: S*S*4; fmt 1: bytes between the two s2d planes (the host checks it fits)
  uint32_t row_step;         // fmt 0: S*4;   fmt 1: Wp*16
};

template <int TAPS>
FLOPE_HD bool roi2_cta_init(const RoiParams& p, int bx, int by, int bz, bool has_mask, Roi2Cta& c) {
  const int32_t* b = p.boxes + (size_t)bz * 5;
  c.crop = bz; c.frame = b[0]; c.xmin = b[1]; c.ymin = b[2];
  c.sw = b[3] - b[1]; c.sh = b[4] - b[2]; c.S = p.S;
  if (c.sw <= 0 || c.sh <= 0) return false;
  c.x_begin = bx * p.cols_cta; c.x_end = imin(p.S, c.x_begin + p.cols_cta);
  c.y_begin = by * p.rows_per_strip; c.y_end = imin(p.S, c.y_begin + p.rows_per_strip);
  c.scale_x = axis_scale(c.sw, p.S);
  c.scale_y = axis_scale(c.sh, p.S);
  float f;
  const int s_first = src_coord(c.x_begin, c.scale_x, f), s_last = src_coord(c.x_end - 1, c.scale_x, f);
  if (TAPS == 8) {
    c.xc0 = iclamp(s_first - 3, 0, c.sw - 1);
    c.xc1 = iclamp(s_last + 4, 0, c.sw - 1);
    c.pad_l = (s_first - 3 < 0) ? 1 : 0;
    c.pad_r = (s_last + 4 > c.sw - 1) ? 1 : 0;
  } else {
    c.xc0 = iclamp(s_first, 0, c.sw - 1);
    c.xc1 = iclamp(s_last + 1, 0, c.sw - 1);
    c.pad_l = c.pad_r = 0;
  }
  const int span = c.xc1 - c.xc0 + 1;
  // 16 bytes in front of the row (replicated border of the 8-tap kernel), up to 15 bytes of alignment, and 16 bytes
  // behind it (aligned-superset tail, window over-read)
  c.pitch_i = ((3 * span + 15 + 15) & ~15) + 32;
  c.pitch_m = ((span + 15 + 15) & ~15) + 32;
  c.rows_fit = imin(kRoi2MaxRows, p.data_bytes / (c.pitch_i + (has_mask ? c.pitch_m : 0)));
  c.row_bytes_i = (long long)p.W * 3;
  c.row_bytes_m = p.W;
  c.img0 = p.frames + (long long)c.frame * p.frame_stride + ((long long)c.ymin * p.W + c.xmin + c.xc0) * 3;
  c.msk0 = has_mask ? p.masks + (long long)c.frame * p.mask_stride + (long long)c.ymin * p.W + c.xmin + c.xc0 : nullptr;
  c.a0 = (int)((uintptr_t)c.img0 & 3);
  c.m0 = (int)((uintptr_t)c.msk0 & 3);
  if (p.out_fmt == 0) {
    c.out_crop = reinterpret_cast<uint8_t*>(p.out) + (long long)bz * 3 * p.S * p.S * 4;
    c.plane_bytes = (uint32_t)p.S * (uint32_t)p.S * 4u;
    c.row_step = (uint32_t)p.S * 4u;
  } else {
    c.out_crop = reinterpret_cast<uint8_t*>(p.out) + (p.g.base + geom_pos(p.g, bz, 0, 0)) * 16;
    c.plane_bytes = (uint32_t)(p.g.plane * 16);
    c.row_step = (uint32_t)p.g.Wp * 16u;
  }
  return true;
}

// ---- tables ----
template <class SV>
FLOPE_HD void roi2_fill_lut(const SV& s, const Roi2Layout& L, int i, int fmt) {
  const float f = normalise_u8(i, 255);
  s.st32(L.lut + 4 * i, fmt == 0 ? f32_bits(f) : bf16_bits_rn(f));
}
template <class SV>
FLOPE_HD void roi2_fill_xtab2(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int i) {
  short ic[2];
  int sx;
  linear_coefs(c.x_begin + i, c.scale_x, c.sw, false, sx, ic);
  U32x2 e;
  e.x = (uint32_t)(sx - c.xc0);
  e.y = (uint32_t)(uint16_t)ic[0] | ((uint32_t)(uint16_t)ic[1] << 16);
  s.st64(L.xtab + i * kXtabBytes2, e);
}
template <class SV>
FLOPE_HD void roi2_fill_ytab2(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int i) {
  short ic[2];
  int sy;
  linear_coefs(c.y_begin + i, c.scale_y, c.sh, true, sy, ic);
  const uint32_t b0 = (uint32_t)(int)ic[0] << 16, b1 = (uint32_t)(int)ic[1] << 16;
  U32x4 e;
  e.x = (uint32_t)sy;
  e.y = (sy & 1) ? b1 : b0;      // coefficient of the even source row of the pair (sy, sy + 1)
  e.z = (sy & 1) ? b0 : b1;      // coefficient of the odd one
  e.w = (uint32_t)iclamp(sy, 0, c.sh - 1) | ((uint32_t)iclamp(sy + 1, 0, c.sh - 1) << 16);
  s.st128(L.ytab + i * kYtabBytes2, e);
}
template <class SV>
FLOPE_HD void roi2_fill_xtab8(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int i) {
  short ic[8];
  int sx;
  lanczos4_coefs(c.x_begin + i, c.scale_x, sx, ic);
  const int o = L.xtab + i * kXtabBytes8;
  s.st32(o, (uint32_t)(sx - 3 - c.xc0));
#pragma unroll
  for (int j = 0; j < 4; ++j)
    s.st32(o + 4 + 4 * j, (uint32_t)(uint16_t)ic[2 * j] | ((uint32_t)(uint16_t)ic[2 * j + 1] << 16));
}
template <class SV>
FLOPE_HD void roi2_fill_ytab8(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int i) {
  short ic[8];
  int sy;
  lanczos4_coefs(c.y_begin + i, c.scale_y, sy, ic);
  const int o = L.ytab + i * kYtabBytes8;
  s.st32(o, (uint32_t)sy);
  // ring slot k holds the source row u with (u & 7) == k; the window is u = sy - 3 + j, j = 0..7
#pragma unroll
  for (int j = 0; j < 8; ++j) s.st32(o + 16 + 4 * ((sy - 3 + j) & 7), (uint32_t)(int)ic[j]);
}
template <int TAPS, class SV>
FLOPE_HD int roi2_ytab_sy(const SV& s, const Roi2Layout& L, int i) {
  return (int)s.ld32(L.ytab + i * (TAPS == 8 ? kYtabBytes8 : kYtabBytes2));
}

// first / last source row (clamped to the crop) that output rows [ya, yb) of the strip read
template <int TAPS, class SV>
FLOPE_HD void roi2_row_span(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int ya, int yb, int& r_lo, int& r_hi) {
  const int s0 = roi2_ytab_sy<TAPS>(s, L, ya - c.y_begin), s1 = roi2_ytab_sy<TAPS>(s, L, yb - 1 - c.y_begin);
  r_lo = iclamp(TAPS == 8 ? s0 - 3 : s0, 0, c.sh - 1);
  r_hi = iclamp(TAPS == 8 ? s1 + 4 : s1 + 1, 0, c.sh - 1);
}

// split the strip into chunks whose source rows fit the staging area (one thread; almost always one chunk)
template <int TAPS, class SV>
FLOPE_HD void roi2_plan_chunks(const Roi2Cta& c, const SV& s, const Roi2Layout& L) {
  int n = 0, ya = c.y_begin;
  while (ya < c.y_end && n < kRoi2MaxChunks) {
    int yb = c.y_end, lo, hi;
    roi2_row_span<TAPS>(c, s, L, ya, yb, lo, hi);
    while (hi - lo + 1 > c.rows_fit && yb > ya + 1) {
      --yb;
      roi2_row_span<TAPS>(c, s, L, ya, yb, lo, hi);
    }
    s.st32(L.plan + 4 * (1 + n++), (uint32_t)yb);
    ya = yb;
  }
  s.st32(L.plan, (uint32_t)n);
}

// one staged source row: where it comes from and where it goes
struct Roi2RowCopy {
  const uint8_t* src;   // 16-byte aligned global address (bulk copy) or the exact first byte (byte copy)
  uint32_t dst;         // smem byte offset
  uint32_t bytes;       // multiple of 16 (bulk copy) or the exact byte count (byte copy)
  bool bulk;
};
// Row `r` of the crop into slot `slot`.  The bulk copy moves the 16-byte aligned superset of the row segment;
// `addr` is the smem offset of the segment's first byte.  Falls back to a byte copy when the aligned superset
// would leave [lo, hi) (first / last bytes of the caller's buffer).
FLOPE_HD Roi2RowCopy roi2_row_copy(const uint8_t* g, int nbytes, const uint8_t* lo, const uint8_t* hi, uint32_t slot_base,
                                   uint32_t& addr) {
  Roi2RowCopy rc;
  const uint32_t mis = (uint32_t)((uintptr_t)g & 15);
  addr = slot_base + 16 + mis;
  const uint8_t* src_al = g - mis;
  const uint32_t bytes = (mis + (uint32_t)nbytes + 15u) & ~15u;
  if (src_al >= lo && src_al + bytes <= hi) {
    rc.src = src_al; rc.dst = slot_base + 16; rc.bytes = bytes; rc.bulk = true;
  } else {
    rc.src = g; rc.dst = addr; rc.bytes = (uint32_t)nbytes; rc.bulk = false;
  }
  return rc;
}
FLOPE_HD void roi2_stage_desc(const RoiParams& p, const Roi2Cta& c, const Roi2Layout& L, bool has_mask, int r_lo, int n_rows,
                              int slot, Roi2RowCopy& ci, Roi2RowCopy& cm, U32x2& rowtab) {
  const int span = c.xc1 - c.xc0 + 1;
  const int r = r_lo + slot;
  const uint32_t base_i = (uint32_t)L.data + (uint32_t)slot * (uint32_t)c.pitch_i;
  ci = roi2_row_copy(c.img0 + (long long)r * c.row_bytes_i, 3 * span, p.frames, p.frames_end, base_i, rowtab.x);
  rowtab.y = 0;
  cm.bytes = 0; cm.bulk = true; cm.src = nullptr; cm.dst = 0;
  if (has_mask) {
    const uint32_t base_m = (uint32_t)L.data + (uint32_t)n_rows * (uint32_t)c.pitch_i + (uint32_t)slot * (uint32_t)c.pitch_m;
    cm = roi2_row_copy(c.msk0 + (long long)r * c.row_bytes_m, span, p.masks, p.masks_end, base_m, rowtab.y);
  }
}
// replicate border of the crop for the 8-tap kernel: 4 pixels in front of / behind the staged row
template <class SV>
FLOPE_HD void roi2_patch_row(const Roi2Cta& c, const SV& s, const Roi2Layout& L, bool has_mask, int slot, int side) {
  U32x2 rt = s.ld64(L.rowtab + slot * 8);
  rt.x -= (uint32_t)s.abs(0); rt.y -= (uint32_t)s.abs(0);      // the row table holds absolute addresses
  const int span = c.xc1 - c.xc0 + 1;
  if (side == 0 && c.pad_l) {
    const int px = (int)rt.x;
    const uint32_t c0 = s.ld8(px), c1 = s.ld8(px + 1), c2 = s.ld8(px + 2);
    for (int j = 1; j <= 4; ++j) { s.st8(px - 3 * j, c0); s.st8(px - 3 * j + 1, c1); s.st8(px - 3 * j + 2, c2); }
    if (has_mask) { const int m = (int)rt.y; const uint32_t v = s.ld8(m); for (int j = 1; j <= 4; ++j) s.st8(m - j, v); }
  }
  if (side == 1 && c.pad_r) {
    const int px = (int)rt.x + 3 * (span - 1);
    const uint32_t c0 = s.ld8(px), c1 = s.ld8(px + 1), c2 = s.ld8(px + 2);
    for (int j = 1; j <= 4; ++j) { s.st8(px + 3 * j, c0); s.st8(px + 3 * j + 1, c1); s.st8(px + 3 * j + 2, c2); }
    if (has_mask) { const int m = (int)rt.y + span - 1; const uint32_t v = s.ld8(m); for (int j = 1; j <= 4; ++j) s.st8(m + j, v); }
  }
}

// masked normalise of the three channels of one pixel that is not fully masked out (m4 >= 4): m4 / v4 are 4 x the cv2
// result (bits 2..9), low bits garbage.  LUT: unmasked values come from the 256-entry table in shared memory, otherwise
// from arithmetic (fmt 1: bf16_rn(fl(4v * fl(1/1020))) == bf16_rn(fp32(v / 255)) for all 256 values; fmt 0: the exact
// fp32 quotient).  Fully masked-out pixels (m4 < 4) are zero whatever the image holds: callers skip their image channels.
template <int FMT, bool LUT, class SV>
FLOPE_HD void roi2_pixel(const SV& s, int lut_abs /*1 KB aligned*/, const uint32_t (&v4)[3], uint32_t m4, uint32_t (&o)[3]) {
  if (m4 >= 1020u) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (LUT) {
        o[c] = s.ld32a((int)((v4[c] & 0x3FCu) | (uint32_t)lut_abs));
      } else if (FMT == 0) {
        o[c] = f32_bits(normalise_u8((int)((v4[c] >> 2) & 0xFFu), 255));
      } else {
        o[c] = f32_bits((float)(v4[c] & 0x3FCu) * (1.0f / 1020.0f));     // packed to bf16 by the caller
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float f = normalise_u8((int)(v4[c] >> 2), (int)(m4 >> 2));
      o[c] = (FMT == 0 || !LUT) ? f32_bits(f) : bf16_bits_rn(f);
    }
  }
}
// fmt 1: two 32-bit words [c0 c1 | c2 0] of bf16 from roi2_pixel's output
template <bool LUT>
FLOPE_HD void roi2_pack_bf16(const uint32_t (&o)[3], uint32_t& lo, uint32_t& hi) {
  if (LUT) {
    lo = prmt(o[0], o[1], 0x5410);
    hi = o[2];
  } else {
#ifdef __CUDA_ARCH__
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(__uint_as_float(o[1])), "f"(__uint_as_float(o[0])));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(0.f), "f"(__uint_as_float(o[2])));
#else
    float f[3];
    memcpy(f, o, 12);
    lo = bf16_bits_rn(f[0]) | (bf16_bits_rn(f[1]) << 16);
    hi = bf16_bits_rn(f[2]);
#endif
  }
}

// =============================================================================================
// bilinear thread program: two adjacent output columns (one 16-byte s2d stem pixel), rows [ya, yb)
// =============================================================================================
struct Roi2Col2 {           // loop-invariant per column
  int iofs;                 // byte offset of the aligned word holding the first tap byte, relative to the row's first byte
  uint32_t ish;             // 8 * (misalignment of the first tap byte)
  uint32_t cx;              // c0 | c1 << 16
  int mofs;                 // same for the mask row
  uint32_t msel;            // byte-permute selector of the two mask taps
};
template <class SV>
FLOPE_HD Roi2Col2 roi2_col2(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int col) {
  const U32x2 e = s.ld64(L.xtab + col * kXtabBytes2);
  const int sxr = (int)e.x;
  Roi2Col2 k;
  k.cx = e.y;
  const int o = (c.a0 + 3 * sxr) & 3;
  k.iofs = 3 * sxr - o;
  k.ish = 8u * (uint32_t)o;
  const int om = (c.m0 + sxr) & 3;
  k.mofs = sxr - om;
  k.msel = (uint32_t)om | ((uint32_t)(om + 1) << 4);
  return k;
}
// horizontally filtered source row, >> 4 (what cv2's vertical pass consumes), for one column
template <bool HAS_MASK, class SV>
FLOPE_HD void roi2_hfilt2(const SV& s, uint32_t img_addr, uint32_t msk_addr, const Roi2Col2& k,
                          uint32_t (&dst)[HAS_MASK ? 4 : 3]) {
  const int wa = (int)img_addr + k.iofs;
  const uint32_t w0 = s.ld32a(wa), w1 = s.ld32a(wa + 4), w2 = s.ld32a(wa + 8);
  const uint32_t u0 = funnel_r(w0, w1, k.ish), u1 = funnel_r(w1, w2, k.ish);   // bytes: a0 a1 a2 b0 | b1 b2 . .
  const uint32_t r1 = prmt(u0, u1, 0x4130);                                    // a0 b0 a1 b1
  const uint32_t r2 = prmt(u0, u1, 0x0052);                                    // a2 b2 . .
  dst[0] = dp2a_lo_uu(k.cx, r1, 0) >> 4;
  dst[1] = dp2a_hi_uu(k.cx, r1, 0) >> 4;
  dst[2] = dp2a_lo_uu(k.cx, r2, 0) >> 4;
  if (HAS_MASK) {
    const int ma = (int)msk_addr + k.mofs;
    const uint32_t r3 = prmt(s.ld32a(ma), s.ld32a(ma + 4), k.msel);
    dst[HAS_MASK ? 3 : 0] = dp2a_lo_uu(k.cx, r3, 0) >> 4;
  }
}

template <bool HAS_MASK, int FMT, bool LUT, class SV>
FLOPE_HD void roi2_linear_thread(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int q /*column pair of the CTA*/,
                                 int ya, int yb, int r_lo) {
  constexpr int NCH = HAS_MASK ? 4 : 3;
  const Roi2Col2 k0 = roi2_col2(c, s, L, 2 * q), k1 = roi2_col2(c, s, L, 2 * q + 1);
  uint32_t E[2][NCH], O[2][NCH];           // filtered even / odd source rows of the current pair, per column
#pragma unroll
  for (int j = 0; j < NCH; ++j) E[0][j] = E[1][j] = O[0][j] = O[1][j] = 0;
  // table pointers as absolute shared addresses, made opaque by adding a zero the compiler cannot see through: they
  // stay in registers instead of being re-derived from the shared window base in every basic block
  const int zero = (int)s.ld32(kRoi2ZeroWord);
  const int rowtab0 = s.abs(L.rowtab - r_lo * 8) + zero;
  const int lut_abs = s.abs(L.lut) + zero;
  auto load_row = [&](bool odd, int r) {   // clamped source row r into the register set of the unclamped row's parity
    const U32x2 rt = s.ld64a(rowtab0 + r * 8);
    if (odd) {
      roi2_hfilt2<HAS_MASK>(s, rt.x, rt.y, k0, O[0]);
      roi2_hfilt2<HAS_MASK>(s, rt.x, rt.y, k1, O[1]);
    } else {
      roi2_hfilt2<HAS_MASK>(s, rt.x, rt.y, k0, E[0]);
      roi2_hfilt2<HAS_MASK>(s, rt.x, rt.y, k1, E[1]);
    }
  };
  const int x = c.x_begin + 2 * q;
  // fmt 0: byte offset of (channel 0, row y, column x); fmt 1: of the s2d pixel (y >> 1, x >> 1) in plane 0
  uint32_t ooff = FMT == 0 ? (uint32_t)(ya * c.S + x) * 4u : (uint32_t)(ya >> 1) * c.row_step + (uint32_t)(x >> 1) * 16u;
  int yofs = s.abs(L.ytab + (ya - c.y_begin) * kYtabBytes2) + zero;
  int u = -0x40000000;
  for (int y = ya; y < yb; ++y, yofs += kYtabBytes2) {
    const U32x4 yt = s.ld128a(yofs);
    const int sy = (int)yt.x;
    if (sy != u) {
      const int r0 = (int)(yt.w & 0xFFFFu), r1 = (int)(yt.w >> 16);
      if (sy == u + 1) {
        load_row(!(sy & 1), r1);
      } else {
        load_row(sy & 1, r0);
        load_row(!(sy & 1), r1);
      }
      u = sy;
    }
    uint32_t o[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t m4 = 1020u;
      if (HAS_MASK) m4 = mulhi_u32(yt.y, E[k][NCH - 1]) + mulhi_u32(yt.z, O[k][NCH - 1]) + 2u;
      if (m4 < 4u) {                       // masked out: exactly zero, the image channels are not needed
        o[k][0] = o[k][1] = o[k][2] = 0u;
      } else {
        uint32_t v4[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) v4[j] = mulhi_u32(yt.y, E[k][j]) + mulhi_u32(yt.z, O[k][j]) + 2u;
        roi2_pixel<FMT, LUT>(s, lut_abs, v4, m4, o[k]);
      }
    }
    if (FMT == 0) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        U32x2 v; v.x = o[0][ch]; v.y = o[1][ch];
        *reinterpret_cast<U32x2*>(c.out_crop + ooff + (uint32_t)ch * c.plane_bytes) = v;
      }
      ooff += c.row_step;
    } else {
      U32x4 v;
      roi2_pack_bf16<LUT>(o[0], v.x, v.y);
      roi2_pack_bf16<LUT>(o[1], v.z, v.w);
      *reinterpret_cast<U32x4*>(c.out_crop + ooff + ((y & 1) ? c.plane_bytes : 0u)) = v;
      if (y & 1) ooff += c.row_step;
    }
  }
}

// =============================================================================================
// Lanczos4 thread program: one output column, rows [ya, yb)
// =============================================================================================
struct Roi2Col8 {
  int iofs; uint32_t ish;
  int mofs; uint32_t msh;
  uint32_t c01, c23, c45, c67;
};
template <class SV>
FLOPE_HD Roi2Col8 roi2_col8(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int col) {
  const int o = L.xtab + col * kXtabBytes8;
  const int sxr = (int)s.ld32(o);     // first tap, relative to the first staged pixel (-4 .. : replicated border)
  Roi2Col8 k;
  k.c01 = s.ld32(o + 4); k.c23 = s.ld32(o + 8);
  k.c45 = s.ld32(o + 12); k.c67 = s.ld32(o + 16);
  const int oi = (c.a0 + 3 * sxr) & 3;  // two's complement & 3 == mod 4 for negative sxr as well
  k.iofs = 3 * sxr - oi;
  k.ish = 8u * (uint32_t)oi;
  const int om = (c.m0 + sxr) & 3;
  k.mofs = sxr - om;
  k.msh = 8u * (uint32_t)om;
  return k;
}
template <bool HAS_MASK, class SV>
FLOPE_HD void roi2_hfilt8(const SV& s, uint32_t img_addr, uint32_t msk_addr, const Roi2Col8& k, int (&dst)[HAS_MASK ? 4 : 3]) {
  const int wa = (int)img_addr + k.iofs;
  uint32_t u[6];
  {
    uint32_t wv[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) wv[i] = s.ld32a(wa + 4 * i);
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = funnel_r(wv[i], wv[i + 1], k.ish);   // bytes 0..23 = 8 pixels x 3 channels
  }
  int a0, a1, a2;
  uint32_t r;
  r = prmt(u[0], u[1], 0x4130); a0 = dp2a_lo_su(k.c01, r, 0); a1 = dp2a_hi_su(k.c01, r, 0);     // taps 0,1: bytes (0,3) (1,4)
  r = prmt(u[0], u[1], 0x0052); a2 = dp2a_lo_su(k.c01, r, 0);                                    //           bytes (2,5)
  r = prmt(u[1], u[2], 0x6352); a0 = dp2a_lo_su(k.c23, r, a0); a1 = dp2a_hi_su(k.c23, r, a1);   // taps 2,3: bytes (6,9) (7,10)
  r = prmt(u[1], u[2], 0x0074); a2 = dp2a_lo_su(k.c23, r, a2);                                   //           bytes (8,11)
  r = prmt(u[3], u[4], 0x4130); a0 = dp2a_lo_su(k.c45, r, a0); a1 = dp2a_hi_su(k.c45, r, a1);   // taps 4,5: bytes (12,15) (13,16)
  r = prmt(u[3], u[4], 0x0052); a2 = dp2a_lo_su(k.c45, r, a2);
  r = prmt(u[4], u[5], 0x6352); a0 = dp2a_lo_su(k.c67, r, a0); a1 = dp2a_hi_su(k.c67, r, a1);   // taps 6,7: bytes (18,21) (19,22)
  r = prmt(u[4], u[5], 0x0074); a2 = dp2a_lo_su(k.c67, r, a2);
  dst[0] = a0; dst[1] = a1; dst[2] = a2;
  if (HAS_MASK) {
    const int ma = (int)msk_addr + k.mofs;
    const uint32_t m0 = s.ld32a(ma), m1 = s.ld32a(ma + 4), m2 = s.ld32a(ma + 8);
    const uint32_t t0 = funnel_r(m0, m1, k.msh), t1 = funnel_r(m1, m2, k.msh);
    int a3 = dp2a_lo_su(k.c01, t0, 0);
    a3 = dp2a_hi_su(k.c23, t0, a3);
    a3 = dp2a_lo_su(k.c45, t1, a3);
    a3 = dp2a_hi_su(k.c67, t1, a3);
    dst[HAS_MASK ? 3 : 0] = a3;
  }
}

template <bool HAS_MASK, int FMT, bool LUT, class SV>
FLOPE_HD void roi2_lanczos_thread(const Roi2Cta& c, const SV& s, const Roi2Layout& L, int col, int ya, int yb, int r_lo) {
  constexpr int NCH = HAS_MASK ? 4 : 3;
  const Roi2Col8 k = roi2_col8(c, s, L, col);
  int R[8][NCH];                           // ring: slot (u & 7) holds filtered source row u
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NCH; ++j) R[i][j] = 0;
  const int zero = (int)s.ld32(kRoi2ZeroWord);
  const int rowtab0 = s.abs(L.rowtab - r_lo * 8) + zero;
  const int lut_abs = s.abs(L.lut) + zero;
  auto load_row = [&](int ur) {
    const int r = iclamp(ur, 0, c.sh - 1);
    const U32x2 rt = s.ld64a(rowtab0 + r * 8);
    switch (ur & 7) {
      case 0: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[0]); break;
      case 1: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[1]); break;
      case 2: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[2]); break;
      case 3: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[3]); break;
      case 4: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[4]); break;
      case 5: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[5]); break;
      case 6: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[6]); break;
      default: roi2_hfilt8<HAS_MASK>(s, rt.x, rt.y, k, R[7]); break;
    }
  };
  const int x = c.x_begin + col;
  uint32_t ooff = FMT == 0 ? (uint32_t)(ya * c.S + x) * 4u
                           : (uint32_t)(ya >> 1) * c.row_step + (uint32_t)(x >> 1) * 16u + (uint32_t)(x & 1) * 8u;
  int yofs = s.abs(L.ytab + (ya - c.y_begin) * kYtabBytes8) + zero;
  int top = -0x40000000;                   // highest source row in the ring
  for (int y = ya; y < yb; ++y, yofs += kYtabBytes8) {
    const int new_top = (int)s.ld32a(yofs) + 4;
    if (new_top > top) {
      for (int ur = imax(top + 1, new_top - 7); ur <= new_top; ++ur) load_row(ur);
      top = new_top;
    }
    const U32x4 ca = s.ld128a(yofs + 16), cb = s.ld128a(yofs + 32);
    const int cf[8] = {(int)ca.x, (int)ca.y, (int)ca.z, (int)ca.w, (int)cb.x, (int)cb.y, (int)cb.z, (int)cb.w};
    auto vpass = [&](int j) {
      int acc = 1 << 21;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += R[i][j] * cf[i];
      return (uint32_t)iclamp(acc >> 22, 0, 255) << 2;
    };
    uint32_t o[3];
    const uint32_t m4 = HAS_MASK ? vpass(NCH - 1) : 1020u;
    if (m4 < 4u) {                         // masked out: exactly zero, the image channels are not needed
      o[0] = o[1] = o[2] = 0u;
    } else {
      uint32_t v4[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) v4[j] = vpass(j);
      roi2_pixel<FMT, LUT>(s, lut_abs, v4, m4, o);
    }
    if (FMT == 0) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) *reinterpret_cast<uint32_t*>(c.out_crop + ooff + (uint32_t)ch * c.plane_bytes) = o[ch];
      ooff += c.row_step;
    } else {
      U32x2 v;
      roi2_pack_bf16<LUT>(o, v.x, v.y);
      *reinterpret_cast<U32x2*>(c.out_crop + ooff + ((y & 1) ? c.plane_bytes : 0u)) = v;
      if (y & 1) ooff += c.row_step;
    }
  }
}

// rows of a chunk dealt to sub-strip `sub` of `n_sub`
FLOPE_HD void roi2_sub_rows(int ya, int yb, int sub, int n_sub, int& a, int& b) {
  const int part = (yb - ya + n_sub - 1) / n_sub;
  a = imin(yb, ya + sub * part);
  b = imin(yb, a + part);
}

#ifdef __CUDACC__
// =============================================================================================
// kernels
// =============================================================================================
__device__ __forceinline__ void roi2_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int TAPS, bool HAS_MASK, int FMT, bool LUT>
__global__ void __launch_bounds__(256, TAPS == 8 ? 2 : 4) roi2_kernel(const __grid_constant__ RoiParams p) {
#ifdef __CUDA_ARCH__
  const SmemDev s{smem_u32(roi_smem)};
  const int t = threadIdx.x, nt = blockDim.x;
  Roi2Cta c;
  if (!roi2_cta_init<TAPS>(p, blockIdx.x, blockIdx.y, blockIdx.z, HAS_MASK, c)) return;   // empty box: the host rejects these
  const int ncols = c.x_end - c.x_begin, nrows = c.y_end - c.y_begin;
  const Roi2Layout L = roi2_layout(TAPS, p.cols_cta, p.rows_per_strip, p.data_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s.at(0));
  if (t == 0) {
    if (s.base & 1023u) __trap();          // the table look-ups OR their index into 1 KB aligned table addresses
    mbar_init(bar, (uint32_t)nt);
    mbar_fence_init();
    s.st32(kRoi2ZeroWord, 0u);
  }
  for (int i = t; i < ncols; i += nt) { if (TAPS == 8) roi2_fill_xtab8(c, s, L, i); else roi2_fill_xtab2(c, s, L, i); }
  for (int i = t; i < nrows; i += nt) { if (TAPS == 8) roi2_fill_ytab8(c, s, L, i); else roi2_fill_ytab2(c, s, L, i); }
  for (int i = t; i < 256; i += nt) roi2_fill_lut(s, L, i, FMT);
  __syncthreads();
  if (t == 0) roi2_plan_chunks<TAPS>(c, s, L);
  __syncthreads();
  const int n_chunks = (int)s.ld32(L.plan);
  uint32_t parity = 0;
  int ya = c.y_begin;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int yb = (int)s.ld32(L.plan + 4 * (1 + ch));
    int r_lo, r_hi;
    roi2_row_span<TAPS>(c, s, L, ya, yb, r_lo, r_hi);
    const int n_rows = r_hi - r_lo + 1;
    if (ch > 0) {
      __syncthreads();                     // every thread is done with the previous chunk's rows
      roi2_fence_proxy_async();
    }
    // ---- stage: thread `slot` copies source row r_lo + slot (image and mask) ----
    uint32_t tx = 0;
    for (int slot = t; slot < n_rows; slot += nt) {
      Roi2RowCopy ci, cm;
      U32x2 rt;
      roi2_stage_desc(p, c, L, HAS_MASK, r_lo, n_rows, slot, ci, cm, rt);
      rt.x = (uint32_t)s.abs((int)rt.x); rt.y = (uint32_t)s.abs((int)rt.y);
      s.st64(L.rowtab + slot * 8, rt);
      if (ci.bulk) tx += ci.bytes; else for (uint32_t i = 0; i < ci.bytes; ++i) *s.at((int)(ci.dst + i)) = __ldg(ci.src + i);
      if (HAS_MASK) { if (cm.bulk) tx += cm.bytes; else for (uint32_t i = 0; i < cm.bytes; ++i) *s.at((int)(cm.dst + i)) = __ldg(cm.src + i); }
    }
    if (tx) mbar_expect_tx(bar, tx); else mbar_arrive(bar);
    for (int slot = t; slot < n_rows; slot += nt) {
      Roi2RowCopy ci, cm;
      U32x2 rt;
      roi2_stage_desc(p, c, L, HAS_MASK, r_lo, n_rows, slot, ci, cm, rt);
      if (ci.bulk) bulk_g2s(s.at((int)ci.dst), ci.src, ci.bytes, bar);
      if (HAS_MASK && cm.bulk) bulk_g2s(s.at((int)cm.dst), cm.src, cm.bytes, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    if (TAPS == 8 && (c.pad_l || c.pad_r)) {
      for (int i = t; i < 2 * n_rows; i += nt) roi2_patch_row(c, s, L, HAS_MASK, i >> 1, i & 1);
      __syncthreads();
    }
    // ---- compute ----
    if (TAPS == 8) {
      if (t < ncols) roi2_lanczos_thread<HAS_MASK, FMT, LUT>(c, s, L, t, ya, yb, r_lo);
    } else {
      const int pairs = ncols >> 1;
      const int sub = t / pairs, q = t - sub * pairs;
      if (sub < p.n_sub) {
        int a, b;
        roi2_sub_rows(ya, yb, sub, p.n_sub, a, b);
        if (a < b) roi2_linear_thread<HAS_MASK, FMT, LUT>(c, s, L, q, a, b, r_lo);
      }
    }
    ya = yb;
  }
#endif
}

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
  if (sw <= 0 || sh <= 0) return;                 // the host rejects empty boxes; never dereference one

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

// flope_b200: streaming ROI crop / resize / mask / normalise kernels (sm_100a) - the production path.
//
// Same arithmetic as roi_crop.cuh (cv2's uint8 fixed-point resize, bit for bit; see the header there and
// oracle/resize.py), organised as a persistent, warp-specialised pipeline:
//
//   work item   = (crop, strip of `rows_per_item` output rows, block of `cols_per_item` output columns); the first item
//                 of a CTA is its index, the rest are claimed from a global counter, so the grid (CTAs per SM x SMs)
//                 balances boxes of different sizes by itself.
//   producer    = the last warp of the CTA.  Per item it writes a header, the vertical table (per output row: the
//                 source row that completes it, the coefficients, the output row offset) and the horizontal table
//                 (per output column: where its taps sit in a staged row, the coefficients) into one of two item
//                 slots, then streams the item's source rows - image and mask - through a ring of shared-memory
//                 stages with one TMA bulk copy per row and matrix (16-byte aligned superset of the row segment, so
//                 every DRAM sector is fetched once).  Rows above / below the crop are the clamped row: the replicated
//                 border of cv2 costs the consumers nothing vertically.
//   consumers   = one thread per output column.  A thread walks DOWN the source rows as they arrive: horizontal
//                 filter of the row into registers (aligned 32-bit words, funnel shift, byte permute, DP2A), then
//                 every output row this source row completes is emitted (vertical filter, mask, normalise, store).
//                 Source rows are filtered once per strip and column, raw rows are dead as soon as they are filtered,
//                 so a stage is handed back to the producer after a few rows: the ring is small (3 x 14 KB), four
//                 CTAs fit an SM and loads run ahead of compute.  The warps of a CTA move through the ring together:
//                 a strip costs what its busiest warp costs, so the per-pixel work is kept branch-free and identical
//                 for every warp (see r3_finish) instead of offering shortcuts to warps with masked-out pixels.
//
// Bilinear: the register window is two rows (the loop is unrolled over a row pair, no data moves).  The vertical
// pass - two separately truncated products in cv2 - runs on the FMA pipe with round-toward-minus-infinity FMAs
// (r3_emit2), not as 32x32 high multiplies; the masked normalise is two more FMAs per channel on the same bit patterns.
// Lanczos4: the register window is a ring of eight rows indexed by (source row & 7); the table rotates the
// coefficients instead of the data.  The replicated border left / right of the crop is written into the padding of
// the staged rows by the warps whose columns reach it (a warp patches for itself, so no CTA-wide barrier).  The
// coefficient tables (cv2's double-precision formula) come from a pre-kernel, roi3_axis_tables_kernel.
//
// Requirements checked by the host (engine.cu: run_roi): W % 16 == 0 (the 16-byte phase of a row segment is the
// same for every row of a crop), cols_per_item a multiple of 32, rows fit a stage.  Anything else takes roi_crop_kernel.
#pragma once
#include "roi_crop.cuh"

namespace flope {

struct Roi3Params {
  const uint8_t* frames;      // (n_frames, H, W, 3) u8
  long long frame_stride;
  const uint8_t* masks;       // (n_frames, H, W) u8 or nullptr
  long long mask_stride;
  int W;
  const int32_t* boxes;       // (n, 5): frame, xmin, ymin, xmax, ymax
  int n;
  int S;
  int out_fmt;
  void* out;
  Geom g;                     // fmt 1 geometry
  int rows_per_item;          // output rows per work item (<= kR3MaxItemRows)
  int cols_per_item;          // output columns per work item = consumer threads (multiple of 32, divides S)
  int col_blocks;             // S / cols_per_item
  int items_per_crop;
  int n_items;
  int stage_bytes;            // multiple of 128
  int n_stages;               // <= kR3MaxStages
  int xtab_slot;              // cols_per_item * bytes per horizontal table entry
  int ring_off;               // r3_xtab_off(TAPS) + 2 * xtab_slot, multiple of 128
  const uint32_t* axis_tab;   // Lanczos4 only: [crop][axis: 0 = x, 1 = y][S] entries of 8 words {floor(src coord), c01, c23, c45, c67, 0, 0, 0},
                              // built by roi3_axis_tables_kernel right before the launch (nullptr: the producer computes them)
  unsigned int* sched;        // {next item, CTAs done}: both zero at launch, reset by the last CTA (nullptr = static round-robin)
};

constexpr int kR3MaxItemRows = 128;
constexpr int kR3MaxStages = 8;
// shared-memory map (bytes)
constexpr int kR3Full = 0;                     // mbarrier per stage: rows have landed
constexpr int kR3Empty = 64;                   // mbarrier per stage: every consumer warp is done with the rows
constexpr int kR3ItemFull = 128;               // mbarrier per item slot: header + tables written
constexpr int kR3ItemEmpty = 144;              // mbarrier per item slot: every consumer warp is done with them
constexpr int kR3Konst = 176;                  // 4 words of fp32 constants the bf16 consumers keep in registers (read once from here:
                                               // an immediate would be re-materialised inside the loop by ptxas)
constexpr int kR3Hdr = 192;                    // 2 x 64 B
constexpr int kR3Ytab = 320;                   // 2 x (kR3MaxItemRows + 1) entries
constexpr int kR3YtabEntry2 = 16;              // linear : {u_top, float b0 * 2^-20, float b1 * 2^-20, output row offset}
constexpr int kR3YtabEntry8 = 48;              // lanczos: {u_top, output row offset, 0, 0, int coef[8] by ring slot}
// per kernel family (TAPS = 2 or 8): bytes of one vertical-table slot, offset of the 256 x u32 normalise table (1 KB
// aligned) and of the horizontal table (2 x cols_per_item entries; the ring follows at Roi3Params::ring_off)
FLOPE_HD constexpr int r3_ytab_slot(int taps) { return (kR3MaxItemRows + 1) * (taps == 8 ? kR3YtabEntry8 : kR3YtabEntry2); }
FLOPE_HD constexpr int r3_lut_off(int taps) { return (kR3Ytab + 2 * r3_ytab_slot(taps) + 1023) & ~1023; }
FLOPE_HD constexpr int r3_xtab_off(int taps) { return r3_lut_off(taps) + 1024; }
constexpr int kR3XtabEntry2 = 16;              // linear : {iofs | ish << 16, a0 | a1 << 16, mofs | msel << 16, 0}
constexpr int kR3XtabEntry8 = 32;              // lanczos: {iofs | ish << 16, mofs | msh << 16, border flags, 0, c01, c23, c45, c67}
constexpr int kR3RingTail = 64;                // over-read of the last row's window
// staged row slot of the Lanczos4 kernel: [16 pad][image superset][32 pad][16 pad][mask superset][16 pad]
constexpr int kR3PadL = 16, kR3PadR = 32, kR3PadM = 16;

// header words
enum { R3H_VALID = 0, R3H_CROP, R3H_UFIRST, R3H_NROWS, R3H_K, R3H_PITCH, R3H_XBEGIN, R3H_RGB0, R3H_MSK0, R3H_SPAN, R3H_PADS };

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t r3_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t r3_lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ U32x4 r3_lds128(uint32_t a) {
  U32x4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void r3_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void r3_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void r3_sts128(uint32_t a, U32x4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void r3_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void r3_bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void r3_bar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must trap (a visible CUDA error), never hang the box.  try_wait suspends the
// thread until the phase completes or a time limit passes, so the loop is not a hot spin.
__device__ __forceinline__ void r3_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (ok) return;
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void r3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ int r3_dp2a_lo_su(uint32_t coef, uint32_t bytes, int c) {      // signed 16-bit x unsigned 8-bit
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef), "r"(bytes), "r"(c));
  return d;
}
__device__ __forceinline__ int r3_dp2a_hi_su(uint32_t coef, uint32_t bytes, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef), "r"(bytes), "r"(c));
  return d;
}

// ---------------------------------------------------------------------------------------------
// producer warp
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int r3_next_item(const Roi3Params& p, int item, int lane) {
  if (!p.sched) return item + (int)gridDim.x;
  unsigned int v = 0;
  if (lane == 0) v = atomicAdd(p.sched, 1u);
  return (int)gridDim.x + (int)__shfl_sync(0xFFFFFFFFu, v, 0);
}

// Lanczos4 coefficients of destination index d along one axis: from the per-launch table when there is one (the double
// precision sin / cos / divisions of cv2's formula are ~1000 instructions per entry - far too much for ONE producer warp
// that has to keep eight consumer warps fed), else computed in place.
__device__ __forceinline__ void r3_lanczos_entry(const Roi3Params& p, int crop, int axis, int d, double scale, int& s_out, short (&ic)[8]) {
  if (p.axis_tab) {
    const uint4* e = reinterpret_cast<const uint4*>(p.axis_tab + (((size_t)crop * 2 + axis) * p.S + d) * 8);
    const uint4 a = __ldg(e);
    const uint32_t c67 = __ldg(reinterpret_cast<const uint32_t*>(e + 1));
    s_out = (int)a.x;
    ic[0] = (short)(a.y & 0xFFFFu); ic[1] = (short)(a.y >> 16); ic[2] = (short)(a.z & 0xFFFFu); ic[3] = (short)(a.z >> 16);
    ic[4] = (short)(a.w & 0xFFFFu); ic[5] = (short)(a.w >> 16); ic[6] = (short)(c67 & 0xFFFFu); ic[7] = (short)(c67 >> 16);
  } else {
    lanczos4_coefs(d, scale, s_out, ic);
  }
}
__global__ void __launch_bounds__(256) roi3_axis_tables_kernel(const int32_t* boxes, int n, int S, uint32_t* tab) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2LL * n * S) return;
  const int d = (int)(i % S), axis = (int)((i / S) & 1), crop = (int)(i / (2LL * S));
  const int32_t* bx = boxes + (size_t)crop * 5;
  const int src = axis ? bx[4] - bx[2] : bx[3] - bx[1];
  uint4 a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, 0u, 0u);
  if (src > 0) {
    short ic[8]; int s;
    lanczos4_coefs(d, axis_scale(src, S), s, ic);
    a.x = (uint32_t)s;
    a.y = (uint32_t)(uint16_t)ic[0] | ((uint32_t)(uint16_t)ic[1] << 16);
    a.z = (uint32_t)(uint16_t)ic[2] | ((uint32_t)(uint16_t)ic[3] << 16);
    a.w = (uint32_t)(uint16_t)ic[4] | ((uint32_t)(uint16_t)ic[5] << 16);
    b.x = (uint32_t)(uint16_t)ic[6] | ((uint32_t)(uint16_t)ic[7] << 16);
  }
  uint4* e = reinterpret_cast<uint4*>(tab + (size_t)i * 8);
  e[0] = a; e[1] = b;
}

template <int TAPS, bool HAS_MASK>
__device__ __forceinline__ void r3_producer(const Roi3Params& p, uint32_t sb, int lane) {
  constexpr int LO = TAPS == 8 ? 3 : 0;          // first / last tap relative to floor(source coordinate)
  constexpr int HI = TAPS == 8 ? 4 : 1;
  constexpr int YE = TAPS == 8 ? kR3YtabEntry8 : kR3YtabEntry2;
  uint32_t stage = 0, sphase = 0;
  int it = 0;
  for (int item = blockIdx.x;; item = r3_next_item(p, item, lane)) {
    const int b = it & 1;
    r3_bar_wait(sb + kR3ItemEmpty + 8 * b, (((uint32_t)it >> 1) & 1u) ^ 1u);
    const uint32_t hdr = sb + kR3Hdr + 64 * b;
    if (item >= p.n_items) {
      if (lane == 0) {
        r3_sts32(hdr + 4 * R3H_VALID, 0u); r3_bar_arrive(sb + kR3ItemFull + 8 * b);
        if (p.sched && atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) { p.sched[0] = 0u; p.sched[1] = 0u; __threadfence(); }
      }
      return;
    }
    const int crop = item / p.items_per_crop, rem = item - crop * p.items_per_crop;
    const int strip = rem / p.col_blocks, cb = rem - strip * p.col_blocks;
    const int32_t* bx = p.boxes + (size_t)crop * 5;
    const int frame = __ldg(bx), xmin = __ldg(bx + 1), ymin = __ldg(bx + 2);
    const int sw = __ldg(bx + 3) - xmin, sh = __ldg(bx + 4) - ymin;
    const int S = p.S;
    const int y_begin = strip * p.rows_per_item, y_end = imin(S, y_begin + p.rows_per_item);
    const int x_begin = cb * p.cols_per_item, x_end = x_begin + p.cols_per_item;
    const int n_out = y_end - y_begin;
    if (sw <= 0 || sh <= 0) {
      // Empty box (the reference's cv2.resize raises; the Python layer rejects it): the crop is defined as zeros, so that
      // its slot never keeps a previous batch's pixels.  A "zero item": the output row offsets only, no source rows.
      const uint32_t zt = sb + kR3Ytab + r3_ytab_slot(TAPS) * b;
      for (int i = lane; i < n_out; i += 32) {
        const int y = y_begin + i;
        const uint32_t ooff = p.out_fmt == 0 ? (uint32_t)y * (uint32_t)S * 4u
                                             : (uint32_t)(y >> 1) * (uint32_t)p.g.Wp * 16u + ((y & 1) ? (uint32_t)(p.g.plane * 16) : 0u);
        r3_sts32(zt + 4 * i, ooff);
      }
      if (lane == 0) {
        r3_sts32(hdr + 4 * R3H_CROP, (uint32_t)crop); r3_sts32(hdr + 4 * R3H_NROWS, (uint32_t)n_out);
        r3_sts32(hdr + 4 * R3H_XBEGIN, (uint32_t)x_begin); r3_sts32(hdr + 4 * R3H_VALID, 2u);
      }
      __syncwarp();
      if (lane == 0) r3_bar_arrive(sb + kR3ItemFull + 8 * b);
      ++it;
      continue;
    }
    const double scale_y = axis_scale(sh, S), scale_x = axis_scale(sw, S);
    float f;
    const int u_first = src_coord(y_begin, scale_y, f) - LO;
    const int u_last = src_coord(y_end - 1, scale_y, f) + HI;
    int n_rows = u_last - u_first + 1;
    if (TAPS == 2) n_rows = (n_rows + 1) & ~1;   // the bilinear consumers take rows in pairs
    // ---- staged segment of a source row: columns [xc0, xc1] of the crop ----
    const int sx_first = src_coord(x_begin, scale_x, f), sx_last = src_coord(x_end - 1, scale_x, f);
    const int xc0 = iclamp(sx_first - LO, 0, sw - 1), xc1 = iclamp(sx_last + HI, 0, sw - 1);
    const int span = xc1 - xc0 + 1;
    const uint32_t pads = TAPS == 8 ? ((sx_first - LO < 0) ? 1u : 0u) | ((sx_last + HI > sw - 1) ? 2u : 0u) : 0u;
    const uint8_t* img0 = p.frames + (long long)frame * p.frame_stride + ((long long)ymin * p.W + xmin + xc0) * 3;
    const uint8_t* msk0 = HAS_MASK ? p.masks + (long long)frame * p.mask_stride + (long long)ymin * p.W + xmin + xc0 : nullptr;
    const uint32_t mis_i = (uint32_t)((uintptr_t)img0 & 15), mis_m = (uint32_t)((uintptr_t)msk0 & 15);
    const uint32_t Li = (mis_i + 3u * (uint32_t)span + 15u) & ~15u;
    const uint32_t Lm = HAS_MASK ? (mis_m + (uint32_t)span + 15u) & ~15u : 0u;
    const uint32_t rgb_off = TAPS == 8 ? (uint32_t)kR3PadL : 0u;                                    // slot offset of the image copy
    const uint32_t msk_off = TAPS == 8 ? rgb_off + Li + (uint32_t)(kR3PadR + kR3PadM) : Li;         // ... of the mask copy
    const uint32_t pitch = TAPS == 8 ? (HAS_MASK ? msk_off + Lm + (uint32_t)kR3PadM : rgb_off + Li + (uint32_t)kR3PadR) : Li + Lm;
    const uint32_t rgb0 = rgb_off + mis_i, msk0o = msk_off + mis_m;    // slot offsets of pixel xc0
    int K = imin(32, p.stage_bytes / (int)pitch);
    if (TAPS == 2) K &= ~1;
    // ---- rows: the first stage is requested before the tables are built (its latency hides behind them) ----
    const long long rb_i = (long long)p.W * 3, rb_m = p.W;
    auto issue = [&](int r0) {
      const int k_this = imin(K, n_rows - r0);
      r3_bar_wait(sb + kR3Empty + 8 * stage, sphase ^ 1u);
      const uint32_t full = sb + kR3Full + 8 * stage;
      if (lane == 0) r3_bar_expect(full, (uint32_t)k_this * (Li + Lm));
      __syncwarp();
      if (lane < k_this) {
        const int r = iclamp(u_first + r0 + lane, 0, sh - 1);
        const uint32_t dst = sb + (uint32_t)p.ring_off + stage * (uint32_t)p.stage_bytes + (uint32_t)lane * pitch;
        r3_bulk_g2s(dst + rgb_off, img0 - mis_i + r * rb_i, Li, full);
        if (HAS_MASK) r3_bulk_g2s(dst + msk_off, msk0 - mis_m + r * rb_m, Lm, full);
      }
      if (++stage == (uint32_t)p.n_stages) { stage = 0; sphase ^= 1u; }
    };
    issue(0);
    // ---- vertical table ----
    const uint32_t ytab = sb + kR3Ytab + r3_ytab_slot(TAPS) * b;
    for (int i = lane; i <= n_out; i += 32) {
      const int y = y_begin + i;
      uint32_t ooff;
      if (p.out_fmt == 0) ooff = (uint32_t)y * (uint32_t)S * 4u;
      else ooff = (uint32_t)(y >> 1) * (uint32_t)p.g.Wp * 16u + ((y & 1) ? (uint32_t)(p.g.plane * 16) : 0u);
      if (TAPS == 2) {
        U32x4 e;
        if (i == n_out) { e.x = 0x7FFFFFFFu; e.y = e.z = e.w = 0u; }
        else {
          short ic[2]; int sy;
          linear_coefs(y, scale_y, sh, true, sy, ic);
          e.x = (uint32_t)(sy + 1); e.w = ooff;
          e.y = __float_as_uint((float)ic[0] * 9.5367431640625e-07f);      // b * 2^-20, exact
          e.z = __float_as_uint((float)ic[1] * 9.5367431640625e-07f);
        }
        r3_sts128(ytab + i * YE, e);
      } else {
        U32x4 e0, e1, e2;
        e0.y = ooff; e0.z = e0.w = 0u;
        e1.x = e1.y = e1.z = e1.w = e2.x = e2.y = e2.z = e2.w = 0u;
        if (i == n_out) e0.x = 0x7FFFFFFFu;
        else {
          short ic[8]; int sy;
          r3_lanczos_entry(p, crop, 1, y, scale_y, sy, ic);
          e0.x = (uint32_t)(sy + 4);
          uint32_t cf[8];
          // ring slot k holds the source row u with (u & 7) == k; the window is u = sy - 3 + j
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = (k - (sy - 3)) & 7;
            int v = 0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) if (jj == j) v = (int)ic[jj];
            cf[k] = (uint32_t)v;
          }
          e1.x = cf[0]; e1.y = cf[1]; e1.z = cf[2]; e1.w = cf[3];
          e2.x = cf[4]; e2.y = cf[5]; e2.z = cf[6]; e2.w = cf[7];
        }
        r3_sts128(ytab + i * YE, e0); r3_sts128(ytab + i * YE + 16, e1); r3_sts128(ytab + i * YE + 32, e2);
      }
    }
    // ---- horizontal table ----
    for (int i = lane; i < p.cols_per_item; i += 32) {
      const uint32_t xa = sb + r3_xtab_off(TAPS) + p.xtab_slot * b + i * (TAPS == 8 ? kR3XtabEntry8 : kR3XtabEntry2);
      if (TAPS == 2) {
        short ic[2]; int sx;
        linear_coefs(x_begin + i, scale_x, sw, false, sx, ic);
        const uint32_t bo = rgb0 + 3u * (uint32_t)(sx - xc0), bm = msk0o + (uint32_t)(sx - xc0);
        U32x4 e;
        e.x = (bo & ~3u) | ((8u * (bo & 3u)) << 16);
        e.y = (uint32_t)(uint16_t)ic[0] | ((uint32_t)(uint16_t)ic[1] << 16);
        e.z = (bm & ~3u) | (((bm & 3u) | (((bm & 3u) + 1u) << 4)) << 16);
        e.w = 0u;
        r3_sts128(xa, e);
      } else {
        short ic[8]; int sx;
        r3_lanczos_entry(p, crop, 0, x_begin + i, scale_x, sx, ic);
        const int rel = sx - 3 - xc0;                  // first tap relative to the first staged pixel (>= -4)
        const uint32_t bo = (uint32_t)((int)rgb0 + 3 * rel), bm = (uint32_t)((int)msk0o + rel);
        U32x4 e0, e1;
        e0.x = (bo & ~3u) | ((8u * (bo & 3u)) << 16);
        e0.y = (bm & ~3u) | ((8u * (bm & 3u)) << 16);
        e0.z = ((pads & 1u) && rel < 0 ? 1u : 0u) | ((pads & 2u) && rel + 7 > span - 1 ? 2u : 0u);
        e0.w = 0u;
        e1.x = (uint32_t)(uint16_t)ic[0] | ((uint32_t)(uint16_t)ic[1] << 16);
        e1.y = (uint32_t)(uint16_t)ic[2] | ((uint32_t)(uint16_t)ic[3] << 16);
        e1.z = (uint32_t)(uint16_t)ic[4] | ((uint32_t)(uint16_t)ic[5] << 16);
        e1.w = (uint32_t)(uint16_t)ic[6] | ((uint32_t)(uint16_t)ic[7] << 16);
        r3_sts128(xa, e0); r3_sts128(xa + 16, e1);
      }
    }
    if (lane == 0) {
      r3_sts32(hdr + 4 * R3H_CROP, (uint32_t)crop); r3_sts32(hdr + 4 * R3H_UFIRST, (uint32_t)u_first);
      r3_sts32(hdr + 4 * R3H_NROWS, (uint32_t)n_rows); r3_sts32(hdr + 4 * R3H_K, (uint32_t)K);
      r3_sts32(hdr + 4 * R3H_PITCH, pitch); r3_sts32(hdr + 4 * R3H_XBEGIN, (uint32_t)x_begin);
      r3_sts32(hdr + 4 * R3H_RGB0, rgb0); r3_sts32(hdr + 4 * R3H_MSK0, msk0o);
      r3_sts32(hdr + 4 * R3H_SPAN, (uint32_t)span); r3_sts32(hdr + 4 * R3H_PADS, pads);
      r3_sts32(hdr + 4 * R3H_VALID, 1u);
    }
    __syncwarp();
    if (lane == 0) r3_bar_arrive(sb + kR3ItemFull + 8 * b);
    for (int r0 = K; r0 < n_rows; r0 += K) issue(r0);
    ++it;
  }
}

// ---------------------------------------------------------------------------------------------
// consumers: shared pieces
// ---------------------------------------------------------------------------------------------
// Normalise + store one pixel.  v4 / m4 hold 4 x the cv2 result in bits 2..9 (other bits arbitrary), `base4` is the
// value of m4's bits above bit 9 (compares run on the raw words).
//   fmt 1: two 32-bit words [c0 c1 | c2 0] of bf16;  bf16_rn(fl(4v * fl(1/1020))) == bf16_rn(fp32(v / 255)) for all v
//   fmt 0: three fp32 words; unmasked values come from the 256-entry table in shared memory
// fp32 output: the branch is warp-uniform (a vote) - a warp with a partially masked pixel (mask edge) takes the general
// quotient for every lane, everything else takes the unmasked path and zeroes its masked-out lanes with a select.
constexpr float kR3MaskedR = (1.0f / 65025.0f) * 0.0625f;     // fl(1/65025) / 16
constexpr float kR3PlainR = 1.0f / 1020.0f;
// PLAIN (fp32 output with a mask only): v4 / m4 hold the plain cv2 result 0..255 instead.
template <bool HAS_MASK, int FMT, bool PLAIN = false>
__device__ __forceinline__ void r3_finish(const uint32_t (&v4)[3], uint32_t m4, uint32_t base4, bool off, uint8_t* q,
                                          uint32_t plane_bytes, uint32_t lut, uint32_t k23, float kn) {
  uint32_t o[3];
  if (FMT == 1 && HAS_MASK) {
    // bf16 output with a mask: ONE branch-free expression for every pixel, two instructions per channel.
    //   mr  = fl(4m * R),  R = fl(1/65025) / 16          (one FFMA on the 2^23 + 4m bit pattern)
    //   out = fl(4v * mr)                                 (one FFMA on the 2^23 + 4v bit pattern: (2^23 + 4v) mr - 2^23 mr, exact, one rounding)
    // bf16_rn(out) equals bf16_rn of the reference's float32((v * (m / 255.0)) / 255.0) for all 65 536 (v, m) pairs (the
    // fp32 values themselves differ below bf16 resolution; tests/test_gpu_roi.py checks the device exhaustively), m = 0
    // gives exactly 0 and m = 255 the unmasked value: no vote, no select, and every warp of a strip does the same work
    // whether the mask edge crosses it or not.
    constexpr float R = kR3MaskedR;                    // == kn, which the caller holds in a register
    uint32_t mb;
    asm("lop3.b32 %0, %1, 0x3FC, %2, 0xEA;" : "=r"(mb) : "r"(m4), "r"(k23));
    const float mr = fmaf(__uint_as_float(mb), kn, -8388608.0f * R);
    const float cc = mr * -8388608.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t vb;
      asm("lop3.b32 %0, %1, 0x3FC, %2, 0xEA;" : "=r"(vb) : "r"(v4[c]), "r"(k23));
      o[c] = __float_as_uint(fmaf(__uint_as_float(vb), mr, cc));
    }
  } else if (HAS_MASK) {
    // fp32 output (the reference's tensor) with a mask: the quotient must be exact in fp32.  normalise_u8(v, m) for every
    // pixel, on X = (4v)(4m) = 16 v m (each step of normalise_u8 scales by an exact power of two: bit-identical, no
    // shifts): m = 0 gives 0, m = 255 the table value.  Branch-free like the bf16 path - the table look-up with its vote,
    // select and bank conflicts only paid off for warps the mask edge does not cross, and those are not the ones a strip
    // waits for.
    const uint32_t mm = PLAIN ? m4 : (m4 & 0x3FCu);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = (float)((PLAIN ? v4[c] : (v4[c] & 0x3FCu)) * mm);
      constexpr float r16 = PLAIN ? 1.0f / 65025.0f : (1.0f / 65025.0f) * 0.0625f;
      const float qq = x * r16;
      o[c] = __float_as_uint(fmaf(fmaf(-qq, PLAIN ? 65025.0f : 16.0f * 65025.0f, x), r16, qq));
    }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (FMT == 0) {
        o[c] = r3_lds32((v4[c] & 0x3FCu) | lut);
      } else {
        // (float)(4v) without a conversion: the integer sits in the mantissa of 2^23 + 4v
        uint32_t xb;
        asm("lop3.b32 %0, %1, 0x3FC, %2, 0xEA;" : "=r"(xb) : "r"(v4[c]), "r"(k23));
        o[c] = __float_as_uint(fmaf(__uint_as_float(xb), kn, -8388608.0f * kR3PlainR));          // kn == kR3PlainR
      }
    }
  }
  if (FMT == 0) {
    *reinterpret_cast<uint32_t*>(q) = o[0];
    *reinterpret_cast<uint32_t*>(q + plane_bytes) = o[1];
    *reinterpret_cast<uint32_t*>(q + 2 * (size_t)plane_bytes) = o[2];
  } else {
    uint2 v;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v.x) : "f"(__uint_as_float(o[1])), "f"(__uint_as_float(o[0])));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v.y) : "f"(0.f), "f"(__uint_as_float(o[2])));
    *reinterpret_cast<uint2*>(q) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear consumer
// ---------------------------------------------------------------------------------------------
struct R3Col2 {
  uint32_t iofs;            // byte offset (from the row slot) of the aligned word that holds the first tap byte
  uint32_t ish;             // 8 * misalignment of the first tap byte
  uint32_t cx;              // a0 | a1 << 16
  uint32_t mofs;            // same for the mask row
  uint32_t msel;            // byte-permute selector of the two mask taps
};
// Horizontally filtered source row for cv2's vertical pass, which consumes A = S >> 4 (S = a0 * p0 + a1 * p1 < 2^20).
// The vertical pass runs on the FMA pipe (see r3_emit2), so the value is kept as the float 2^23 + 16 * A: its bit
// pattern is 0x4B000000 | (S & ~15), one logic operation on the DP2A result.
__device__ __forceinline__ uint32_t r3_f16a(uint32_t S, uint32_t k23) {      // k23 = 0x4B000000, kept in a register by the caller
  uint32_t d;
  asm("lop3.b32 %0, %1, 0xFFFFF0, %2, 0xEA;" : "=r"(d) : "r"(S), "r"(k23));
  return d;
}
template <bool HAS_MASK>
__device__ __forceinline__ void r3_hfilt2(uint32_t wa, uint32_t ma, const R3Col2& k, uint32_t k23, uint32_t (&dst)[HAS_MASK ? 4 : 3]) {
  const uint32_t w0 = r3_lds32(wa), w1 = r3_lds32(wa + 4), w2 = r3_lds32(wa + 8);
  const uint32_t u0 = __funnelshift_r(w0, w1, k.ish), u1 = __funnelshift_r(w1, w2, k.ish);   // a0 a1 a2 b0 | b1 b2 . .
  const uint32_t r1 = __byte_perm(u0, u1, 0x4130);                                           // a0 b0 a1 b1
  const uint32_t r2 = __byte_perm(u0, u1, 0x0052);                                           // a2 b2 . .
  dst[0] = r3_f16a(__dp2a_lo(k.cx, r1, 0u), k23);
  dst[1] = r3_f16a(__dp2a_hi(k.cx, r1, 0u), k23);
  dst[2] = r3_f16a(__dp2a_lo(k.cx, r2, 0u), k23);
  if (HAS_MASK) {
    const uint32_t r3 = __byte_perm(r3_lds32(ma), r3_lds32(ma + 4), k.msel);
    dst[HAS_MASK ? 3 : 0] = r3_f16a(__dp2a_lo(k.cx, r3, 0u), k23);
  }
}
// One output pixel.  cv2's vertical pass is ((b0 * A0 >> 16) + (b1 * A1 >> 16) + 2) >> 2 with two separate truncations.
// Both happen on the FMA pipe: with F = 2^23 + 16 A (r3_hfilt2), bf = b * 2^-20 and round-toward-minus-infinity,
//   fma.rm(bf0, F0, M0)          = M0 + 8 b0 + floor(b0 A0 / 65536)              (exact: one rounding, ulp 1 in [2^23, 2^24))
//   fma.rm(bf1, F1, the above)   = M0 + 8 (b0 + b1) + floor(..) + floor(..)
// so M0 = 1.5 * 2^23 + 2 - 8 (b0 + b1) leaves the bit pattern 0x4B400000 + v4, v4 = 4 x the cv2 result + (0..3).
__device__ __forceinline__ float r3_fma_rm(float a, float b, float c) {
  float d;
  asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <bool HAS_MASK, int FMT>
__device__ __forceinline__ void r3_emit2(const uint32_t (&lo)[HAS_MASK ? 4 : 3], const uint32_t (&up)[HAS_MASK ? 4 : 3],
                                         const U32x4& yt, uint8_t* out, uint32_t plane_bytes, uint32_t lut, uint32_t k23, float kn) {
  constexpr int NCH = HAS_MASK ? 4 : 3;
  constexpr uint32_t kBase = 0x4B400000u;    // 1.5 * 2^23
  const float b0 = __uint_as_float(yt.y), b1 = __uint_as_float(yt.z);
  // b0 + b1 == 2048 for every destination row of every source size up to 16384 and every S this kernel is launched
  // with (S % 32 == 0, S <= 512): the vertical fraction is within 2^-27 of a rational with denominator 2S, which is never
  // closer than 1/lcm(2S, 4096) >= 2^-22 to a rounding tie of 2048 * frac, so the two roundings are complementary;
  // tests/test_roi_coefs.py checks all of them.  Hence M0 is a constant.
  constexpr float m0 = 12582914.0f - 8.0f * 2048.0f;
  uint32_t m4 = kBase + 1020u;
  if (HAS_MASK) m4 = __float_as_uint(r3_fma_rm(b1, __uint_as_float(up[NCH - 1]), r3_fma_rm(b0, __uint_as_float(lo[NCH - 1]), m0)));
  const bool off = HAS_MASK && m4 < kBase + 4u;      // masked out: exactly zero whatever the image holds
  // (No shortcut for warps whose 32 pixels are all masked out: the warps of a CTA advance stage by stage together, so
  //  the time of a strip is the time of its busiest warp and the vote would only add instructions to that one.)
  uint32_t v4[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) v4[j] = __float_as_uint(r3_fma_rm(b1, __uint_as_float(up[j]), r3_fma_rm(b0, __uint_as_float(lo[j]), m0)));
  r3_finish<HAS_MASK, FMT>(v4, m4, kBase, off, out + yt.w, plane_bytes, lut, k23, kn);
}

template <bool HAS_MASK, int FMT>
__device__ __forceinline__ void r3_consumer2(const Roi3Params& p, uint32_t sb, int t, int lane) {
  constexpr int NCH = HAS_MASK ? 4 : 3, TAPS = 2;
  const int S = p.S;
  const uint32_t lut = sb + r3_lut_off(TAPS);
  const uint32_t ring = sb + (uint32_t)p.ring_off;
  const uint32_t bars_end = sb + kR3Full + 8u * (uint32_t)p.n_stages;
  uint32_t stage_row = ring, full = sb + kR3Full, sphase = 0;    // ring position; `empty` sits kR3Empty behind `full`
  uint8_t* out0;
  uint32_t plane_bytes = 0u;
  if (FMT == 0) {
    out0 = reinterpret_cast<uint8_t*>(p.out) + (long long)t * 4;
    plane_bytes = (uint32_t)S * (uint32_t)S * 4u;
  } else {
    out0 = reinterpret_cast<uint8_t*>(p.out) + (long long)p.g.base * 16 + (long long)t * 8;   // (x >> 1) * 16 + (x & 1) * 8 == x * 8
  }
  const long long crop_bytes = FMT == 0 ? 3LL * S * S * 4 : (long long)p.g.Hp * p.g.Wp * 16;
  const int col_bytes = FMT == 0 ? 4 : 8;
  const uint32_t k23 = r3_lds32(sb + kR3Konst + 8);   // the bit pattern of 2^23; read from shared memory so that it stays in a register
  const float kn = __uint_as_float(r3_lds32(sb + kR3Konst + (HAS_MASK ? 0 : 4)));   // the normalise factor of r3_finish (bf16 output)
  for (int it = 0;; ++it) {
    const int b = it & 1;
    r3_bar_wait(sb + kR3ItemFull + 8 * b, ((uint32_t)it >> 1) & 1u);
    const uint32_t hdr = sb + kR3Hdr + 64 * b;
    const U32x4 h0 = r3_lds128(hdr), h1 = r3_lds128(hdr + 16);
    if (h0.x == 0u) return;
    if (h0.x == 2u) {                          // zero item (empty box): write zeros to the item's rows of this column
      uint8_t* zo = out0 + (int)h0.y * crop_bytes + (int)h1.z * col_bytes;
      const uint32_t zt = sb + kR3Ytab + r3_ytab_slot(TAPS) * b;
      for (int i = 0; i < (int)h0.w; ++i) {
        uint8_t* q = zo + r3_lds32(zt + 4 * i);
        if (FMT == 0) {
          *reinterpret_cast<uint32_t*>(q) = 0u;
          *reinterpret_cast<uint32_t*>(q + plane_bytes) = 0u;
          *reinterpret_cast<uint32_t*>(q + 2 * (size_t)plane_bytes) = 0u;
        } else {
          *reinterpret_cast<uint2*>(q) = make_uint2(0u, 0u);
        }
      }
      __syncwarp();
      if (lane == 0) r3_bar_arrive(sb + kR3ItemEmpty + 8 * b);
      continue;
    }
    int u = (int)h0.z, rows_left = (int)h0.w;
    const int K = (int)h1.x;
    const uint32_t pitch = h1.y;
    R3Col2 col;
    {
      const U32x4 e = r3_lds128(sb + r3_xtab_off(TAPS) + p.xtab_slot * b + kR3XtabEntry2 * t);
      col.iofs = e.x & 0xFFFFu; col.ish = e.x >> 16; col.cx = e.y; col.mofs = e.z & 0xFFFFu; col.msel = e.z >> 16;
    }
    uint8_t* out = out0 + (int)h0.y * crop_bytes + (int)h1.z * col_bytes;
    asm volatile("" : "+l"(out));            // one 64-bit register pair: the emit adds the row offset to it, nothing else
    uint32_t A[NCH], B[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) A[j] = B[j] = 0u;
    uint32_t ya = sb + kR3Ytab + r3_ytab_slot(TAPS) * b;
    U32x4 yt = r3_lds128(ya);
    while (rows_left > 0) {
      const int k_this = imin(K, rows_left);
      rows_left -= k_this;
      r3_bar_wait(full, sphase);
      uint32_t wa = stage_row + col.iofs, ma = stage_row + col.mofs;      // this column's words in the current row
      const int u_end = u + k_this;
      while (u != u_end) {
        r3_hfilt2<HAS_MASK>(wa, ma, col, k23, A);
        while ((int)yt.x == u) {             // output rows whose upper source row is u
          r3_emit2<HAS_MASK, FMT>(B, A, yt, out, plane_bytes, lut, k23, kn);
          ya += kR3YtabEntry2;
          yt = r3_lds128(ya);
        }
        ++u; wa += pitch; ma += pitch;
        r3_hfilt2<HAS_MASK>(wa, ma, col, k23, B);
        while ((int)yt.x == u) {
          r3_emit2<HAS_MASK, FMT>(A, B, yt, out, plane_bytes, lut, k23, kn);
          ya += kR3YtabEntry2;
          yt = r3_lds128(ya);
        }
        ++u; wa += pitch; ma += pitch;
      }
      __syncwarp();
      if (lane == 0) r3_bar_arrive(full + kR3Empty);
      stage_row += (uint32_t)p.stage_bytes; full += 8u;
      if (full == bars_end) { stage_row = ring; full = sb + kR3Full; sphase ^= 1u; }
    }
    __syncwarp();
    if (lane == 0) r3_bar_arrive(sb + kR3ItemEmpty + 8 * b);
  }
}

// ---------------------------------------------------------------------------------------------
// Lanczos4 consumer
// ---------------------------------------------------------------------------------------------
struct R3Col8 {
  uint32_t iofs, ish, mofs, msh;
  uint32_t c01, c23, c45, c67;
};
template <bool HAS_MASK>
__device__ __forceinline__ void r3_hfilt8(uint32_t row, const R3Col8& k, int (&dst)[HAS_MASK ? 4 : 3]) {
  const uint32_t wa = row + k.iofs;
  uint32_t u[6];
  {
    uint32_t wv[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) wv[i] = r3_lds32(wa + 4 * i);
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = __funnelshift_r(wv[i], wv[i + 1], k.ish);   // bytes 0..23 = 8 pixels x 3 channels
  }
  int a0, a1, a2;
  uint32_t r;
  r = __byte_perm(u[0], u[1], 0x4130); a0 = r3_dp2a_lo_su(k.c01, r, 0); a1 = r3_dp2a_hi_su(k.c01, r, 0);     // taps 0,1: bytes (0,3) (1,4)
  r = __byte_perm(u[0], u[1], 0x0052); a2 = r3_dp2a_lo_su(k.c01, r, 0);                                      //           bytes (2,5)
  r = __byte_perm(u[1], u[2], 0x6352); a0 = r3_dp2a_lo_su(k.c23, r, a0); a1 = r3_dp2a_hi_su(k.c23, r, a1);   // taps 2,3: bytes (6,9) (7,10)
  r = __byte_perm(u[1], u[2], 0x0074); a2 = r3_dp2a_lo_su(k.c23, r, a2);                                     //           bytes (8,11)
  r = __byte_perm(u[3], u[4], 0x4130); a0 = r3_dp2a_lo_su(k.c45, r, a0); a1 = r3_dp2a_hi_su(k.c45, r, a1);   // taps 4,5: bytes (12,15) (13,16)
  r = __byte_perm(u[3], u[4], 0x0052); a2 = r3_dp2a_lo_su(k.c45, r, a2);
  r = __byte_perm(u[4], u[5], 0x6352); a0 = r3_dp2a_lo_su(k.c67, r, a0); a1 = r3_dp2a_hi_su(k.c67, r, a1);   // taps 6,7: bytes (18,21) (19,22)
  r = __byte_perm(u[4], u[5], 0x0074); a2 = r3_dp2a_lo_su(k.c67, r, a2);
  dst[0] = a0; dst[1] = a1; dst[2] = a2;
  if (HAS_MASK) {
    const uint32_t ma = row + k.mofs;
    const uint32_t m0 = r3_lds32(ma), m1 = r3_lds32(ma + 4), m2 = r3_lds32(ma + 8);
    const uint32_t t0 = __funnelshift_r(m0, m1, k.msh), t1 = __funnelshift_r(m1, m2, k.msh);
    int a3 = r3_dp2a_lo_su(k.c01, t0, 0);
    a3 = r3_dp2a_hi_su(k.c23, t0, a3);
    a3 = r3_dp2a_lo_su(k.c45, t1, a3);
    a3 = r3_dp2a_hi_su(k.c67, t1, a3);
    dst[HAS_MASK ? 3 : 0] = a3;
  }
}
// replicated border of the crop: 4 pixels in front of / behind the staged segment of every row of a stage
template <bool HAS_MASK>
__device__ __forceinline__ void r3_patch8(uint32_t stage_row, int k_this, uint32_t pitch, uint32_t rgb0, uint32_t msk0, int span,
                                          bool left, bool right, int lane) {
  for (int i = lane; i < 4 * k_this; i += 32) {
    const uint32_t rowa = stage_row + (uint32_t)(i >> 2) * pitch;
    const int j = (i & 3) + 1;
    if (left) {
      const uint32_t px = rowa + rgb0;
      const uint32_t c0 = r3_lds8(px), c1 = r3_lds8(px + 1), c2 = r3_lds8(px + 2);
      r3_sts8(px - 3 * j, c0); r3_sts8(px - 3 * j + 1, c1); r3_sts8(px - 3 * j + 2, c2);
      if (HAS_MASK) r3_sts8(rowa + msk0 - j, r3_lds8(rowa + msk0));
    }
    if (right) {
      const uint32_t px = rowa + rgb0 + 3u * (uint32_t)(span - 1);
      const uint32_t c0 = r3_lds8(px), c1 = r3_lds8(px + 1), c2 = r3_lds8(px + 2);
      r3_sts8(px + 3 * j, c0); r3_sts8(px + 3 * j + 1, c1); r3_sts8(px + 3 * j + 2, c2);
      if (HAS_MASK) r3_sts8(rowa + msk0 + (uint32_t)(span - 1 + j), r3_lds8(rowa + msk0 + (uint32_t)(span - 1)));
    }
  }
}

template <bool HAS_MASK, int FMT>
__device__ __forceinline__ void r3_consumer8(const Roi3Params& p, uint32_t sb, int t, int lane) {
  constexpr int NCH = HAS_MASK ? 4 : 3, TAPS = 8;
  const int S = p.S;
  const uint32_t lut = sb + r3_lut_off(TAPS);
  const uint32_t ring = sb + (uint32_t)p.ring_off;
  const uint32_t bars_end = sb + kR3Full + 8u * (uint32_t)p.n_stages;
  uint32_t stage_row = ring, full = sb + kR3Full, sphase = 0;
  uint8_t* out0;
  uint32_t plane_bytes = 0u;
  if (FMT == 0) {
    out0 = reinterpret_cast<uint8_t*>(p.out) + (long long)t * 4;
    plane_bytes = (uint32_t)S * (uint32_t)S * 4u;
  } else {
    out0 = reinterpret_cast<uint8_t*>(p.out) + (long long)p.g.base * 16 + (long long)t * 8;
  }
  const long long crop_bytes = FMT == 0 ? 3LL * S * S * 4 : (long long)p.g.Hp * p.g.Wp * 16;
  const int col_bytes = FMT == 0 ? 4 : 8;
  for (int it = 0;; ++it) {
    const int b = it & 1;
    r3_bar_wait(sb + kR3ItemFull + 8 * b, ((uint32_t)it >> 1) & 1u);
    const uint32_t hdr = sb + kR3Hdr + 64 * b;
    const U32x4 h0 = r3_lds128(hdr), h1 = r3_lds128(hdr + 16), h2 = r3_lds128(hdr + 32);
    if (h0.x == 0u) return;
    if (h0.x == 2u) {                          // zero item (empty box): write zeros to the item's rows of this column
      uint8_t* zo = out0 + (int)h0.y * crop_bytes + (int)h1.z * col_bytes;
      const uint32_t zt = sb + kR3Ytab + r3_ytab_slot(TAPS) * b;
      for (int i = 0; i < (int)h0.w; ++i) {
        uint8_t* q = zo + r3_lds32(zt + 4 * i);
        if (FMT == 0) {
          *reinterpret_cast<uint32_t*>(q) = 0u;
          *reinterpret_cast<uint32_t*>(q + plane_bytes) = 0u;
          *reinterpret_cast<uint32_t*>(q + 2 * (size_t)plane_bytes) = 0u;
        } else {
          *reinterpret_cast<uint2*>(q) = make_uint2(0u, 0u);
        }
      }
      __syncwarp();
      if (lane == 0) r3_bar_arrive(sb + kR3ItemEmpty + 8 * b);
      continue;
    }
    int u = (int)h0.z, rows_left = (int)h0.w;
    const int K = (int)h1.x;
    const uint32_t pitch = h1.y, rgb0 = h1.w, msk0 = h2.x;
    const int span = (int)h2.y;
    R3Col8 col;
    bool patch_l, patch_r;
    {
      const uint32_t xa = sb + r3_xtab_off(TAPS) + p.xtab_slot * b + kR3XtabEntry8 * t;
      const U32x4 e0 = r3_lds128(xa), e1 = r3_lds128(xa + 16);
      col.iofs = e0.x & 0xFFFFu; col.ish = e0.x >> 16; col.mofs = e0.y & 0xFFFFu; col.msh = e0.y >> 16;
      col.c01 = e1.x; col.c23 = e1.y; col.c45 = e1.z; col.c67 = e1.w;
      patch_l = __any_sync(0xFFFFFFFFu, e0.z & 1u);
      patch_r = __any_sync(0xFFFFFFFFu, e0.z & 2u);
    }
    uint8_t* out = out0 + (int)h0.y * crop_bytes + (int)h1.z * col_bytes;
    asm volatile("" : "+l"(out));
    int R[8][NCH];                           // ring: slot (u & 7) holds filtered source row u
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < NCH; ++j) R[i][j] = 0;
    uint32_t ya = sb + kR3Ytab + r3_ytab_slot(TAPS) * b;
    U32x4 yt = r3_lds128(ya);
    while (rows_left > 0) {
      const int k_this = imin(K, rows_left);
      rows_left -= k_this;
      r3_bar_wait(full, sphase);
      if (patch_l || patch_r) {
        r3_patch8<HAS_MASK>(stage_row, k_this, pitch, rgb0, msk0, span, patch_l, patch_r, lane);
        __syncwarp();
      }
      uint32_t row = stage_row;
      for (int k = 0; k < k_this; ++k, ++u, row += pitch) {
        switch (u & 7) {
          case 0: r3_hfilt8<HAS_MASK>(row, col, R[0]); break;
          case 1: r3_hfilt8<HAS_MASK>(row, col, R[1]); break;
          case 2: r3_hfilt8<HAS_MASK>(row, col, R[2]); break;
          case 3: r3_hfilt8<HAS_MASK>(row, col, R[3]); break;
          case 4: r3_hfilt8<HAS_MASK>(row, col, R[4]); break;
          case 5: r3_hfilt8<HAS_MASK>(row, col, R[5]); break;
          case 6: r3_hfilt8<HAS_MASK>(row, col, R[6]); break;
          default: r3_hfilt8<HAS_MASK>(row, col, R[7]); break;
        }
        while ((int)yt.x == u) {             // output rows whose last source row is u
          const U32x4 ca = r3_lds128(ya + 16), cb = r3_lds128(ya + 32);
          const int cf[8] = {(int)ca.x, (int)ca.y, (int)ca.z, (int)ca.w, (int)cb.x, (int)cb.y, (int)cb.z, (int)cb.w};
          // 4 x the cv2 result, clamped: ((acc + 2^21) >> 22 clamped to 0..255) * 4 == ((acc + 2^21) >> 20 clamped to 0..1023) & 0x3FC
          // (what the table look-up and the bf16 bit patterns index with); the masked fp32 quotient takes the plain 0..255
          constexpr bool PLAIN = HAS_MASK && FMT == 0;
          auto vpass = [&](int j) {
            int acc = 1 << 21;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += R[i][j] * cf[i];
            return (uint32_t)__vimin_s32_relu(acc >> (PLAIN ? 22 : 20), PLAIN ? 255 : 1023);   // max(min(x, hi), 0) in one VIMNMX.RELU
          };
          const uint32_t m4 = HAS_MASK ? vpass(NCH - 1) : 1020u;
          const bool off = HAS_MASK && m4 < 4u;
          uint8_t* q = out + yt.y;
          uint32_t v4[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) v4[j] = vpass(j);
          r3_finish<HAS_MASK, FMT, PLAIN>(v4, m4, 0u, off, q, plane_bytes, lut, 0x4B000000u, HAS_MASK ? kR3MaskedR : kR3PlainR);
          ya += kR3YtabEntry8;
          yt = r3_lds128(ya);
        }
      }
      __syncwarp();
      if (lane == 0) r3_bar_arrive(full + kR3Empty);
      stage_row += (uint32_t)p.stage_bytes; full += 8u;
      if (full == bars_end) { stage_row = ring; full = sb + kR3Full; sphase ^= 1u; }
    }
    __syncwarp();
    if (lane == 0) r3_bar_arrive(sb + kR3ItemEmpty + 8 * b);
  }
}

// ---------------------------------------------------------------------------------------------
// kernel: blockDim = cols_per_item + 32; consumer thread t owns output column x_begin + t of the item
// ---------------------------------------------------------------------------------------------
template <int TAPS, bool HAS_MASK, int FMT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) roi3_kernel(const __grid_constant__ Roi3Params p) {
  extern __shared__ __align__(1024) uint8_t r3_smem[];
  const uint32_t sb = smem_u32(r3_smem);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n_cons = (int)(blockDim.x >> 5) - 1;
  if (t == 0) {
    if (sb & 1023u) __trap();
    for (int s = 0; s < p.n_stages; ++s) { r3_bar_init(sb + kR3Full + 8 * s, 1u); r3_bar_init(sb + kR3Empty + 8 * s, (uint32_t)n_cons); }
    for (int b = 0; b < 2; ++b) { r3_bar_init(sb + kR3ItemFull + 8 * b, 1u); r3_bar_init(sb + kR3ItemEmpty + 8 * b, (uint32_t)n_cons); }
    r3_sts32(sb + kR3Konst, f32_bits(kR3MaskedR));
    r3_sts32(sb + kR3Konst + 4, f32_bits(kR3PlainR));
    r3_sts32(sb + kR3Konst + 8, 0x4B000000u);
    mbar_fence_init();
  }
  if (FMT == 0) for (int i = t; i < 256; i += blockDim.x) r3_sts32(sb + r3_lut_off(TAPS) + 4 * i, f32_bits(normalise_u8(i, 255)));
  __syncthreads();
  if (warp == n_cons) r3_producer<TAPS, HAS_MASK>(p, sb, lane);
  else if (TAPS == 8) r3_consumer8<HAS_MASK, FMT>(p, sb, t, lane);
  else r3_consumer2<HAS_MASK, FMT>(p, sb, t, lane);
}
#endif  // __CUDACC__

}  // namespace flope

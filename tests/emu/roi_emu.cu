// Host emulation of the staged ROI kernels (flope_b200/csrc/roi_crop.cuh) - TEST INFRASTRUCTURE ONLY.
//
// The thread programs of roi2_kernel are plain functions of (shared-memory image, thread index).  This file
// compiles them for the host and walks the grid sequentially, phase by phase, with memcpy in place of the TMA
// bulk copies, so that the index arithmetic (aligned-window loads, funnel shifts, byte permutes, DP2A pairing,
// ring rotation, chunking, border replication) can be checked bit for bit against cv2 on a machine without a
// GPU (tests/test_roi_emu.py).  Nothing under flope_b200/ links or loads this; the product path is the CUDA
// kernel and fails loudly without a GPU.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../flope_b200/csrc/roi_crop.cuh"

using namespace flope;

template <int TAPS, bool HAS_MASK, int FMT, bool LUT>
static int emu_cta(const RoiParams& p, int bx, int by, int bz, int nt) {
  Roi2Cta c;
  if (!roi2_cta_init<TAPS>(p, bx, by, bz, HAS_MASK, c)) return 0;
  const int ncols = c.x_end - c.x_begin, nrows = c.y_end - c.y_begin;
  const Roi2Layout L = roi2_layout(TAPS, p.cols_cta, p.rows_per_strip, p.data_bytes);
  std::vector<uint8_t> buf(L.total + 1024, 0xCD);
  uint8_t* smem = buf.data() + ((1024 - ((uintptr_t)buf.data() & 1023)) & 1023);      // 1 KB aligned like the device buffer
  const SmemHost s{smem};
  s.st32(kRoi2ZeroWord, 0u);
  for (int i = 0; i < ncols; ++i) { if (TAPS == 8) roi2_fill_xtab8(c, s, L, i); else roi2_fill_xtab2(c, s, L, i); }
  for (int i = 0; i < nrows; ++i) { if (TAPS == 8) roi2_fill_ytab8(c, s, L, i); else roi2_fill_ytab2(c, s, L, i); }
  for (int i = 0; i < 256; ++i) roi2_fill_lut(s, L, i, FMT);
  roi2_plan_chunks<TAPS>(c, s, L);
  const int* plan = reinterpret_cast<const int*>(smem + L.plan);
  const int n_chunks = plan[0];
  int ya = c.y_begin;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int yb = plan[1 + ch];
    int r_lo, r_hi;
    roi2_row_span<TAPS>(c, s, L, ya, yb, r_lo, r_hi);
    const int n_rows = r_hi - r_lo + 1;
    if (n_rows > c.rows_fit && yb - ya > 1) { fprintf(stderr, "emu: chunk does not fit\n"); return -1; }
    memset(smem + L.data, 0xCD, p.data_bytes);
    for (int slot = 0; slot < n_rows; ++slot) {
      Roi2RowCopy ci, cm;
      U32x2 rt;
      roi2_stage_desc(p, c, L, HAS_MASK, r_lo, n_rows, slot, ci, cm, rt);
      rt.x = (uint32_t)s.abs((int)rt.x); rt.y = (uint32_t)s.abs((int)rt.y);
      memcpy(smem + L.rowtab + slot * 8, &rt, 8);
      const Roi2RowCopy* cs[2] = {&ci, &cm};
      for (int k = 0; k < (HAS_MASK ? 2 : 1); ++k) {
        const Roi2RowCopy& r = *cs[k];
        if (r.bulk && (((uintptr_t)r.src & 15) || (r.dst & 15) || (r.bytes & 15))) { fprintf(stderr, "emu: misaligned bulk copy\n"); return -2; }
        if ((int)(r.dst + r.bytes) > L.total) { fprintf(stderr, "emu: staged row overruns shared memory\n"); return -3; }
        memcpy(smem + r.dst, r.src, r.bytes);
      }
    }
    if (TAPS == 8 && (c.pad_l || c.pad_r))
      for (int i = 0; i < 2 * n_rows; ++i) roi2_patch_row(c, s, L, HAS_MASK, i >> 1, i & 1);
    for (int t = 0; t < nt; ++t) {
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
    }
    ya = yb;
  }
  return n_chunks;
}

extern "C" int roi_emu(const uint8_t* frames, int n_frames, int H, int W, const uint8_t* masks, const int32_t* boxes, int n,
                       int S, int interp, int fmt, void* out, const int* geom /*C,H,W,Hp,Wp,base*/, long long plane,
                       int rows_per_strip, int cols_cta, int n_sub, int data_bytes, int block, int lut, int* max_chunks) {
  RoiParams p{};
  p.frames = frames; p.frame_stride = (long long)H * W * 3; p.masks = masks; p.mask_stride = (long long)H * W;
  p.H = H; p.W = W; p.boxes = boxes; p.n = n; p.S = S; p.out_fmt = fmt; p.out = out;
  p.g.C = geom[0]; p.g.H = geom[1]; p.g.W = geom[2]; p.g.Hp = geom[3]; p.g.Wp = geom[4]; p.g.base = geom[5]; p.g.plane = plane;
  p.rows_per_strip = rows_per_strip; p.cols_cta = cols_cta; p.n_sub = n_sub; p.data_bytes = data_bytes;
  p.frames_end = frames + (long long)n_frames * p.frame_stride;
  p.masks_end = masks ? masks + (long long)n_frames * p.mask_stride : nullptr;
  const bool hm = masks != nullptr;
  int mc = 0;
  for (int bz = 0; bz < n; ++bz)
    for (int by = 0; by < (S + rows_per_strip - 1) / rows_per_strip; ++by)
      for (int bx = 0; bx < (S + cols_cta - 1) / cols_cta; ++bx) {
        int rc;
        const int key = (interp ? 4 : 0) | (hm ? 2 : 0) | fmt;
        switch (key) {
          case 0: rc = lut ? emu_cta<2, false, 0, true>(p, bx, by, bz, block) : emu_cta<2, false, 0, false>(p, bx, by, bz, block); break;
          case 1: rc = lut ? emu_cta<2, false, 1, true>(p, bx, by, bz, block) : emu_cta<2, false, 1, false>(p, bx, by, bz, block); break;
          case 2: rc = lut ? emu_cta<2, true, 0, true>(p, bx, by, bz, block) : emu_cta<2, true, 0, false>(p, bx, by, bz, block); break;
          case 3: rc = lut ? emu_cta<2, true, 1, true>(p, bx, by, bz, block) : emu_cta<2, true, 1, false>(p, bx, by, bz, block); break;
          case 4: rc = lut ? emu_cta<8, false, 0, true>(p, bx, by, bz, block) : emu_cta<8, false, 0, false>(p, bx, by, bz, block); break;
          case 5: rc = lut ? emu_cta<8, false, 1, true>(p, bx, by, bz, block) : emu_cta<8, false, 1, false>(p, bx, by, bz, block); break;
          case 6: rc = lut ? emu_cta<8, true, 0, true>(p, bx, by, bz, block) : emu_cta<8, true, 0, false>(p, bx, by, bz, block); break;
          default: rc = lut ? emu_cta<8, true, 1, true>(p, bx, by, bz, block) : emu_cta<8, true, 1, false>(p, bx, by, bz, block); break;
        }
        if (rc < 0) return rc;
        if (rc > mc) mc = rc;
      }
  if (max_chunks) *max_chunks = mc;
  return 0;
}

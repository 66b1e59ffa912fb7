// flope_b200: shift-GEMM convolution on tcgen05 / TMEM fed by TMA bulk copies (sm_100a).
//
// Replaces every cuDNN/cuBLAS call site of PoseResNet.forward
// (sunflower/models/posenet.py:24-34: base.conv1, layer1..4 convs, downsample convs,
// base.fc) - SURVEY.md section 2c, K2/K3/K5.
//
// One CTA computes TM = MT*128 consecutive pixel positions x N_TILE output channels.
// In the blocked-pixel layout (common.cuh) a filter tap is a constant position shift, so
// the A operand of tap (dy,dx) is the *same* shared-memory halo tile read through a UMMA
// descriptor whose start address is moved by shift*16 bytes: im2col costs nothing and
// the halo tile is fetched once per 64-channel group instead of once per tap.
//   warp 0   : TMA producer (cp.async.bulk -> mbarrier complete_tx)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-5: epilogue (tcgen05.ld -> folded BN scale/bias, residual, ReLU -> bf16/fp32 store)
#pragma once
#include "common.cuh"

namespace flope {

constexpr int kMaxGroups = 32;
constexpr int kMaxTaps = 16;
constexpr int kConvThreads = 192;

enum OutMode : int { OUT_PLAIN = 0, OUT_PARITY = 1, OUT_F32_ROWS = 2 };

struct ConvParams {
  // ---- A operand (activations, blocked-pixel bf16) ----
  const __nv_bfloat16* in;
  long long in_plane;          // pixels per plane
  int in_base;                 // guard pixels in front of position 0
  int n_groups;                // K groups (each kc8 planes = kc8*8 channels)
  int kc8;                     // planes per group: 8 (64 channels) or 2 (stem, 16 channels)
  int halo_before, halo_after; // pixels needed before / after the tile for all taps
  int group_plane[kMaxGroups]; // first input plane of each group
  int group_tapofs[kMaxGroups];
  int group_ntaps[kMaxGroups];
  int tap_shift[kMaxTaps];     // position shift per tap-table entry
  int taps_total;              // sum of group_ntaps = weight tiles per N tile
  // ---- B operand (weights, packed [n_tile][tile][kc8][N_TILE][8] bf16) ----
  const __nv_bfloat16* wgt;
  // ---- position space (validity + (n,h,w) decode) ----
  long long n_positions;       // N*Hp*Wp
  int Hp, Wp, H, W;
  // ---- epilogue ----
  const float* scale;          // [Cout] folded BN scale (1 for fc)
  const float* bias;           // [Cout] folded BN bias / fc bias
  int relu;
  int out_mode;
  int Cout;
  void* out;
  long long out_plane;
  int out_base, out_Hp, out_Wp;        // for OUT_PARITY these describe the half-resolution grid
  const __nv_bfloat16* res;            // residual (plain layout) or nullptr
  long long res_plane;
  int res_base, res_Hp, res_Wp;
  // ---- smem ring sizes ----
  int n_a_slots, n_b_slots;
};

template <int N_TILE, int MT>
struct ConvSmem {
  static constexpr int TM = MT * 128;
  static size_t a_plane_bytes(int halo) { return (size_t)(TM + halo) * 16; }
  static size_t bytes(int halo, int kc8, int n_a, int n_b) {
    return 1024 /*barriers + align slack*/ + (size_t)n_a * kc8 * a_plane_bytes(halo) + (size_t)n_b * kc8 * N_TILE * 16 +
           2 * N_TILE * sizeof(float);
  }
};

template <int N_TILE, int MT>
__global__ void __launch_bounds__(kConvThreads, 2) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  constexpr int TM = MT * 128;
  constexpr int TMEM_COLS = (N_TILE * MT < 32) ? 32 : N_TILE * MT;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
  constexpr uint32_t IDESC = umma_idesc_bf16(128, N_TILE);

  extern __shared__ __align__(128) uint8_t smem_raw[];
  // carve: [barriers 512 B][tmem ptr][scale][bias][A ring][B ring]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_empty = a_full + 8;
  uint64_t* b_full = a_empty + 8;
  uint64_t* b_empty = b_full + 16;
  uint64_t* acc_full = b_empty + 16;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_scale = reinterpret_cast<float*>(smem_raw + 512);
  float* s_bias = s_scale + N_TILE;
  const int halo = p.halo_before + p.halo_after;
  const uint32_t a_plane_bytes = (uint32_t)(TM + halo) * 16u;
  const uint32_t a_slot_bytes = a_plane_bytes * p.kc8;
  const uint32_t b_tile_bytes = (uint32_t)p.kc8 * N_TILE * 16u;
  uint8_t* a_ring = smem_raw + 512 + 2 * N_TILE * sizeof(float);
  a_ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a_ring) + 127) & ~uintptr_t(127));
  uint8_t* b_ring = a_ring + (size_t)p.n_a_slots * a_slot_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long tile_start = (long long)blockIdx.x * TM;
  const int n_tile = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_a_slots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.n_b_slots; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < N_TILE; i += 128) {
      s_scale[i] = p.scale[n_tile * N_TILE + i];
      s_bias[i] = p.bias[n_tile * N_TILE + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const __nv_bfloat16* wtile = p.wgt + (size_t)n_tile * p.taps_total * (b_tile_bytes / 2);
      int b_idx = 0;
      for (int g = 0; g < p.n_groups; ++g) {
        const int a_slot = g % p.n_a_slots;
        const uint32_t a_phase = (g / p.n_a_slots) & 1;
        mbar_wait(&a_empty[a_slot], a_phase ^ 1);
        mbar_expect_tx(&a_full[a_slot], a_slot_bytes);
        uint8_t* a_dst = a_ring + (size_t)a_slot * a_slot_bytes;
        for (int j = 0; j < p.kc8; ++j) {
          const __nv_bfloat16* src =
              p.in + ((long long)(p.group_plane[g] + j) * p.in_plane + p.in_base + tile_start - p.halo_before) * 8;
          bulk_g2s(a_dst + (size_t)j * a_plane_bytes, src, a_plane_bytes, &a_full[a_slot]);
        }
        for (int t = 0; t < p.group_ntaps[g]; ++t, ++b_idx) {
          const int b_slot = b_idx % p.n_b_slots;
          const uint32_t b_phase = (b_idx / p.n_b_slots) & 1;
          mbar_wait(&b_empty[b_slot], b_phase ^ 1);
          mbar_expect_tx(&b_full[b_slot], b_tile_bytes);
          bulk_g2s(b_ring + (size_t)b_slot * b_tile_bytes, wtile + (size_t)b_idx * (b_tile_bytes / 2), b_tile_bytes,
                   &b_full[b_slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // descriptor strides: A planes are a_plane_bytes apart (K direction), 8-pixel groups 128 B apart
      const uint64_t a_desc0 = umma_desc(0, /*lbo=*/a_plane_bytes, /*sbo=*/128);
      const uint64_t b_desc0 = umma_desc(0, /*lbo=*/N_TILE * 16, /*sbo=*/128);
      const uint32_t a_ring_addr = smem_u32(a_ring);
      const uint32_t b_ring_addr = smem_u32(b_ring);
      const int kpairs = p.kc8 >> 1;
      int b_idx = 0;
      uint32_t first = 1;
      for (int g = 0; g < p.n_groups; ++g) {
        const int a_slot = g % p.n_a_slots;
        const uint32_t a_phase = (g / p.n_a_slots) & 1;
        mbar_wait(&a_full[a_slot], a_phase);
        const uint32_t a_base = a_ring_addr + a_slot * a_slot_bytes + (uint32_t)p.halo_before * 16u;
        const int tofs = p.group_tapofs[g];
        for (int t = 0; t < p.group_ntaps[g]; ++t, ++b_idx) {
          const int b_slot = b_idx % p.n_b_slots;
          const uint32_t b_phase = (b_idx / p.n_b_slots) & 1;
          mbar_wait(&b_full[b_slot], b_phase);
          tc_fence_after();
          const uint32_t a_tap = a_base + (uint32_t)(p.tap_shift[tofs + t] * 16);
          const uint32_t b_base = b_ring_addr + b_slot * b_tile_bytes;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            for (int k = 0; k < kpairs; ++k) {
              const uint32_t a_addr = a_tap + (uint32_t)mt * 2048u + (uint32_t)(2 * k) * a_plane_bytes;
              const uint32_t b_addr = b_base + (uint32_t)(2 * k) * (N_TILE * 16u);
              umma_bf16(tmem_base + mt * N_TILE, a_desc0 | (uint64_t)((a_addr >> 4) & 0x3FFF),
                        b_desc0 | (uint64_t)((b_addr >> 4) & 0x3FFF), IDESC, (first && k == 0) ? 0u : 1u);
            }
          }
          first = 0;
          tc_commit(&b_empty[b_slot]);
        }
        tc_commit(&a_empty[a_slot]);
      }
      tc_commit(acc_full);
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int img = p.Hp * p.Wp;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const long long pos = tile_start + mt * 128 + quarter * 32 + lane;
      const int n = (int)(pos / img);
      const int r = (int)(pos - (long long)n * img);
      const int h = r / p.Wp;
      const int w = r - h * p.Wp;
      const bool valid = pos < p.n_positions && h < p.H && w < p.W;
      long long out_pix = 0, res_pix = 0;
      int out_plane_ofs = 0;
      if (p.out_mode == OUT_PLAIN) {
        out_pix = p.out_base + ((long long)n * p.out_Hp + h) * p.out_Wp + w;
      } else if (p.out_mode == OUT_PARITY) {
        out_pix = p.out_base + ((long long)n * p.out_Hp + (h >> 1)) * p.out_Wp + (w >> 1);
        out_plane_ofs = (((h & 1) << 1) | (w & 1)) * (p.Cout >> 3);
      } else {
        out_pix = pos;
      }
      if (p.res) res_pix = p.res_base + ((long long)n * p.res_Hp + h) * p.res_Wp + w;
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * N_TILE + c0), acc);
        tmem_ld_wait();
        if (valid) {
          const int cout0 = n_tile * N_TILE + c0;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = c0 + j8 * 8 + j;
              v[j] = fmaf(__uint_as_float(acc[j8 * 8 + j]), s_scale[c], s_bias[c]);
            }
            const int plane = (cout0 >> 3) + j8;
            if (p.res) {
              const uint4 rr = *reinterpret_cast<const uint4*>(p.res + ((long long)plane * p.res_plane + res_pix) * 8);
              v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
              v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.out_mode == OUT_F32_ROWS) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_pix * p.Cout + cout0 + j8 * 8);
              dst[0] = make_float4(v[0], v[1], v[2], v[3]);
              dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                   ((long long)(plane + out_plane_ofs) * p.out_plane + out_pix) * 8;
              *reinterpret_cast<uint4*>(dst) = o;
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace flope

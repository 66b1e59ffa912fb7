// flope_b200: shift-GEMM convolution on tcgen05 / TMEM fed by TMA bulk copies (sm_100a).
//
// Replaces every cuDNN/cuBLAS call site of PoseResNet.forward
// (sunflower/models/posenet.py:24-34: base.conv1, layer1..4 convs, downsample convs,
// base.fc) - SURVEY.md section 2c, K2/K3/K5.
//
// A tile is TM = MT*128 consecutive pixel positions x N_TILE output channels.  In the
// blocked-pixel layout (common.cuh) a filter tap is a constant position shift, so the A
// operand of tap (dy,dx) is the *same* shared-memory halo tile read through a UMMA
// descriptor whose start address is moved by shift*16 bytes: im2col costs nothing and the
// halo tile is fetched once per 64-channel group instead of once per tap.
//
// Persistent, warp-specialised, one CTA per SM, tiles assigned round-robin:
//   warp 0   : TMA producer (cp.async.bulk -> mbarrier complete_tx), A-halo ring + weight-tile ring
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one elected lane; the loop itself is warp-uniform)
//   warps 2-17: epilogue (tcgen05.ld -> +bias (BN scale is folded into the weights), residual, ReLU -> store);
//              16 warps so that each SM sub-partition has four of them to hide TMEM/LDG/STG latency
// The accumulator is double-buffered in TMEM (2 x MT*N_TILE columns), so the epilogue of tile i
// overlaps the loads and MMAs of tile i+1.
#pragma once
#include "common.cuh"

namespace flope {

constexpr int kMaxGroups = 32;
constexpr int kMaxTaps = 16;
constexpr int kEpiWarps = 16;                  // 4 per TMEM lane quarter
constexpr int kConvThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxASlots = 8;
constexpr int kMaxBSlots = 16;

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
  int n_positions;             // N*Hp*Wp
  int Hp, Wp, H, W;
  int n_m_tiles, n_n_tiles;
  // ---- epilogue ----
  const float* bias;           // [Cout] folded BN bias / fc bias (the BN scale is folded into the packed weights)
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
  // ---- fused 3x3/s2 max-pool (stem only, POOL kernels): one tile = the three conv rows 2i-1..2i+1 of one
  //      crop, pooled into row i of the output tensor (out_* then describe the pooled grid) ----
  int pool_rows;               // pooled rows per crop (H/2)
  int pool_cols;               // pooled columns (W/2)
  int b_resident;              // all weight tiles fit the ring and n_n_tiles == 1: load them once per CTA
  // ---- second A source: K groups >= first_group2 are read from in2 (the projection shortcut of a
  //      ResNet block folded into its conv2 as extra K: same position space, shift 0) ----
  const __nv_bfloat16* in2;
  long long in2_plane;
  int in2_base;
  int first_group2;            // == n_groups when there is no second source
};

__host__ __device__ constexpr int pow2_at_least(int v) { int r = 32; while (r < v) r <<= 1; return r; }

// First position of M-tile `m_tile`: consecutive 128*MT-position ranges, or (POOL) the conv rows 2i-1..2i+1.
template <bool POOL>
__device__ __forceinline__ int conv_tile_start(const ConvParams& p, int m_tile, int TM) {
  if (!POOL) return m_tile * TM;
  const int n = m_tile / p.pool_rows;
  const int i = m_tile - n * p.pool_rows;
  return n * (p.Hp * p.Wp) + (2 * i - 1) * p.Wp;
}

// PAIR: the kernel is launched as clusters of two CTAs (one TPC).  A pair tile is 2*TM positions x N_TILE
// channels: CTA r owns positions [r*TM, (r+1)*TM) of it (its own A halo tiles, accumulators and epilogue) and
// loads rows [r*N_TILE/2, (r+1)*N_TILE/2) of every weight tile; the leader CTA (rank 0) issues
// tcgen05.mma.cta_group::2 (M = 256) for both.  Per SM and per MMA the shared-memory port then serves
// 128 A rows + N_TILE/2 B rows instead of 128 + N_TILE, which is what bounds the single-CTA kernel.
//   full barriers : each CTA's TMA completes on its own barrier; rank 1's warp 1 relays every completed phase to
//                   the leader's barrier (count 2 there: local producer + relay)
//   empty barriers: tcgen05.commit multicast arrives in both CTAs
//   accumulator   : acc_full by multicast commit; acc_empty lives in the leader (both CTAs' epilogue warps arrive)
template <int N_TILE, int MT, int KP, bool POOL, bool PAIR>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  static_assert(!(POOL && PAIR), "the fused-pool epilogue is single-CTA only");
  constexpr int TM = MT * 128;
  constexpr int NB_ROWS = PAIR ? N_TILE / 2 : N_TILE;            // weight rows this CTA holds per tile
  constexpr int ACC_COLS = pow2_at_least(N_TILE * MT);          // column stride of one accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "two accumulator stages must fit the 512 TMEM columns");
  constexpr uint32_t IDESC = umma_idesc_bf16(PAIR ? 256 : 128, N_TILE);
  constexpr int NCHUNK = N_TILE / 32;
  constexpr int KC8 = 2 * KP;                                   // planes per K group

  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_empty = a_full + kMaxASlots;
  uint64_t* b_full = a_empty + kMaxASlots;
  uint64_t* b_empty = b_full + kMaxBSlots;
  uint64_t* acc_full = b_empty + kMaxBSlots;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem_raw + 512);
  const int halo = p.halo_before + p.halo_after;
  const uint32_t a_plane_bytes = (uint32_t)(TM + halo) * 16u;
  const uint32_t a_slot_bytes = a_plane_bytes * KC8;
  constexpr uint32_t b_tile_bytes = (uint32_t)KC8 * NB_ROWS * 16u;
  uint8_t* a_ring = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 512 + (size_t)p.Cout * sizeof(float) + 127) & ~uintptr_t(127));
  uint8_t* b_ring = a_ring + (size_t)p.n_a_slots * a_slot_bytes;
  // POOL: two staging buffers [N_TILE/8 planes][TM positions][8] bf16 for the post-ReLU conv rows
  constexpr uint32_t stage_bytes = POOL ? (uint32_t)(N_TILE / 8) * TM * 16u : 0u;
  uint8_t* pool_stage = b_ring + (size_t)p.n_b_slots * b_tile_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_m_tiles * p.n_n_tiles;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs)
  const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // tiles are dealt to pairs (or CTAs) round-robin
  const int tile_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int rank_ofs = PAIR ? (int)rank * TM : 0;                // this CTA's first position inside a pair tile
  constexpr int TILE_POS = PAIR ? 2 * TM : TM;                   // positions per (pair) tile

  if (threadIdx.x == 0) {
    const uint32_t full_count = (PAIR && rank == 0) ? 2u : 1u;   // leader: own producer + the peer's relay
    for (int i = 0; i < p.n_a_slots; ++i) { mbar_init(&a_full[i], full_count); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.n_b_slots; ++i) { mbar_init(&b_full[i], full_count); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], POOL ? kEpiWarps / 2 : (PAIR ? 2 * kEpiWarps : kEpiWarps));
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc2(tmem_ptr, TMEM_COLS); tmem_relinquish2(); }
    else { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  }
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // everything above touched only static data (bias) and on-chip state; the activations of the previous layer
  // are complete and visible after this point, and the next kernel may begin its own prologue
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ===================== TMA producer: whole warp runs the loop, one lane issues =====================
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_ring_addr = smem_u32(a_ring);
    const uint32_t b_ring_addr = smem_u32(b_ring);
    int a_slot = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
      const int n_tile = tile % p.n_n_tiles;
      const int tile_start = conv_tile_start<POOL>(p, tile / p.n_n_tiles, TILE_POS) + rank_ofs;
      // PAIR: weights are packed [n_tile][tile][rank][k8][N_TILE/2][8], so each CTA's half is one contiguous copy
      const __nv_bfloat16* wtile = p.wgt + ((size_t)n_tile * p.taps_total * (PAIR ? 2 : 1) + rank) * (b_tile_bytes / 2);
      for (int g = 0; g < p.n_groups; ++g) {
        mbar_wait(&a_empty[a_slot], a_phase ^ 1);
        mbar_expect_tx_if(leader, &a_full[a_slot], a_slot_bytes);
        const uint32_t a_dst = a_ring_addr + a_slot * a_slot_bytes;
        const bool second = g >= p.first_group2;
        const long long a_plane = second ? p.in2_plane : p.in_plane;
        const __nv_bfloat16* src = (second ? p.in2 : p.in) +
            ((long long)p.group_plane[g] * a_plane + (second ? p.in2_base : p.in_base) + tile_start - p.halo_before) * 8;
#pragma unroll
        for (int j = 0; j < KC8; ++j)
          bulk_g2s_if(leader, a_dst + j * a_plane_bytes, src + (long long)j * a_plane * 8, a_plane_bytes, &a_full[a_slot]);
        if (++a_slot == p.n_a_slots) { a_slot = 0; a_phase ^= 1; }
        const int ntaps = p.group_ntaps[g];
        if (!p.b_resident || tile == first_tile) {
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(&b_empty[b_slot], b_phase ^ 1);
            mbar_expect_tx_if(leader, &b_full[b_slot], b_tile_bytes);
            bulk_g2s_if(leader, b_ring_addr + b_slot * b_tile_bytes, wtile, b_tile_bytes, &b_full[b_slot]);
            wtile += (PAIR ? 2 : 1) * (b_tile_bytes / 2);
            if (++b_slot == p.n_b_slots) { b_slot = 0; b_phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && PAIR && rank != 0) {
    // ===================== relay (non-leader CTA of a pair): forward every completed full-barrier phase of this
    // CTA to the leader's barrier of the same slot, in the order the leader's MMA loop waits for them =====================
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_full_remote = mapa_u32(smem_u32(a_full), 0);
    const uint32_t b_full_remote = mapa_u32(smem_u32(b_full), 0);
    int a_slot = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    uint32_t it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
      for (int g = 0; g < p.n_groups; ++g) {
        mbar_wait(&a_full[a_slot], a_phase);
        mbar_arrive_remote_if(leader, a_full_remote + a_slot * 8);
        const int ntaps = p.group_ntaps[g];
        for (int t = 0; t < ntaps; ++t) {
          if (!p.b_resident || it == 0) {
            mbar_wait(&b_full[b_slot], b_phase);
            mbar_arrive_remote_if(leader, b_full_remote + b_slot * 8);
          }
          if (++b_slot == p.n_b_slots) { b_slot = 0; b_phase ^= (p.b_resident ? 0u : 1u); }
        }
        if (++a_slot == p.n_a_slots) { a_slot = 0; a_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: whole warp runs the loop, one lane issues =====================
    // Descriptors (K-major, no swizzle): A planes are a_plane_bytes apart in K, 8-pixel groups 128 B apart;
    // weight tiles are [k8][N_TILE][8], so K-adjacent core matrices are N_TILE*16 B apart.  Only the
    // 14-bit start-address field changes between MMAs, so each descriptor costs one integer add.
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
    const uint32_t a_lo0 = ((a_plane_bytes >> 4) << 16) + (smem_u32(a_ring) >> 4) + (uint32_t)p.halo_before;
    const uint32_t b_lo0 = (((uint32_t)NB_ROWS * 16u >> 4) << 16) + (smem_u32(b_ring) >> 4);
    const uint32_t a_kstep = 2u * (a_plane_bytes >> 4);                      // two planes per K=16 MMA
    constexpr uint32_t b_kstep = 2u * NB_ROWS;
    const uint32_t a_slot_units = a_slot_bytes >> 4;
    constexpr uint32_t b_tile_units = b_tile_bytes >> 4;
    int a_slot = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    uint32_t it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
      const uint32_t stage = it & 1;
      if (PAIR) mbar_wait_cluster(&acc_empty[stage], ((it >> 1) & 1) ^ 1);
      else mbar_wait(&acc_empty[stage], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + stage * ACC_COLS;
      uint32_t accumulate = 0;
      for (int g = 0; g < p.n_groups; ++g) {
        if (PAIR) mbar_wait_cluster(&a_full[a_slot], a_phase);
        else mbar_wait(&a_full[a_slot], a_phase);
        const uint32_t a_grp = a_lo0 + a_slot * a_slot_units;
        const int tofs = p.group_tapofs[g];
        const int ntaps = p.group_ntaps[g];
        for (int t = 0; t < ntaps; ++t) {
          if (!p.b_resident || it == 0) {
            if (PAIR) mbar_wait_cluster(&b_full[b_slot], b_phase);
            else mbar_wait(&b_full[b_slot], b_phase);
          }
          tc_fence_after();
          const uint32_t a_tap = a_grp + (uint32_t)p.tap_shift[tofs + t];    // shift in pixels == 16-byte units
          const uint32_t b_tap = b_lo0 + b_slot * b_tile_units;
#pragma unroll
          for (int k = 0; k < KP; ++k) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (PAIR)
                umma2_bf16_if(leader, acc + mt * N_TILE, a_tap + k * a_kstep + mt * 128u, desc_hi, b_tap + k * b_kstep, desc_hi,
                              IDESC, (k == 0) ? accumulate : 1u);
              else
                umma_bf16_if(leader, acc + mt * N_TILE, a_tap + k * a_kstep + mt * 128u, desc_hi, b_tap + k * b_kstep, desc_hi,
                             IDESC, (k == 0) ? accumulate : 1u);
            }
          }
          // commits free the weight slot / halo slot (in both CTAs of a pair) once these MMAs retire
          if (!p.b_resident) { if (PAIR) tc_commit2_if(leader, &b_empty[b_slot]); else tc_commit_if(leader, &b_empty[b_slot]); }
          if (t == ntaps - 1) {
            if (PAIR) tc_commit2_if(leader, &a_empty[a_slot]); else tc_commit_if(leader, &a_empty[a_slot]);
            if (g == p.n_groups - 1) { if (PAIR) tc_commit2_if(leader, &acc_full[stage]); else tc_commit_if(leader, &acc_full[stage]); }
          }
          accumulate = 1;
          if (++b_slot == p.n_b_slots) { b_slot = 0; b_phase ^= (p.b_resident ? 0u : 1u); }
        }
        if (++a_slot == p.n_a_slots) { a_slot = 0; a_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: kEpiWarps warps, 4 per TMEM lane quarter =====================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    // POOL: the 16 epilogue warps form two groups of 8, one per accumulator stage, so the latency chains
    // (TMEM load -> stage -> barrier -> pool -> store) of consecutive tiles overlap.
    const int group = POOL ? (((warp - 2) >> 2) & 1) : 0;
    const int sub = POOL ? ((warp - 2) >> 3) : ((warp - 2) >> 2);   // share of the (mt, 32-column chunk) list
    constexpr int NSUB = POOL ? kEpiWarps / 8 : kEpiWarps / 4;
    const int img = p.Hp * p.Wp;
    uint32_t it = POOL ? group : 0;
    const int tile_step = POOL ? 2 * tile_stride : tile_stride;
    const uint32_t acc_empty_remote = PAIR ? mapa_u32(smem_u32(acc_empty), 0) : 0u;   // the leader's acc_empty barriers
    for (int tile = first_tile + (POOL ? group * tile_stride : 0); tile < total_tiles; tile += tile_step, it += (POOL ? 2 : 1)) {
      const uint32_t stage = it & 1;
      const int n_tile = tile % p.n_n_tiles;
      const int tile_start = conv_tile_start<POOL>(p, tile / p.n_n_tiles, TILE_POS) + rank_ofs;
      const int cout_base = n_tile * N_TILE;
      uint8_t* stage_buf = pool_stage + (it & 1) * stage_bytes;
      const uint32_t acc = tmem_base + stage * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      bool waited = false;
#pragma unroll 1
      for (int c = sub; c < MT * NCHUNK; c += NSUB) {
        const int mt = c / NCHUNK;
        const int c0 = (c - mt * NCHUNK) * 32;
        const int pos = tile_start + mt * 128 + quarter * 32 + lane;
        const int n = pos / img;
        const int r = pos - n * img;
        const int h = r / p.Wp;
        const int w = r - h * p.Wp;
        const bool valid = pos >= 0 && pos < p.n_positions && h < p.H && w < p.W;
        const int plane0 = (cout_base + c0) >> 3;
        uint4 res[4];
        if (p.res != nullptr && valid) {     // issued before the accumulator wait: latency hidden behind the MMAs
          const __nv_bfloat16* rp = p.res + ((long long)plane0 * p.res_plane + p.res_base + ((long long)n * p.res_Hp + h) * p.res_Wp + w) * 8;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) res[j8] = __ldg(reinterpret_cast<const uint4*>(rp + (long long)j8 * p.res_plane * 8));
        }
        if (!waited) {
          mbar_wait(&acc_full[stage], (it >> 1) & 1);
          tc_fence_after();
          waited = true;
        }
        uint32_t v32[32];
        tmem_ld32(acc + (uint32_t)(mt * N_TILE + c0), v32);
        tmem_ld_wait();
        if (c + NSUB >= MT * NCHUNK) {
          // this warp's last TMEM read of the stage: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (PAIR) mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
          else if (lane == 0) mbar_arrive(&acc_empty[stage]);
        }
        if (POOL) {
          // stage relu(acc + bias) as bf16 (zeros at padded positions: the pool's padding) for the pooling pass
          const int lp = mt * 128 + quarter * 32 + lane;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (valid) {
              const float4 ba = *reinterpret_cast<const float4*>(s_bias + cout_base + c0 + j8 * 8);
              const float4 bb = *reinterpret_cast<const float4*>(s_bias + cout_base + c0 + j8 * 8 + 4);
              o.x = pack_bf16x2_relu(__uint_as_float(v32[j8 * 8 + 0]) + ba.x, __uint_as_float(v32[j8 * 8 + 1]) + ba.y);
              o.y = pack_bf16x2_relu(__uint_as_float(v32[j8 * 8 + 2]) + ba.z, __uint_as_float(v32[j8 * 8 + 3]) + ba.w);
              o.z = pack_bf16x2_relu(__uint_as_float(v32[j8 * 8 + 4]) + bb.x, __uint_as_float(v32[j8 * 8 + 5]) + bb.y);
              o.w = pack_bf16x2_relu(__uint_as_float(v32[j8 * 8 + 6]) + bb.z, __uint_as_float(v32[j8 * 8 + 7]) + bb.w);
            }
            *reinterpret_cast<uint4*>(stage_buf + ((size_t)((c0 >> 3) + j8) * TM + lp) * 16) = o;
          }
        } else if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(s_bias + cout_base + c0 + j);
            v[j] = __uint_as_float(v32[j]) + b.x; v[j + 1] = __uint_as_float(v32[j + 1]) + b.y;
            v[j + 2] = __uint_as_float(v32[j + 2]) + b.z; v[j + 3] = __uint_as_float(v32[j + 3]) + b.w;
          }
          if (p.res != nullptr) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const uint4 rr = res[j8];
              v[j8 * 8 + 0] += bf16_lo(rr.x); v[j8 * 8 + 1] += bf16_hi(rr.x); v[j8 * 8 + 2] += bf16_lo(rr.y); v[j8 * 8 + 3] += bf16_hi(rr.y);
              v[j8 * 8 + 4] += bf16_lo(rr.z); v[j8 * 8 + 5] += bf16_hi(rr.z); v[j8 * 8 + 6] += bf16_lo(rr.w); v[j8 * 8 + 7] += bf16_hi(rr.w);
            }
          }
          if (p.out_mode == OUT_F32_ROWS) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)pos * p.Cout + cout_base + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (p.relu) dst[j >> 2] = make_float4(fmaxf(v[j], 0.f), fmaxf(v[j + 1], 0.f), fmaxf(v[j + 2], 0.f), fmaxf(v[j + 3], 0.f));
              else dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            long long out_pix;
            int plane = plane0;
            if (p.out_mode == OUT_PLAIN) {
              out_pix = p.out_base + ((long long)n * p.out_Hp + h) * p.out_Wp + w;
            } else {
              out_pix = p.out_base + ((long long)n * p.out_Hp + (h >> 1)) * p.out_Wp + (w >> 1);
              plane += (((h & 1) << 1) | (w & 1)) * (p.Cout >> 3);
            }
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)plane * p.out_plane + out_pix) * 8;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              uint4 o;
              if (p.relu) {
                o.x = pack_bf16x2_relu(v[j8 * 8 + 0], v[j8 * 8 + 1]); o.y = pack_bf16x2_relu(v[j8 * 8 + 2], v[j8 * 8 + 3]);
                o.z = pack_bf16x2_relu(v[j8 * 8 + 4], v[j8 * 8 + 5]); o.w = pack_bf16x2_relu(v[j8 * 8 + 6], v[j8 * 8 + 7]);
              } else {
                o.x = pack_bf16x2(v[j8 * 8 + 0], v[j8 * 8 + 1]); o.y = pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
                o.z = pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]); o.w = pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
              }
              *reinterpret_cast<uint4*>(dst + (long long)j8 * p.out_plane * 8) = o;
            }
          }
        }
      }
      if (POOL) {
        // all epilogue warps have staged their chunks -> 3x3/s2 max over the three staged conv rows
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kEpiWarps * 16) : "memory");
        const int m_tile = tile / p.n_n_tiles;
        const int n = m_tile / p.pool_rows;
        const int i = m_tile - n * p.pool_rows;
        const int items = p.pool_cols * (N_TILE / 8);
        const int gtid = (warp & 3) * 32 + lane + (sub << 7);          // thread index inside the group (0..255)
        for (int t = gtid; t < items; t += kEpiWarps * 16) {
          const int plane = t / p.pool_cols;
          const int j = t - plane * p.pool_cols;
          const uint8_t* src = stage_buf + (size_t)plane * TM * 16;
          uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int lr = 0; lr < 3; ++lr) {
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int lp = lr * p.Wp + 2 * j + dx;
              if (lp >= 0) {
                const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)lp * 16);
                m.x = bf16x2_max_u32(m.x, v.x); m.y = bf16x2_max_u32(m.y, v.y);
                m.z = bf16x2_max_u32(m.z, v.z); m.w = bf16x2_max_u32(m.w, v.w);
              }
            }
          }
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                               ((long long)((cout_base >> 3) + plane) * p.out_plane + p.out_base +
                                ((long long)n * p.out_Hp + i) * p.out_Wp + j) * 8;
          *reinterpret_cast<uint4*>(dst) = m;
        }
        // the group's next tile reuses this staging buffer: everyone must be done reading it
        asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kEpiWarps * 16) : "memory");
      }
      if (!waited) {
        // a warp with no chunk in this configuration still takes part in the accumulator hand-back
        mbar_wait(&acc_full[stage], (it >> 1) & 1);
        __syncwarp();
        if (PAIR) mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
        else if (lane == 0) mbar_arrive(&acc_empty[stage]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();            // the peer may still read this CTA's smem / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace flope

// flope_b200: layer1 .. layer4 of PoseResNet (sixteen convs, sunflower/models/posenet.py:28-31) as ONE persistent
// launch (sm_100a).
//
// conv_igemm.cuh runs one launch per ResNet stage (a ConvChain).  tools/timeline.py shows what every launch boundary
// still costs at 256 crops: the SMs free up over 8-17 us as the predecessor's last tiles finish (entry skew) and then
// sit in griddepcontrol.wait until ALL of them have (2-10 us), each CTA spends ~2 us in its prologue and 2-3 us more
// until its first operands land, and the last tile's epilogue (4-7 us) runs with the tensor pipe idle: 15-35 us per
// boundary that no other stream can fill, because a CTA owns the whole SM's shared memory.
//
// This kernel walks the four stages with the same warp roles, rings, barriers and TMEM accumulators, so none of
// that happens between stages:
//   * work items of all stages are numbered stage-major, layer-major and dealt round-robin to the CTA pairs; every role
//     (producer / MMA issuer or relay / epilogue warps) carries its own item counter across the stages;
//   * the tile shape is a compile-time property (64x4, 128x2, 256x1 = N_TILE x MT sub-tiles; 64x1 latency tiles for
//     stages that are too few items at a small max_batch), so each role calls one template instance per stage; all
//     shapes use the same two 256-column TMEM accumulator stages;
//   * the shared-memory rings are re-cut per stage (slot sizes differ).  The producer drains the old rings (waits for
//     every slot's release) before its first copy into the new geometry - a bubble of one tile's MMAs per pair, at
//     a moment that differs from pair to pair; barrier phases are tracked per slot (bit masks), because slot counts
//     change between stages;
//   * dependencies inside a stage are the ConvChain tile flags; the first conv of stages 2-4 (3x3 stride 2 over the
//     previous stage's parity-split output) waits for the previous stage's tiles that cover its halo rows: half-res
//     rows h0..h1 of crop range n0..n1 <- full-res rows 2*h0 .. 2*h1+1, a contiguous tile range.  The 1x1 projection
//     (folded into conv2 as extra K) reads positions its own conv1 neighbours already waited for;
//   * the epilogue warps re-stage the stage's folded-BN biases between two named barriers of their own; each stage's
//     rings start right behind its own bias block;
//   * latency tiles (TrunkStage<64, 1, true>: small max_batch, e.g. a streamed frame with 8 flowers) keep three weight
//     tiles per ring slot - one barrier round per three taps for producer, relay and issuer - and may split K: a
//     stage whose layers are at most a third of the CTA pairs deals every tile as k_splits items over disjoint K-group
//     ranges; the first k_splits-1 store raw fp32 accumulators and count themselves in split_flags, the last adds them
//     in index order and runs the epilogue (ConvChain::k_splits, engine.cu plan_splits).
// Static dealing needs every CTA resident (grid <= SM count, one such launch on the device at a time) - the same
// condition as the per-stage chains.  engine.cu enforces it: forwards that contain these launches pass a per-device
// gate (an event chain under a host mutex), so two engines on two streams run their backbones one after the other;
// engines meant to overlap (pipeline.EnginePool) use per-layer launches, which wait for nothing.
#pragma once
#include "conv_igemm.cuh"

namespace flope {

constexpr int kTrunkStages = 4;
struct TrunkParams {
  ConvChain st[kTrunkStages];          // st[s].flags: that stage's tile counters
  int tile_pos[kTrunkStages];          // positions per pair tile of each stage
  unsigned long long* stamps;          // optional phase stamps (see ConvChain::stamps)
};

struct TrunkCtx {
  uint64_t *a_full, *a_empty, *b_full, *b_empty, *acc_full, *acc_empty;
  float* s_bias;
  uint8_t* smem;                       // start of dynamic shared memory (barriers, bias block, rings)
  uint32_t tmem_base;
  uint32_t rank;
  int lane, warp;
  int first_work, work_stride;
};

// per-role running state across the stages
struct TrunkRole {
  int k = 0;                           // items this CTA pair has taken so far (all stages)
  uint32_t it = 0;                     // accumulator-stage counter (MMA / epilogue)
  uint32_t a_par = 0, b_par = 0;       // per-slot parity of completed uses of the role's side of the rings
};

template <int N_TILE, int MT, bool SPLIT = false>
struct TrunkStage {
  static constexpr int KP = 4, KC8 = 8;
  static constexpr int TM = MT * 128;
  static constexpr int NB_ROWS = N_TILE / 2;
  static constexpr int ACC_COLS = 256;         // accumulator stage stride: the same for every stage shape
  static constexpr uint32_t IDESC = umma_idesc_bf16(256, N_TILE);
  static constexpr int NCHUNK = N_TILE / 32;
  static constexpr int TILE_POS = 2 * TM;
  static constexpr uint32_t b_tile_bytes = (uint32_t)KC8 * NB_ROWS * 16u;
  static constexpr int TB = SPLIT ? kLatencyTapsPerSlot : 1;          // weight tiles per ring slot (one barrier round each)
  static constexpr uint32_t b_slot_bytes = (uint32_t)TB * b_tile_bytes;
  static_assert(N_TILE * MT <= ACC_COLS, "a tile's accumulators fit one 256-column stage");
  static_assert(!SPLIT || MT == 1, "split-K is for the latency tiles");

  // Work item g of a stage -> (layer, tile, split) and the K-group range / first weight tile of the split.
  struct Item { int l, w, sp, ks, gi0, gi1, tap0; };
  static __device__ __forceinline__ int items_per_layer(const ConvChain& ch) {
    return ch.L[0].n_work * (SPLIT && ch.k_splits > 1 ? ch.k_splits : 1);
  }
  static __device__ __forceinline__ Item decode(const ConvChain& ch, int g) {
    Item it;
    it.ks = SPLIT && ch.k_splits > 1 ? ch.k_splits : 1;
    const int per_layer = ch.L[0].n_work * it.ks;
    it.l = g / per_layer;
    const int r = g - it.l * per_layer;
    it.w = r / it.ks;
    it.sp = r - it.w * it.ks;
    it.gi0 = 0; it.gi1 = ch.L[it.l].n_groups; it.tap0 = 0;
    if (SPLIT && it.ks > 1) { it.gi0 = ch.split_group[it.l][it.sp]; it.gi1 = ch.split_group[it.l][it.sp + 1]; it.tap0 = ch.split_tap[it.l][it.sp]; }
    return it;
  }

  // The rings of a stage start behind ITS bias block (kMaxChain x Cout floats), exactly as conv_igemm.cuh lays a chain
  // out, so the launch needs no more shared memory than the largest stage.  A later stage's larger bias block grows
  // into the previous stage's ring area - which is dead by then: the epilogue warps write it only after their last
  // accumulator of the previous stage is complete, i.e. after every MMA that read those rings.
  static __device__ __forceinline__ uint32_t ring_addr(const ConvChain& ch, const TrunkCtx& c) {
    return (smem_u32(c.smem) + 512u + (uint32_t)(ch.n_layers * ch.L[0].Cout) * 4u + 127u) & ~127u;
  }

  // ===================== TMA producer (warp 0 of both CTAs) =====================
  // prev: the previous stage (nullptr for the first), prev_tile_pos: positions per pair tile there
  static __device__ __forceinline__ void producer(const ConvChain& ch, const ConvChain* prev, int prev_tile_pos, int item_base,
                                                  const TrunkCtx& c, TrunkRole& s) {
    const ConvParams& p = ch.L[0];
    const int halo = p.halo_before + p.halo_after;
    const uint32_t a_plane_bytes = (uint32_t)(TM + halo) * 16u;
    const uint32_t a_slot_bytes = a_plane_bytes * KC8;
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_ring_addr = ring_addr(ch, c);
    const uint32_t b_ring_addr = a_ring_addr + (uint32_t)p.n_a_slots * a_slot_bytes;
    const int total = ch.n_layers * items_per_layer(ch);
    const int img = p.Hp * p.Wp;
    int a_slot = 0, b_slot = 0;
    bool drained = prev == nullptr;
    for (;; ++s.k) {
      const int g = c.first_work + s.k * c.work_stride - item_base;
      if (g >= total) break;
      if (!drained) {
        // the new ring geometry overlaps the old slots arbitrarily: every old slot must have been released
        // (all barriers, not just the previous stage's: this pair may have had no item there; unused ones pass at once)
        for (int i = 0; i < kMaxASlots; ++i) mbar_wait(&c.a_empty[i], ((s.a_par >> i) & 1u) ^ 1u);
        for (int i = 0; i < kMaxBSlots; ++i) mbar_wait(&c.b_empty[i], ((s.b_par >> i) & 1u) ^ 1u);
        drained = true;
      }
      const Item it = decode(ch, g);
      const int l = it.l, w = it.w;
      const ConvParams& q = ch.L[l];
      const int m = w / p.n_n_tiles;
      const int tile_start = m * TILE_POS + (int)c.rank * TM;
      if (l > 0) {
        const uint32_t* fl = ch.flags + (size_t)(l - 1) * ch.n_m_tiles;
        if (m > 0) wait_tile_flag(fl + m - 1, ch.expected, ch.fail, ch.host_err);
        wait_tile_flag(fl + m, ch.expected, ch.fail, ch.host_err);
        if (m + 1 < ch.n_m_tiles) wait_tile_flag(fl + m + 1, ch.expected, ch.fail, ch.host_err);
        asm volatile("fence.proxy.async;" ::: "memory");
      } else if (prev != nullptr) {
        // stride-2 conv over the previous stage's output: half-res rows under this CTA's halo tile -> full-res rows
        const ConvParams& pp = prev->L[0];
        int a = tile_start - p.halo_before, b = tile_start + TM - 1 + p.halo_after;
        a = a < 0 ? 0 : a;
        b = b >= p.n_positions ? p.n_positions - 1 : b;
        if (a <= b) {
          const int na = a / img, ha = (a - na * img) / p.Wp;
          const int nb = b / img, hb = (b - nb * img) / p.Wp;
          const int ra = 2 * ha < pp.H ? 2 * ha : pp.H - 1;
          const int rb = 2 * hb + 1 < pp.H ? 2 * hb + 1 : pp.H - 1;
          const int lo = (na * pp.Hp + ra) * pp.Wp;
          const int hi = (nb * pp.Hp + rb) * pp.Wp + pp.Wp - 1;
          const int t_lo = lo / prev_tile_pos;
          int t_hi = hi / prev_tile_pos;
          t_hi = t_hi < prev->n_m_tiles ? t_hi : prev->n_m_tiles - 1;
          const uint32_t* fl = prev->flags + (size_t)(prev->n_layers - 1) * prev->n_m_tiles;
          for (int t = t_lo; t <= t_hi; ++t) wait_tile_flag(fl + t, prev->expected, ch.fail, ch.host_err);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
      }
      const int n_tile = w % p.n_n_tiles;
      const __nv_bfloat16* wtile = q.wgt + (((size_t)n_tile * q.taps_total + it.tap0) * 2 + c.rank) * (b_tile_bytes / 2);
      for (int gi = it.gi0; gi < it.gi1; ++gi) {
        mbar_wait(&c.a_empty[a_slot], ((s.a_par >> a_slot) & 1u) ^ 1u);
        mbar_expect_tx_if(leader, &c.a_full[a_slot], a_slot_bytes);
        const uint32_t a_dst = a_ring_addr + a_slot * a_slot_bytes;
        const bool second = gi >= q.first_group2;
        const long long a_plane = second ? q.in2_plane : q.in_plane;
        const __nv_bfloat16* src = (second ? q.in2 : q.in) +
            ((long long)q.group_plane[gi] * a_plane + (second ? q.in2_base : q.in_base) + tile_start - p.halo_before) * 8;
#pragma unroll
        for (int j = 0; j < KC8; ++j)
          bulk_g2s_if(leader, a_dst + j * a_plane_bytes, src + (long long)j * a_plane * 8, a_plane_bytes, &c.a_full[a_slot]);
        s.a_par ^= 1u << a_slot;
        if (++a_slot == p.n_a_slots) a_slot = 0;
        const int ntaps = q.group_ntaps[gi];
        for (int t = 0; t < ntaps; t += TB) {
          const int nt = ntaps - t < TB ? ntaps - t : TB;
          mbar_wait(&c.b_empty[b_slot], ((s.b_par >> b_slot) & 1u) ^ 1u);
          mbar_expect_tx_if(leader, &c.b_full[b_slot], (uint32_t)nt * b_tile_bytes);
#pragma unroll
          for (int j = 0; j < TB; ++j) {
            if (j < nt) {
              bulk_g2s_if(leader, b_ring_addr + b_slot * b_slot_bytes + j * b_tile_bytes, wtile, b_tile_bytes, &c.b_full[b_slot]);
              wtile += 2 * (b_tile_bytes / 2);
            }
          }
          s.b_par ^= 1u << b_slot;
          if (++b_slot == p.n_b_slots) b_slot = 0;
        }
      }
    }
  }

  // ===================== relay (warp 1 of the non-leader CTA) =====================
  static __device__ __forceinline__ void relay(const ConvChain& ch, int item_base, const TrunkCtx& c, TrunkRole& s) {
    const ConvParams& p = ch.L[0];
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_full_remote = mapa_u32(smem_u32(c.a_full), 0);
    const uint32_t b_full_remote = mapa_u32(smem_u32(c.b_full), 0);
    const int total = ch.n_layers * items_per_layer(ch);
    int a_slot = 0, b_slot = 0;
    for (;; ++s.k) {
      const int g = c.first_work + s.k * c.work_stride - item_base;
      if (g >= total) break;
      const Item it = decode(ch, g);
      const ConvParams& q = ch.L[it.l];
      for (int gi = it.gi0; gi < it.gi1; ++gi) {
        mbar_wait(&c.a_full[a_slot], (s.a_par >> a_slot) & 1u);
        mbar_arrive_remote_if(leader, a_full_remote + a_slot * 8);
        s.a_par ^= 1u << a_slot;
        const int ntaps = q.group_ntaps[gi];
        for (int t = 0; t < ntaps; t += TB) {
          mbar_wait(&c.b_full[b_slot], (s.b_par >> b_slot) & 1u);
          mbar_arrive_remote_if(leader, b_full_remote + b_slot * 8);
          s.b_par ^= 1u << b_slot;
          if (++b_slot == p.n_b_slots) b_slot = 0;
        }
        if (++a_slot == p.n_a_slots) a_slot = 0;
      }
    }
  }

  // ===================== MMA issuer (warp 1 of the leader CTA) =====================
  static __device__ __forceinline__ void mma(const ConvChain& ch, int item_base, const TrunkCtx& c, TrunkRole& s) {
    const ConvParams& p = ch.L[0];
    const int halo = p.halo_before + p.halo_after;
    const uint32_t a_plane_bytes = (uint32_t)(TM + halo) * 16u;
    const uint32_t a_slot_bytes = a_plane_bytes * KC8;
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_ring_addr = ring_addr(ch, c);
    const uint32_t a_lo0 = ((a_plane_bytes >> 4) << 16) + (a_ring_addr >> 4) + (uint32_t)p.halo_before;
    const uint32_t b_lo0 = (((uint32_t)NB_ROWS * 16u >> 4) << 16) + ((a_ring_addr + (uint32_t)p.n_a_slots * a_slot_bytes) >> 4);
    const uint32_t a_kstep = 2u * (a_plane_bytes >> 4);
    constexpr uint32_t b_kstep = 2u * NB_ROWS;
    const uint32_t a_slot_units = a_slot_bytes >> 4;
    constexpr uint32_t b_tile_units = b_tile_bytes >> 4;
    constexpr uint32_t b_slot_units = b_slot_bytes >> 4;
    const int total = ch.n_layers * items_per_layer(ch);
    int a_slot = 0, b_slot = 0;
    for (;; ++s.k) {
      const int g = c.first_work + s.k * c.work_stride - item_base;
      if (g >= total) break;
      const Item it = decode(ch, g);
      const ConvParams& q = ch.L[it.l];
      const uint32_t stage = s.it & 1;
      mbar_wait(&c.acc_empty[stage], ((s.it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = c.tmem_base + stage * ACC_COLS;
      uint32_t accumulate = 0;
      for (int gi = it.gi0; gi < it.gi1; ++gi) {
        mbar_wait(&c.a_full[a_slot], (s.a_par >> a_slot) & 1u);
        s.a_par ^= 1u << a_slot;
        const uint32_t a_grp = a_lo0 + a_slot * a_slot_units;
        const bool last_group = gi == it.gi1 - 1;
        const int tofs = q.group_tapofs[gi];
        const int ntaps = q.group_ntaps[gi];
        for (int t = 0; t < ntaps; t += TB) {
          const int nt = ntaps - t < TB ? ntaps - t : TB;
          mbar_wait(&c.b_full[b_slot], (s.b_par >> b_slot) & 1u);
          s.b_par ^= 1u << b_slot;
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < TB; ++j) {
            if (j < nt) {
              const uint32_t a_tap = a_grp + (uint32_t)q.tap_shift[tofs + t + j];
              const uint32_t b_tap = b_lo0 + b_slot * b_slot_units + j * b_tile_units;
#pragma unroll
              for (int k = 0; k < KP; ++k) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                  umma2_bf16_if(leader, acc + mt * N_TILE, a_tap + k * a_kstep + mt * 128u, desc_hi, b_tap + k * b_kstep, desc_hi,
                                IDESC, k == 0 ? accumulate : 1u);
              }
              accumulate = 1;
            }
          }
          tc_commit2_if(leader, &c.b_empty[b_slot]);
          if (t + nt == ntaps) {
            tc_commit2_if(leader, &c.a_empty[a_slot]);
            if (last_group) tc_commit2_if(leader, &c.acc_full[stage]);
          }
          if (++b_slot == p.n_b_slots) b_slot = 0;
        }
        if (++a_slot == p.n_a_slots) a_slot = 0;
      }
      ++s.it;
    }
  }

  // ===================== epilogue (warps 2..17 of both CTAs) =====================
  static __device__ __forceinline__ void epilogue(const ConvChain& ch, int item_base, const TrunkCtx& c, TrunkRole& s) {
    const ConvParams& p = ch.L[0];
    const int quarter = c.warp & 3;
    const int sub = (c.warp - 2) >> 2;
    constexpr int NSUB = kEpiWarps / 4;
    const int img = p.Hp * p.Wp;
    const int lane = c.lane;
    const uint32_t acc_empty_remote = mapa_u32(smem_u32(c.acc_empty), 0);
    const int total = ch.n_layers * items_per_layer(ch);
    // this stage's folded-BN biases: every epilogue warp is done with the previous stage's before they are replaced
    asm volatile("bar.sync 6, %0;" ::"n"(kEpiWarps * 32) : "memory");
    for (int i = (int)threadIdx.x - 64; i < ch.n_layers * p.Cout; i += kEpiWarps * 32) c.s_bias[i] = ch.L[i / p.Cout].bias[i % p.Cout];
    asm volatile("bar.sync 6, %0;" ::"n"(kEpiWarps * 32) : "memory");
    for (;; ++s.k) {
      const int g = c.first_work + s.k * c.work_stride - item_base;
      if (g >= total) break;
      const Item it = decode(ch, g);
      const int l = it.l, w = it.w;
      const ConvParams& q = ch.L[l];
      const float* bias_l = c.s_bias + l * p.Cout;
      const uint32_t stage = s.it & 1;
      const int n_tile = w % p.n_n_tiles;
      const int tile_start = (w / p.n_n_tiles) * TILE_POS + (int)c.rank * TM;
      const int cout_base = n_tile * N_TILE;
      const uint32_t acc = c.tmem_base + stage * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      bool waited = false;
      if (SPLIT && it.sp + 1 < it.ks) {
        // a partial: the raw accumulators go to the workspace ([column quad][row]: a warp's 32 rows are contiguous)
        float4* ws = ch.split_ws + ((((size_t)w * (it.ks - 1) + it.sp) * 2 + c.rank) * (N_TILE / 4)) * 128 + quarter * 32 + lane;
#pragma unroll 1
        for (int cc = sub; cc < NCHUNK; cc += NSUB) {
          if (!waited) {
            mbar_wait(&c.acc_full[stage], (s.it >> 1) & 1);
            tc_fence_after();
            waited = true;
          }
          uint32_t v32[32];
          tmem_ld32(acc + (uint32_t)(cc * 32), v32);
          tmem_ld_wait();
          if (cc + NSUB >= NCHUNK) {
            tc_fence_before();
            __syncwarp();
            mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            ws[(size_t)(cc * 8 + j4) * 128] = make_float4(__uint_as_float(v32[4 * j4]), __uint_as_float(v32[4 * j4 + 1]),
                                                         __uint_as_float(v32[4 * j4 + 2]), __uint_as_float(v32[4 * j4 + 3]));
        }
        if (!waited) {
          mbar_wait(&c.acc_full[stage], (s.it >> 1) & 1);
          tc_fence_before();
          __syncwarp();
          mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
        }
        asm volatile("bar.sync 5, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (threadIdx.x == 64) {
          __threadfence();
          atomicAdd(ch.split_flags + (size_t)l * p.n_work + w, 1u);
        }
        ++s.it;
        continue;
      }
      bool partials_ready = false;
      if (q.res_layer >= 0) {
        wait_tile_flag(ch.flags + (size_t)q.res_layer * ch.n_m_tiles + w / p.n_n_tiles, ch.expected, ch.fail, ch.host_err);
        __syncwarp();
      }
#pragma unroll 1
      for (int cc = sub; cc < MT * NCHUNK; cc += NSUB) {
        const int mt = cc / NCHUNK;
        const int c0 = (cc - mt * NCHUNK) * 32;
        const int pos = tile_start + mt * 128 + quarter * 32 + lane;
        const int n = pos / img;
        const int r = pos - n * img;
        const int h = r / p.Wp;
        const int ww = r - h * p.Wp;
        const bool valid = pos >= 0 && pos < p.n_positions && h < p.H && ww < p.W;
        const int plane0 = (cout_base + c0) >> 3;
        uint4 res[4];
        if (q.res != nullptr && valid) {
          const __nv_bfloat16* rp = q.res + ((long long)plane0 * q.res_plane + q.res_base + ((long long)n * q.res_Hp + h) * q.res_Wp + ww) * 8;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) res[j8] = __ldcg(reinterpret_cast<const uint4*>(rp + (long long)j8 * q.res_plane * 8));
        }
        if (!waited) {
          mbar_wait(&c.acc_full[stage], (s.it >> 1) & 1);
          tc_fence_after();
          waited = true;
        }
        uint32_t v32[32];
        tmem_ld32(acc + (uint32_t)(mt * N_TILE + c0), v32);
        tmem_ld_wait();
        if (cc + NSUB >= MT * NCHUNK) {
          tc_fence_before();
          __syncwarp();
          mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
        }
        if (SPLIT && it.ks > 1) {
          if (!partials_ready) {
            wait_tile_flag(ch.split_flags + (size_t)l * p.n_work + w, 2u * (uint32_t)(it.ks - 1), ch.fail, ch.host_err);
            __syncwarp();
            partials_ready = true;
          }
          if (valid) {
            for (int s2 = 0; s2 < it.ks - 1; ++s2) {       // index order: the sum does not depend on who finished first
              const float4* ws = ch.split_ws + ((((size_t)w * (it.ks - 1) + s2) * 2 + c.rank) * (N_TILE / 4) + (c0 >> 2)) * 128 + quarter * 32 + lane;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 t = __ldcg(ws + (size_t)j4 * 128);
                v32[4 * j4] = __float_as_uint(__uint_as_float(v32[4 * j4]) + t.x);
                v32[4 * j4 + 1] = __float_as_uint(__uint_as_float(v32[4 * j4 + 1]) + t.y);
                v32[4 * j4 + 2] = __float_as_uint(__uint_as_float(v32[4 * j4 + 2]) + t.z);
                v32[4 * j4 + 3] = __float_as_uint(__uint_as_float(v32[4 * j4 + 3]) + t.w);
              }
            }
          }
        }
        if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias_l + cout_base + c0 + j);
            v[j] = __uint_as_float(v32[j]) + b.x; v[j + 1] = __uint_as_float(v32[j + 1]) + b.y;
            v[j + 2] = __uint_as_float(v32[j + 2]) + b.z; v[j + 3] = __uint_as_float(v32[j + 3]) + b.w;
          }
          if (q.res != nullptr) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const uint4 rr = res[j8];
              v[j8 * 8 + 0] += bf16_lo(rr.x); v[j8 * 8 + 1] += bf16_hi(rr.x); v[j8 * 8 + 2] += bf16_lo(rr.y); v[j8 * 8 + 3] += bf16_hi(rr.y);
              v[j8 * 8 + 4] += bf16_lo(rr.z); v[j8 * 8 + 5] += bf16_hi(rr.z); v[j8 * 8 + 6] += bf16_lo(rr.w); v[j8 * 8 + 7] += bf16_hi(rr.w);
            }
          }
          long long out_pix;
          int plane = plane0;
          if (q.out_mode == OUT_PLAIN) {
            out_pix = q.out_base + ((long long)n * q.out_Hp + h) * q.out_Wp + ww;
          } else {
            out_pix = q.out_base + ((long long)n * q.out_Hp + (h >> 1)) * q.out_Wp + (ww >> 1);
            plane += (((h & 1) << 1) | (ww & 1)) * (p.Cout >> 3);
          }
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(q.out) + ((long long)plane * q.out_plane + out_pix) * 8;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            uint4 o;
            if (q.relu) {
              o.x = pack_bf16x2_relu(v[j8 * 8 + 0], v[j8 * 8 + 1]); o.y = pack_bf16x2_relu(v[j8 * 8 + 2], v[j8 * 8 + 3]);
              o.z = pack_bf16x2_relu(v[j8 * 8 + 4], v[j8 * 8 + 5]); o.w = pack_bf16x2_relu(v[j8 * 8 + 6], v[j8 * 8 + 7]);
            } else {
              o.x = pack_bf16x2(v[j8 * 8 + 0], v[j8 * 8 + 1]); o.y = pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
              o.z = pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]); o.w = pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
            }
            *reinterpret_cast<uint4*>(dst + (long long)j8 * q.out_plane * 8) = o;
          }
        }
      }
      if (!waited) {
        mbar_wait(&c.acc_full[stage], (s.it >> 1) & 1);
        tc_fence_before();
        __syncwarp();
        mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
      }
      // publish this CTA's part of the tile (conv_igemm.cuh: barrier -> fence -> counter)
      asm volatile("bar.sync 5, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        atomicAdd(ch.flags + (size_t)l * ch.n_m_tiles + w / p.n_n_tiles, 1u);
      }
      ++s.it;
    }
  }
};

// Stage shapes of ResNet-18 with the CTA-pair plan (engine.cu plan_conv): 64 channels -> 64x4, 128 -> 128x2, >= 256 -> 256x1;
// a stage whose layers would be too few items at the engine's max_batch runs on the latency tiles 64x1 instead
// (bit s of SMALL; small batches switch the late stages first: 0, 8, 12, 14, 15 are the masks that occur).
template <int S, bool SMALL> struct TrunkShape { using type = TrunkStage<64, 1, true>; };
template <> struct TrunkShape<0, false> { using type = TrunkStage<64, 4>; };
template <> struct TrunkShape<1, false> { using type = TrunkStage<128, 2>; };
template <> struct TrunkShape<2, false> { using type = TrunkStage<256, 1>; };
template <> struct TrunkShape<3, false> { using type = TrunkStage<256, 1>; };

template <int SMALL>
__global__ void __launch_bounds__(kConvThreads, 1) trunk_chain_kernel(const __grid_constant__ TrunkParams tp) {
  using St0 = typename TrunkShape<0, (SMALL & 1) != 0>::type;
  using St1 = typename TrunkShape<1, (SMALL & 2) != 0>::type;
  using St2 = typename TrunkShape<2, (SMALL & 4) != 0>::type;
  using St3 = typename TrunkShape<3, (SMALL & 8) != 0>::type;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  TrunkCtx c;
  c.a_full = reinterpret_cast<uint64_t*>(smem_raw);
  c.a_empty = c.a_full + kMaxASlots;
  c.b_full = c.a_empty + kMaxASlots;
  c.b_empty = c.b_full + kMaxBSlots;
  c.acc_full = c.b_empty + kMaxBSlots;
  c.acc_empty = c.acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(c.acc_empty + 2);
  c.s_bias = reinterpret_cast<float*>(smem_raw + 512);
  c.smem = smem_raw;
  c.warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  c.lane = threadIdx.x & 31;
  c.rank = cluster_ctarank();
  c.first_work = (int)(blockIdx.x >> 1);
  c.work_stride = (int)(gridDim.x >> 1);

  unsigned long long* stamps = tp.stamps ? tp.stamps + (size_t)blockIdx.x * kStampWords : nullptr;
  auto stamp = [&](int i) {
    if (!stamps) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    stamps[i] = t;
  };
  if (threadIdx.x == 0) stamp(0);
  if (threadIdx.x == 0) {
    const uint32_t full_count = c.rank == 0 ? 2u : 1u;
    for (int i = 0; i < kMaxASlots; ++i) { mbar_init(&c.a_full[i], full_count); mbar_init(&c.a_empty[i], 1); }
    for (int i = 0; i < kMaxBSlots; ++i) { mbar_init(&c.b_full[i], full_count); mbar_init(&c.b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&c.acc_full[i], 1); mbar_init(&c.acc_empty[i], 2 * kEpiWarps); }
    mbar_fence_init();
  }
  if (c.warp == 1) { tmem_alloc2(tmem_ptr, 512); tmem_relinquish2(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  c.tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) stamp(1);
  griddep_wait();
  griddep_launch();
  if (threadIdx.x == 0) stamp(2);

  int base[kTrunkStages + 1];
  base[0] = 0;
  for (int s = 0; s < kTrunkStages; ++s)
    base[s + 1] = base[s] + tp.st[s].n_layers * tp.st[s].L[0].n_work * (tp.st[s].k_splits > 1 ? tp.st[s].k_splits : 1);
  TrunkRole role;
  if (c.warp == 0) {
    St0::producer(tp.st[0], nullptr, 0, base[0], c, role);
    St1::producer(tp.st[1], &tp.st[0], tp.tile_pos[0], base[1], c, role);
    St2::producer(tp.st[2], &tp.st[1], tp.tile_pos[1], base[2], c, role);
    St3::producer(tp.st[3], &tp.st[2], tp.tile_pos[2], base[3], c, role);
  } else if (c.warp == 1 && c.rank != 0) {
    St0::relay(tp.st[0], base[0], c, role);
    St1::relay(tp.st[1], base[1], c, role);
    St2::relay(tp.st[2], base[2], c, role);
    St3::relay(tp.st[3], base[3], c, role);
  } else if (c.warp == 1) {
    St0::mma(tp.st[0], base[0], c, role);
    St1::mma(tp.st[1], base[1], c, role);
    St2::mma(tp.st[2], base[2], c, role);
    St3::mma(tp.st[3], base[3], c, role);
    if (c.lane == 0) stamp(4);
  } else {
    St0::epilogue(tp.st[0], base[0], c, role);
    St1::epilogue(tp.st[1], base[1], c, role);
    St2::epilogue(tp.st[2], base[2], c, role);
    St3::epilogue(tp.st[3], base[3], c, role);
    if (threadIdx.x == 64) stamp(6);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(7);
  cluster_sync_all();
  if (c.warp == 1) {
    tc_fence_after();
    tmem_dealloc2(c.tmem_base, 512);
  }
}

}  // namespace flope

// flope_b200: shift-GEMM convolution on tcgen05 / TMEM fed by TMA bulk copies (sm_100a).
//
// Replaces every cuDNN/cuBLAS call site of PoseResNet.forward
// (sunflower/models/posenet.py:24-34: base.conv1, base.maxpool, layer1..4 convs, downsample convs,
// base.fc) - SURVEY.md section 2c, K2/K3/K5.
//
// A tile is TM = MT*128 pixel positions x N_TILE output channels.  In the blocked-pixel layout
// (common.cuh) a filter tap is a constant position shift, so the A operand of tap (dy,dx) is the
// *same* shared-memory halo tile read through a UMMA descriptor whose start address is moved by
// shift*16 bytes: im2col costs nothing and the halo tile is fetched once per 64-channel group
// instead of once per tap.
//
// Persistent, warp-specialised, one CTA per SM, work dealt round-robin:
//   warp 0   : TMA producer (cp.async.bulk -> mbarrier complete_tx), A-halo ring + weight-tile ring
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one elected lane; the loop itself is warp-uniform)
//   warps 2-17: epilogue (tcgen05.ld -> +bias (BN scale is folded into the weights), residual, ReLU -> store);
//              16 warps so that each SM sub-partition has four of them to hide TMEM/LDG/STG latency
// The accumulator is double-buffered in TMEM (2 x MT*N_TILE columns), so the epilogue of tile i
// overlaps the loads and MMAs of tile i+1.
//
// PAIR: the kernel is launched as clusters of two CTAs (one TPC).  A pair tile is 2*TM positions x N_TILE
// channels: CTA r owns TM of them (its own A halo tiles, accumulators and epilogue) and loads rows
// [r*N_TILE/2, (r+1)*N_TILE/2) of every weight tile; the leader CTA (rank 0) issues
// tcgen05.mma.cta_group::2 (M = 256) for both.  Per SM and per MMA the shared-memory port then serves
// 128 A rows + N_TILE/2 B rows instead of 128 + N_TILE, which is what bounds the single-CTA kernel.
//   full barriers : each CTA's TMA completes on its own barrier; rank 1's warp 1 relays every completed phase to
//                   the leader's barrier (count 2 there: local producer + relay)
//   empty barriers: tcgen05.commit multicast arrives in both CTAs
//   accumulator   : acc_full by multicast commit; acc_empty lives in the leader (both CTAs' epilogue warps arrive)
//
// CHAIN (struct ConvChain below): the four convs of a ResNet stage in one launch, work items ordered layer-major, tile
// (l, t) waiting on the completion counters of tiles (l-1, t-1 .. t+1): a layer's last partial wave and the next layer's
// first wave run together instead of draining the machine at every layer boundary.
//
// TAPS: 16 = the stem's 4x4 window (K = 16 per tap, so only MT MMAs per tap): the MMA warp issues one window row
// (4 taps) per loop iteration with arithmetic shifts; 0 = per-group tap tables (3x3, stride-2 phases, fc).
//
// POOL (stem): conv + BN + ReLU + 3x3/s2 max-pool in one kernel, the 112x112x64 stem tensor never reaches HBM.
// Sub-tile mt of a tile is one *conv row* (lane = column, so W + 2 <= 128) and a tile is four consecutive rows
// 4t..4t+3 of one crop; a work item is a band of a crop (a quarter; an eighth-ish for small batches), walked top to
// bottom by one CTA:
//   vertical max  : the four rows of a column sit in the same TMEM lane -> pure register work; pooled row 2t
//                   needs conv row 4t-1, which the thread carries over from the previous tile (a one-row
//                   "carry" tile opens every quarter)
//   horizontal max: neighbouring columns are neighbouring lanes -> two warp shuffles per packed bf16x2 register;
//                   the single cross-warp neighbour (lane 0 <- lane 31 of the previous lane quarter) goes through
//                   a 2 KB shared-memory exchange
#pragma once
#include <type_traits>

#include "common.cuh"

namespace flope {

constexpr int kMaxGroups = 32;
constexpr int kMaxTaps = 16;
constexpr int kEpiWarps = 16;                  // 4 per TMEM lane quarter
constexpr int kConvThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxASlots = 8;
constexpr int kMaxBSlots = 16;
constexpr int kPoolXchBytes = 2 * 4 * 4 * 16 * 4;

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
  // ---- B operand (weights, packed [n_tile][tile][kc8][N_TILE][8] bf16; PAIR: [n_tile][tile][rank][kc8][N_TILE/2][8]) ----
  const __nv_bfloat16* wgt;
  // ---- position space (validity + (n,h,w) decode) ----
  int n_positions;             // N*Hp*Wp
  int Hp, Wp, H, W;
  int n_work;                  // work items dealt round-robin to CTAs (pairs): (pair) tiles, or POOL work items (pairs of them)
  int n_n_tiles;
  // ---- epilogue ----
  const float* bias;           // [Cout] folded BN bias / fc bias (the BN scale is folded into the packed weights)
  int relu;
  int out_mode;
  int Cout;
  void* out;
  long long out_plane;
  int out_base, out_Hp, out_Wp;        // for OUT_PARITY these describe the half-resolution grid; POOL: the pooled grid
  const __nv_bfloat16* res;            // residual (plain layout) or nullptr
  long long res_plane;
  int res_base, res_Hp, res_Wp;
  // ---- smem ring sizes ----
  int n_a_slots, n_b_slots;
  int pool_rows;               // POOL: conv rows per work item (a multiple of 4 that divides H); 0 otherwise
  int pool_split;              // POOL: work items per crop (H / pool_rows)
  int b_resident;              // all weight tiles fit the ring and n_n_tiles == 1: load them once per CTA
  // ---- second A source: K groups >= first_group2 are read from in2 (the projection shortcut of a
  //      ResNet block folded into its conv2 as extra K: same position space, shift 0) ----
  const __nv_bfloat16* in2;
  long long in2_plane;
  int in2_base;
  int first_group2;            // == n_groups when there is no second source
  int res_layer;               // chain index of the layer that writes `res` inside this launch, or -1
};

// A chain = up to four layers of one ResNet stage (same position space, same tile configuration) in ONE persistent
// launch.  Work items are ordered layer-major, (layer l, tile t) = l * n_work + t, and dealt round-robin as before, so
// the CTAs that finish layer l's last partial wave start on layer l+1 at once.  Tile (l, t) reads the outputs of tiles
// (l-1, t-1 .. t+1) (its halo never reaches further) - and, being the only reader of those ranges in layer l-1's
// input buffer, may also overwrite that buffer's tile t once they are done; every finished tile publishes itself in
// `flags` (a counter per position tile: n_n_tiles x CTAs of a pair parts), consumers wait on the three counters.
constexpr int kMaxChain = 4;
constexpr int kMaxSplit = 4;
// trunk_chain_kernel, latency tiles: weight tiles (taps) per ring slot.  With 4 KB tiles and four short MMAs per tap the
// per-slot handshakes (producer: empty -> expect -> copy; relay across the pair; issuer: full -> fence -> commit) pace the
// K loop, not the tensor pipe (profiles/r2_notes.md): one barrier round per three taps.
constexpr int kLatencyTapsPerSlot = 3;
struct ConvChain {
  ConvParams L[kMaxChain];
  int n_layers;
  int n_m_tiles;               // position tiles per layer
  uint32_t* flags;             // [n_layers][n_m_tiles] completion counters, zeroed before the launch (nullptr: single layer)
  uint32_t expected;           // counter value of a finished position tile
  uint32_t* fail;              // device word set when a dependency wait timed out (see wait_tile_flag); zeroed with the flags
  uint32_t* host_err;          // the same, in mapped host memory, sticky until the host reads it
  // Dynamic work claiming (optional, chains only): items are handed out in index order by an atomic counter instead
  // of round-robin by block index.  Every claimed item then belongs to a CTA that is RESIDENT, and all the items it
  // waits on have lower indices, i.e. were claimed earlier by resident CTAs - so the tile-flag waits cannot deadlock
  // even when the grid is only partly resident (several engines sharing the device), without a cooperative launch.
  uint32_t* counter;           // next unclaimed item, zeroed before the launch (nullptr: static round-robin)
  int claim_static;
  // Optional phase stamps (diagnosis, tools/timeline.py): [CTA][kStampWords] = %globaltimer (ns) at entry, prologue
  // done, dependency wait done, first operands landed, last MMA issued, first accumulator ready, last store issued,
  // exit.  (Not clock64: the SM cycle counter was measured to advance at 0.45-0.85 of the SM clock over a CTA's
  // life - it does not count while every warp of the SM is parked in a barrier wait.)  nullptr in production.
  unsigned long long* stamps;
  uint32_t* claims;            // [CTA pairs][kClaimRing] leader -> peer hand-over of the claimed items (pair kernels)
  // Split-K (trunk_chain_kernel, latency tiles only; 0 or 1 = off).  A streamed frame with a handful of flowers leaves
  // layer3 / layer4 with 28 / 16 tiles for 74 CTA pairs, each a serial chain of 144 / 288 MMAs: every (position, channel)
  // tile becomes k_splits consecutive work items over disjoint K-group ranges.  Items 0 .. k_splits-2 store their raw
  // fp32 accumulators in split_ws and count themselves in split_flags; the last one adds them to its own in index
  // order (a fixed order: results do not depend on timing) and runs the usual epilogue.
  int k_splits;
  uint32_t* split_flags;       // [n_layers][n_work] partial counters, zeroed with the tile flags
  float4* split_ws;            // [n_work][k_splits-1][2 ranks][N_TILE/4 column quads][128 rows] of this stage
  unsigned char split_group[kMaxChain][kMaxSplit + 1];    // K-group range of every split, per layer
  unsigned short split_tap[kMaxChain][kMaxSplit + 1];     // weight tiles (taps) ahead of every split
};
constexpr int kStampWords = 8;
constexpr int kClaimQ = 4;     // claimed items a CTA may hold ahead of its epilogue
constexpr int kClaimRing = 8;  // hand-over slots per CTA pair (> kClaimQ + the claims in flight)

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until the counter reaches `expected`.  The wait is bounded in TIME (a tile whose producer CTA never became
// resident - another context's kernels hold the SMs - must not hang the box): after kTileWaitNs the CTA records the
// failure in `fail` (device word, zeroed with the flags; every other wait of the launch then gives up at once) and in
// `host_err` (mapped host word: the next API call on any engine of the process reports FLOPE_ECUDA), and carries on
// with whatever is in memory, so that the launch terminates instead of trapping the context.
constexpr unsigned long long kTileWaitNs = 5000000000ull;
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_tile_flag(const uint32_t* flag, uint32_t expected, uint32_t* fail, uint32_t* host_err) {
  if (ld_acquire_gpu(flag) >= expected) return;
  const unsigned long long t0 = global_timer_ns();
  uint32_t spins = 0;
  while (ld_acquire_gpu(flag) < expected) {
    if ((++spins & 1023u) == 0) {
      if (fail && ld_acquire_gpu(fail)) return;
      if (global_timer_ns() - t0 > kTileWaitNs) {
        if (fail) atomicExch(fail, 1u);
        if (host_err) { atomicExch_system(host_err, 1u); __threadfence_system(); }
        return;
      }
    }
  }
}

__host__ __device__ constexpr int pow2_at_least(int v) { int r = 32; while (r < v) r <<= 1; return r; }

// 32 lanes x 16 columns of fp32 accumulator -> 16 registers
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <int N_TILE, int MT, int KP, bool POOL, bool PAIR, int TAPS>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvChain ch) {
  const ConvParams& p = ch.L[0];             // everything the layers of a chain share: geometry, tile counts, ring sizes, halo
  static_assert(!POOL || (N_TILE == 64 && MT == 4), "the pooled stem uses four conv rows x 64 channels per tile");
  constexpr int TM = MT * 128;
  constexpr int NB_ROWS = PAIR ? N_TILE / 2 : N_TILE;            // weight rows this CTA holds per tile
  constexpr int ACC_COLS = pow2_at_least(N_TILE * MT);          // column stride of one accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "two accumulator stages must fit the 512 TMEM columns");
  constexpr uint32_t IDESC = umma_idesc_bf16(PAIR ? 256 : 128, N_TILE);
  constexpr int NCHUNK = N_TILE / 32;
  constexpr int KC8 = 2 * KP;                                   // planes per K group
  constexpr int TILE_POS = PAIR ? 2 * TM : TM;                  // positions per (pair) tile

  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_empty = a_full + kMaxASlots;
  uint64_t* b_full = a_empty + kMaxASlots;
  uint64_t* b_empty = b_full + kMaxBSlots;
  uint64_t* acc_full = b_empty + kMaxBSlots;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint64_t* q_full = acc_empty + 3;                              // dynamic claiming: item queue producer warp -> other warps
  uint64_t* q_empty = q_full + kClaimQ;
  int* q_item = reinterpret_cast<int*>(q_empty + kClaimQ);
  static_assert((2 * kMaxASlots + 2 * kMaxBSlots + 4 + 1 + 2 * kClaimQ) * 8 + kClaimQ * 4 <= 512, "barrier area");
  float* s_bias = reinterpret_cast<float*>(smem_raw + 512);
  const int halo = p.halo_before + p.halo_after;
  // positions one tile spans in the halo tile: TM consecutive ones, or (POOL) four rows of 128 lanes, Wp apart
  const int span = POOL ? 3 * p.Wp + 128 : TM;
  const uint32_t mt_stride = POOL ? (uint32_t)p.Wp : 128u;      // sub-tile mt starts mt_stride positions further
  const uint32_t a_plane_bytes = (uint32_t)(span + halo) * 16u;
  const uint32_t a_slot_bytes = a_plane_bytes * KC8;
  constexpr uint32_t b_tile_bytes = (uint32_t)KC8 * NB_ROWS * 16u;
  uint8_t* a_ring = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 512 + (size_t)ch.n_layers * p.Cout * sizeof(float) + 127) & ~uintptr_t(127));
  uint8_t* b_ring = a_ring + (size_t)p.n_a_slots * a_slot_bytes;
  // POOL: lane-31 hand-over between the four lane-quarter warps of a channel slice: [parity][slice][quarter][2 rows][8]
  uint32_t* pool_xch = reinterpret_cast<uint32_t*>(b_ring + (size_t)p.n_b_slots * b_tile_bytes);

  // broadcast from lane 0 so that the compiler knows the warp index is warp-uniform: the role branches become
  // uniform branches and the MMA / TMA issue code can live in the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs)
  const int first_work = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // work is dealt to pairs (or CTAs) round-robin
  const int work_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_per_work = POOL ? (p.pool_rows >> 2) + 1 : 1; // POOL: one carry tile + pool_rows/4 four-row tiles
  const int total_work = ch.n_layers * p.n_work;                 // layer-major work items of the chain
  const bool dyn = ch.counter != nullptr;
  // k-th work item of this CTA (pair), or -1 after the last one.  Static: round-robin by block index.  Dynamic: the
  // producer warp claims it (claim_item below) and queues it in shared memory for the other warps.
  auto take_item = [&](int k) -> int {
    if (!dyn) { const int gw = first_work + k * work_stride; return gw < total_work ? gw : -1; }
    mbar_wait(&q_full[k & (kClaimQ - 1)], (uint32_t)(k / kClaimQ) & 1u);
    return q_item[k & (kClaimQ - 1)];
  };
  auto done_item = [&](int k) {                                  // this warp has finished the k-th item (non-producer warps)
    if (!dyn) return;
    __syncwarp();
    if (lane == 0) mbar_arrive(&q_empty[k & (kClaimQ - 1)]);
  };
  // The leader keeps ONE claim in flight (item k+1, issued while item k is being loaded), so its producer never waits
  // for an atomic's round trip after the first.  Not more: claims issued back to back by one CTA come back as
  // consecutive items, and a CTA that holds (i, i+1, i+2) works through a layer's first tiles one after the other
  // while the CTAs that drew the next layer's first tiles wait for them - measured +23 % on the layer4 chain.
  int c_cur = 0, c_nxt = 0;
  auto claim_item = [&](int k) -> int {                          // producer warp only
    if (!dyn) { const int gw = first_work + k * work_stride; return gw < total_work ? gw : -1; }
    const int slot = k & (kClaimQ - 1);
    mbar_wait(&q_empty[slot], ((uint32_t)(k / kClaimQ) & 1u) ^ 1u);   // the item that used this queue slot is finished
    uint32_t* hand = ch.claims + (size_t)first_work * kClaimRing;
    auto enc = [&](int kk, int item) { return (((uint32_t)(kk + 1) & 0xFFFu) << 20) | (item >= total_work ? 0xFFFFFu : (uint32_t)item); };
    int gw;
    if (ch.claim_static) {                                       // diagnosis: the static sequence through the item queue
      gw = first_work + k * work_stride;
      if (gw >= total_work) gw = -1;
    } else if (rank == 0) {
      if (lane == 0) {
        c_cur = k == 0 ? (int)atomicAdd(ch.counter, 1u) : c_nxt;
        if (PAIR) *reinterpret_cast<volatile uint32_t*>(hand + (k & (kClaimRing - 1))) = enc(k, c_cur);
        c_nxt = (int)atomicAdd(ch.counter, 1u);                  // item k+1: not needed before the next call
      }
      gw = __shfl_sync(0xffffffffu, c_cur, 0);
      if (gw >= total_work) gw = -1;
    } else {
      const uint32_t tag = (uint32_t)(k + 1) & 0xFFFu;
      uint32_t v, spins = 0;
      while (((v = ld_acquire_gpu(hand + (k & (kClaimRing - 1)))) >> 20) != tag) {
        if (++spins > (1u << 22)) { printf("flope: claim hand-over timeout block=%d\n", blockIdx.x); __trap(); }
      }
      gw = (v & 0xFFFFFu) == 0xFFFFFu ? -1 : (int)(v & 0xFFFFFu);
    }
    if (lane == 0) { q_item[slot] = gw; mbar_arrive(&q_full[slot]); }
    return gw;
  };

  // First position of this CTA's part of tile `tt` of work item `w`.
  //   plain: tile w / n_n_tiles covers TILE_POS consecutive positions; CTA r of a pair takes the r-th half
  //   POOL : work item (2w + r for pairs) = (crop, row quarter); tile tt covers conv rows q*R + 4(tt-1) .. +3
  auto tile_first_pos = [&](int w, int tt) -> int {
    if (!POOL) return (w / p.n_n_tiles) * TILE_POS + (PAIR ? (int)rank * TM : 0);
    const int unit = PAIR ? 2 * w + (int)rank : w;
    const int n = unit / p.pool_split, q = unit - n * p.pool_split;
    return (n * p.Hp + q * p.pool_rows + 4 * (tt - 1)) * p.Wp;
  };

  unsigned long long* stamps = ch.stamps ? ch.stamps + (size_t)blockIdx.x * kStampWords : nullptr;
  auto stamp = [&](int i) {
    if (!stamps) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    stamps[i] = t;
  };
  if (threadIdx.x == 0) stamp(0);
  if (threadIdx.x == 0) {
    const uint32_t full_count = (PAIR && rank == 0) ? 2u : 1u;   // leader: own producer + the peer's relay
    for (int i = 0; i < p.n_a_slots; ++i) { mbar_init(&a_full[i], full_count); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.n_b_slots; ++i) { mbar_init(&b_full[i], full_count); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 2 * kEpiWarps : kEpiWarps); }
    for (int i = 0; i < kClaimQ; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1 + kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc2(tmem_ptr, TMEM_COLS); tmem_relinquish2(); }
    else { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  }
  for (int l = 0; l < ch.n_layers; ++l)
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias[l * p.Cout + i] = ch.L[l].bias[i];
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) stamp(1);
  // everything above touched only static data (bias) and on-chip state; the activations of the previous layer
  // are complete and visible after this point, and the next kernel may begin its own prologue
  griddep_wait();
  griddep_launch();
  if (threadIdx.x == 0) stamp(2);

  if (warp == 0) {
    // ===================== TMA producer: whole warp runs the loop, one lane issues =====================
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t a_ring_addr = smem_u32(a_ring);
    const uint32_t b_ring_addr = smem_u32(b_ring);
    int a_slot = 0, b_slot = 0;
    uint32_t a_phase = 0, b_phase = 0;
    bool first = true;
    for (int k = 0;; ++k) {
      const int gw = claim_item(k);
      if (gw < 0) break;
      const int l = gw / p.n_work, w = gw - l * p.n_work;
      const ConvParams& q = ch.L[l];
      if (l > 0) {
        // the producing layer's tiles under this tile's halo must be complete (and visible to the TMA unit)
        const int m = w / p.n_n_tiles;
        const uint32_t* fl = ch.flags + (size_t)(l - 1) * ch.n_m_tiles;
        if (m > 0) wait_tile_flag(fl + m - 1, ch.expected, ch.fail, ch.host_err);
        wait_tile_flag(fl + m, ch.expected, ch.fail, ch.host_err);
        if (m + 1 < ch.n_m_tiles) wait_tile_flag(fl + m + 1, ch.expected, ch.fail, ch.host_err);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      for (int tt = 0; tt < tiles_per_work; ++tt) {
        const int n_tile = POOL ? 0 : w % p.n_n_tiles;
        const int tile_start = tile_first_pos(w, tt);
        // PAIR: weights are packed [n_tile][tile][rank][k8][N_TILE/2][8], so each CTA's half is one contiguous copy
        const __nv_bfloat16* wtile = q.wgt + ((size_t)n_tile * q.taps_total * (PAIR ? 2 : 1) + rank) * (b_tile_bytes / 2);
        for (int g = 0; g < q.n_groups; ++g) {
          mbar_wait(&a_empty[a_slot], a_phase ^ 1);
          mbar_expect_tx_if(leader, &a_full[a_slot], a_slot_bytes);
          const uint32_t a_dst = a_ring_addr + a_slot * a_slot_bytes;
          const bool second = g >= q.first_group2;
          const long long a_plane = second ? q.in2_plane : q.in_plane;
          const __nv_bfloat16* src = (second ? q.in2 : q.in) +
              ((long long)q.group_plane[g] * a_plane + (second ? q.in2_base : q.in_base) + tile_start - p.halo_before) * 8;
#pragma unroll
          for (int j = 0; j < KC8; ++j)
            bulk_g2s_if(leader, a_dst + j * a_plane_bytes, src + (long long)j * a_plane * 8, a_plane_bytes, &a_full[a_slot]);
          if (++a_slot == p.n_a_slots) { a_slot = 0; a_phase ^= 1; }
          const int ntaps = q.group_ntaps[g];
          if (!p.b_resident || first) {
            for (int t = 0; t < ntaps; ++t) {
              mbar_wait(&b_empty[b_slot], b_phase ^ 1);
              mbar_expect_tx_if(leader, &b_full[b_slot], b_tile_bytes);
              bulk_g2s_if(leader, b_ring_addr + b_slot * b_tile_bytes, wtile, b_tile_bytes, &b_full[b_slot]);
              wtile += (PAIR ? 2 : 1) * (b_tile_bytes / 2);
              if (++b_slot == p.n_b_slots) { b_slot = 0; b_phase ^= 1; }
            }
          }
        }
        first = false;
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
    bool first = true;
    for (int k = 0;; ++k) {
      const int gw = take_item(k);
      if (gw < 0) break;
      const ConvParams& q = ch.L[gw / p.n_work];
      for (int tt = 0; tt < tiles_per_work; ++tt) {
        for (int g = 0; g < q.n_groups; ++g) {
          mbar_wait(&a_full[a_slot], a_phase);
          mbar_arrive_remote_if(leader, a_full_remote + a_slot * 8);
          const int ntaps = q.group_ntaps[g];
          for (int t = 0; t < ntaps; ++t) {
            if (!p.b_resident || first) {
              mbar_wait(&b_full[b_slot], b_phase);
              mbar_arrive_remote_if(leader, b_full_remote + b_slot * 8);
            }
            if (++b_slot == p.n_b_slots) { b_slot = 0; b_phase ^= (p.b_resident ? 0u : 1u); }
          }
          if (++a_slot == p.n_a_slots) { a_slot = 0; a_phase ^= 1; }
        }
        first = false;
      }
      done_item(k);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: whole warp runs the loop, one lane issues =====================
    // Descriptors (K-major, no swizzle): A planes are a_plane_bytes apart in K, 8-pixel groups 128 B apart;
    // weight tiles are [k8][NB_ROWS][8], so K-adjacent core matrices are NB_ROWS*16 B apart.  Only the
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
    for (int k = 0;; ++k) {
      const int gw = take_item(k);
      if (gw < 0) break;
      const ConvParams& q = ch.L[gw / p.n_work];
      for (int tt = 0; tt < tiles_per_work; ++tt, ++it) {
        const uint32_t stage = it & 1;
        // POOL: the carry tile that opens a work item only needs its last row (a warp-uniform branch: the issue
        // predicate itself must stay the elected lane, or ptxas wraps every MMA in an ELECT/R2UR loop)
        const bool carry_tile = POOL && tt == 0;
        mbar_wait(&acc_empty[stage], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + stage * ACC_COLS;
        uint32_t accumulate = 0;
        // NT consecutive weight tiles starting at the current slot, tap j reading the halo tile at a_tap0 + j
        // (NT = 1: one filter tap; NT = 4: one row of the stem's 4x4 window).  Waits for the weight tiles, issues the
        // NT*KP*MT MMAs back to back (every operand = a loop-carried base + compile-time constants), then commits
        // the slots these MMAs were the last users of.
        auto issue_taps = [&](auto nt_tag, uint32_t a_tap0, bool last_of_group, bool last_group) {
          constexpr int NT = decltype(nt_tag)::value;
          if (!p.b_resident || it == 0) {
#pragma unroll
            for (int j = 0; j < NT; ++j) mbar_wait(&b_full[b_slot + j], b_phase);
          }
          tc_fence_after();
          const uint32_t b_tap = b_lo0 + b_slot * b_tile_units;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                if (POOL && carry_tile && mt < MT - 1) continue;
                const uint32_t acc_flag = (j == 0 && k == 0) ? accumulate : 1u;
                if (PAIR)
                  umma2_bf16_if(leader, acc + mt * N_TILE, a_tap0 + j + k * a_kstep + mt * mt_stride, desc_hi,
                                b_tap + j * b_tile_units + k * b_kstep, desc_hi, IDESC, acc_flag);
                else
                  umma_bf16_if(leader, acc + mt * N_TILE, a_tap0 + j + k * a_kstep + mt * mt_stride, desc_hi,
                               b_tap + j * b_tile_units + k * b_kstep, desc_hi, IDESC, acc_flag);
              }
            }
          }
          // commits free the weight slot / halo slot (in both CTAs of a pair) once these MMAs retire
          if (!p.b_resident) {
#pragma unroll
            for (int j = 0; j < NT; ++j) { if (PAIR) tc_commit2_if(leader, &b_empty[b_slot + j]); else tc_commit_if(leader, &b_empty[b_slot + j]); }
          }
          if (last_of_group) {
            if (PAIR) tc_commit2_if(leader, &a_empty[a_slot]); else tc_commit_if(leader, &a_empty[a_slot]);
            if (last_group) { if (PAIR) tc_commit2_if(leader, &acc_full[stage]); else tc_commit_if(leader, &acc_full[stage]); }
          }
          accumulate = 1;
          b_slot += NT;
          if (b_slot == p.n_b_slots) { b_slot = 0; b_phase ^= (p.b_resident ? 0u : 1u); }
        };
        for (int g = 0; g < q.n_groups; ++g) {
          mbar_wait(&a_full[a_slot], a_phase);
          if (it == 0 && g == 0 && lane == 0) stamp(3);
          const uint32_t a_grp = a_lo0 + a_slot * a_slot_units;
          const bool last_group = g == q.n_groups - 1;
          if (TAPS == 16) {
            // stem: 4x4 stride-1 window on the space-to-depth grid, one window row (4 taps, shifts +0..+3) per
            // iteration; all 16 weight tiles are resident (n_b_slots == 16), so b_slot + j never wraps mid-row
            uint32_t a_row = a_grp - (uint32_t)(2 * p.Wp + 2);
#pragma unroll 1
            for (int r = 0; r < 4; ++r, a_row += (uint32_t)p.Wp)
              issue_taps(std::integral_constant<int, 4>{}, a_row, r == 3, last_group);
          } else {
            // per-group tap tables (3x3, stride-2 phases, 1x1, fc)
            const int tofs = q.group_tapofs[g];
            const int ntaps = q.group_ntaps[g];
            for (int t = 0; t < ntaps; ++t)
              issue_taps(std::integral_constant<int, 1>{}, a_grp + (uint32_t)q.tap_shift[tofs + t], t == ntaps - 1, last_group);
          }
          if (++a_slot == p.n_a_slots) { a_slot = 0; a_phase ^= 1; }
        }
      }
      done_item(k);
    }
    if (lane == 0) stamp(4);
  } else {
    // ===================== epilogue: kEpiWarps warps, 4 per TMEM lane quarter =====================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;       // share of the (mt, 32-column chunk) list; POOL: 16-channel slice
    constexpr int NSUB = kEpiWarps / 4;
    const int img = p.Hp * p.Wp;
    const uint32_t acc_empty_remote = PAIR ? mapa_u32(smem_u32(acc_empty), 0) : 0u;   // the leader's acc_empty barriers
    auto release_acc = [&](uint32_t stage) {       // this warp's last TMEM read of the stage is done
      tc_fence_before();
      __syncwarp();
      if (PAIR) mbar_arrive_remote_if(lane == 0 ? 1u : 0u, acc_empty_remote + stage * 8);
      else if (lane == 0) mbar_arrive(&acc_empty[stage]);
    };
    uint32_t it = 0;
    for (int k = 0;; ++k) {
      const int gw = take_item(k);
      if (gw < 0) break;
      const int l = gw / p.n_work, w = gw - l * p.n_work;
      const ConvParams& q = ch.L[l];
      const float* bias_l = s_bias + l * p.Cout;
      if constexpr (POOL) {
        // ---------- stem: conv + BN + ReLU rows -> 3x3/s2 max-pool, two pooled rows per four-row tile ----------
        const int unit = PAIR ? 2 * w + (int)rank : w;
        const int n = unit / p.pool_split, q = unit - n * p.pool_split;
        const int col = quarter * 32 + lane;                  // conv column of this thread
        const bool col_ok = col < p.W;
        float bias16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) bias16[j] = bias_l[sub * 16 + j];
        uint32_t carry[8];                                    // conv row 4t-1 (post-ReLU bf16x2), 16 channels
#pragma unroll
        for (int j = 0; j < 8; ++j) carry[j] = 0u;
        for (int tt = 0; tt < tiles_per_work; ++tt, ++it) {
          const uint32_t stage = it & 1;
          const uint32_t acc = tmem_base + stage * ACC_COLS + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sub * 16);
          mbar_wait(&acc_full[stage], (it >> 1) & 1);
          tc_fence_after();
          uint32_t a[4][8];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            if (tt == 0 && mt < 3) {                          // carry tile: only its last row was computed
#pragma unroll
              for (int j = 0; j < 8; ++j) a[mt][j] = 0u;
              continue;
            }
            uint32_t v16[16];
            tmem_ld16(acc + (uint32_t)(mt * N_TILE), v16);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              a[mt][j] = col_ok ? pack_bf16x2_relu(__uint_as_float(v16[2 * j]) + bias16[2 * j],
                                                   __uint_as_float(v16[2 * j + 1]) + bias16[2 * j + 1])
                                : 0u;                         // pad / out-of-row columns are the pool's zero padding
          }
          release_acc(stage);
          if (tt == 0) {
            // the row above the first quarter is the pool's padding (post-ReLU values are >= 0, so 0 is neutral)
#pragma unroll
            for (int j = 0; j < 8; ++j) carry[j] = q == 0 ? 0u : a[3][j];
            continue;
          }
          uint32_t v[2][8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[0][j] = bf16x2_max_u32(bf16x2_max_u32(carry[j], a[0][j]), a[1][j]);
            v[1][j] = bf16x2_max_u32(bf16x2_max_u32(a[1][j], a[2][j]), a[3][j]);
            carry[j] = a[3][j];
          }
          // horizontal: pooled column j = max(conv columns 2j-1, 2j, 2j+1); even lanes own a pooled column
          uint32_t* xch = pool_xch + (((it & 1) * 4 + sub) * 4 + quarter) * 16;
          if (lane == 31) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { xch[j] = v[0][j]; xch[8 + j] = v[1][j]; }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + sub) : "memory");   // the four lane-quarter warps of this channel slice
          uint32_t o[2][8];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t left = __shfl_up_sync(0xffffffffu, v[r][j], 1);
              const uint32_t right = __shfl_down_sync(0xffffffffu, v[r][j], 1);
              if (lane == 0) left = quarter == 0 ? 0u : *(xch - 16 + r * 8 + j);   // lane 31 of the previous lane quarter
              o[r][j] = bf16x2_max_u32(bf16x2_max_u32(left, v[r][j]), right);
            }
          }
          if (!(lane & 1) && col_ok) {
            const int orow = ((q * p.pool_rows) >> 1) + 2 * (tt - 1);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                ((long long)(2 * sub) * p.out_plane + p.out_base + ((long long)n * p.out_Hp + orow) * p.out_Wp + (col >> 1)) * 8;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              *reinterpret_cast<uint4*>(dst + (long long)r * p.out_Wp * 8) = make_uint4(o[r][0], o[r][1], o[r][2], o[r][3]);
              *reinterpret_cast<uint4*>(dst + ((long long)r * p.out_Wp + p.out_plane) * 8) = make_uint4(o[r][4], o[r][5], o[r][6], o[r][7]);
            }
          }
        }
      } else {
        const uint32_t stage = it & 1;
        const int n_tile = w % p.n_n_tiles;
        const int tile_start = tile_first_pos(w, 0);
        const int cout_base = n_tile * N_TILE;
        const uint32_t acc = tmem_base + stage * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
        bool waited = false;
        if (q.res_layer >= 0) {                 // residual written earlier in this launch: its tile must be published
          wait_tile_flag(ch.flags + (size_t)q.res_layer * ch.n_m_tiles + w / p.n_n_tiles, ch.expected, ch.fail, ch.host_err);
          __syncwarp();
        }
#pragma unroll 1
        for (int c = sub; c < MT * NCHUNK; c += NSUB) {
          const int mt = c / NCHUNK;
          const int c0 = (c - mt * NCHUNK) * 32;
          const int pos = tile_start + mt * 128 + quarter * 32 + lane;
          const int n = pos / img;
          const int r = pos - n * img;
          const int h = r / p.Wp;
          const int ww = r - h * p.Wp;
          const bool valid = pos >= 0 && pos < p.n_positions && h < p.H && ww < p.W;
          const int plane0 = (cout_base + c0) >> 3;
          uint4 res[4];
          if (q.res != nullptr && valid) {     // issued before the accumulator wait: latency hidden behind the MMAs
            // (L2 loads: inside a chain the residual may have been written by another SM moments ago)
            const __nv_bfloat16* rp = q.res + ((long long)plane0 * q.res_plane + q.res_base + ((long long)n * q.res_Hp + h) * q.res_Wp + ww) * 8;
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) res[j8] = __ldcg(reinterpret_cast<const uint4*>(rp + (long long)j8 * q.res_plane * 8));
          }
          if (!waited) {
            mbar_wait(&acc_full[stage], (it >> 1) & 1);
            tc_fence_after();
            waited = true;
            if (it == 0 && threadIdx.x == 64) stamp(5);
          }
          uint32_t v32[32];
          tmem_ld32(acc + (uint32_t)(mt * N_TILE + c0), v32);
          tmem_ld_wait();
          if (c + NSUB >= MT * NCHUNK) release_acc(stage);   // hand the accumulator back to the MMA warp
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
            if (q.out_mode == OUT_F32_ROWS) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(q.out) + (long long)pos * p.Cout + cout_base + c0);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (q.relu) dst[j >> 2] = make_float4(fmaxf(v[j], 0.f), fmaxf(v[j + 1], 0.f), fmaxf(v[j + 2], 0.f), fmaxf(v[j + 3], 0.f));
                else dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            } else {
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
        }
        if (!waited) {
          // a warp with no chunk in this configuration still takes part in the accumulator hand-back
          mbar_wait(&acc_full[stage], (it >> 1) & 1);
          release_acc(stage);
        }
        if (ch.flags != nullptr) {
          // publish this CTA's part of the tile: every epilogue warp has issued its stores -> barrier -> one thread
          // fences (cumulative at GPU scope) and bumps the tile's counter
          asm volatile("bar.sync 5, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (threadIdx.x == 64) {
            __threadfence();
            atomicAdd(ch.flags + (size_t)l * ch.n_m_tiles + w / p.n_n_tiles, 1u);
          }
        }
        ++it;
      }
      done_item(k);
    }
    if (threadIdx.x == 64) stamp(6);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(7);
  if (PAIR) cluster_sync_all();            // the peer may still read this CTA's smem / signal its barriers until here
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace flope

// flope_b200: engine + C ABI (include/flope_b200.h) for the B200-native FloPE pose path.
//
// Host side only orchestrates: it owns the blocked-pixel activation buffers, folds
// BatchNorm, packs weights into UMMA-ready tiles and launches the sm_100a kernels of
// roi_crop.cuh / conv_igemm.cuh / pointwise.cuh / pose_head.cuh on the caller's stream.
// There is no CPU compute path: without a CUDA device every entry point fails loudly.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/flope_b200.h"
#include "conv_igemm.cuh"
#include "depth.cuh"
#include "mask_post.cuh"
#include "pointwise.cuh"
#include "pose_head.cuh"
#include "trunk_chain.cuh"
#include "roi_crop.cuh"
#include "roi_stream.cuh"

using namespace flope;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(x)                                                                           \
  do {                                                                                        \
    cudaError_t _e = (x);                                                                     \
    if (_e != cudaSuccess) return fail(FLOPE_ECUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); \
  } while (0)

constexpr int kMaxSmem = 232448;      // 227 KB opt-in dynamic shared memory per CTA
constexpr int kMaxTM = 1024;

long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

Geom make_geom(int C, int H, int W, int pad, int max_batch) {
  Geom g;
  g.C = C; g.H = H; g.W = W; g.Hp = H + pad; g.Wp = W + pad;
  g.base = (int)round_up(8 * g.Wp + 8, 8);     // zero guard: the pooled stem's carry tile reaches 6 rows above a crop
  const long long npos = (long long)max_batch * g.Hp * g.Wp;
  g.plane = g.base + round_up(npos, 1024) + kMaxTM + round_up(4 * g.Wp + 8, 8);
  return g;
}

struct ActBuf {
  Geom g{};
  __nv_bfloat16* d = nullptr;
  int planes = 0;
  bool parity = false;
  size_t bytes() const { return (size_t)planes * g.plane * 8 * sizeof(__nv_bfloat16); }
};

enum ConvKind { K_CONV3 = 0, K_CONV3_S2 = 1, K_DOWN1_S2 = 2, K_STEM = 3, K_FC = 4 };

struct TapDef { int r, s; };

struct ConvLayer {
  std::string name, wkey, bnkey, biaskey;
  ConvKind kind;
  int cin, cout;
  int in_buf, out_buf, res_buf;
  int relu;
  int out_mode;
  bool pool = false;             // stem only: fused 3x3/s2 max-pool epilogue
  bool pair = false;             // CTA-pair kernel (cluster of 2, tcgen05 cta_group::2, M = 256)
  // projection shortcut folded into this conv as extra K groups (block 0 of layers 2-4):
  int short_buf = -1, cin2 = 0;
  std::string wkey2, bnkey2;
  // derived
  int n_tile = 64, mt = 4;
  size_t smem = 0;
  ConvParams p{};
  std::vector<std::vector<TapDef>> group_taps;   // per group: (r,s) of each weight tile, for packing
  std::vector<int> group_cin;                    // first input channel of each group
  __nv_bfloat16* d_w = nullptr;
  float* d_bias = nullptr;
};

template <int N_TILE, int MT, int KP, bool POOL, bool PAIR, int TAPS>
cudaError_t conv_set_attr() {
  return cudaFuncSetAttribute(conv_igemm_kernel<N_TILE, MT, KP, POOL, PAIR, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
}
// Launch with optional cluster size and programmatic dependent launch (the kernel may begin while its stream
// predecessor drains; every kernel here calls griddep_wait() before touching activations).
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, int pdl,
                     Args&&... args) {   // pdl: bit 0 = programmatic dependent launch, bit 1 = cooperative launch
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = 0;
  if (pdl & 2) {                                          // cooperative: the whole grid becomes resident at once or not at all
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;   // the two CTAs of a pair share a TPC
    attr[na].val.clusterDim.x = cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl & 1) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <int N_TILE, int MT, int KP, bool POOL, bool PAIR, int TAPS>
cudaError_t conv_launch_t(const ConvChain& c, dim3 grid, size_t smem, cudaStream_t st, int pdl) {
  return launch_k(conv_igemm_kernel<N_TILE, MT, KP, POOL, PAIR, TAPS>, grid, dim3(kConvThreads), smem, st, PAIR ? 2 : 1, pdl, c);
}

// (N_TILE, MT, KP, POOL, PAIR, TAPS): output channels per tile, 128-pixel sub-tiles per CTA and tile, K=16 MMAs per weight
// tile, fused 3x3/s2 max-pool epilogue (stem), CTA-pair (cta_group::2) kernel, tap issue order (16 = stem rows, 0 = tables)
#define FOR_EACH_CONV_CFG(X)                                                                                   \
  /* single-CTA */                                                                                             \
  X(64, 4, 4, false, false, 0) X(128, 2, 4, false, false, 0) X(64, 1, 4, false, false, 0)                                                   \
  X(64, 4, 1, false, false, 16) X(64, 4, 1, true, false, 16)                                                   \
  /* CTA pair */                                                                                               \
  X(64, 4, 4, false, true, 0) X(128, 2, 4, false, true, 0) X(256, 1, 4, false, true, 0) X(64, 1, 4, false, true, 0) \
  X(64, 4, 1, false, true, 16) X(64, 4, 1, true, true, 16)

cudaError_t conv_set_all_attrs() {
  cudaError_t e;
#define X(N, M, K, P, R, T) if ((e = conv_set_attr<N, M, K, P, R, T>()) != cudaSuccess) return e;
  FOR_EACH_CONV_CFG(X)
#undef X
  if ((e = cudaFuncSetAttribute(trunk_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(trunk_chain_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(trunk_chain_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(trunk_chain_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(trunk_chain_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
}
// returns false when no kernel instance matches; *err receives the launch status otherwise
bool conv_launch(int n_tile, int mt, bool pair, int taps, const ConvChain& c, dim3 grid, size_t smem, cudaStream_t st, int pdl,
                 cudaError_t* err) {
  const ConvParams& p = c.L[0];
#define X(N, M, K, P, R, T)                                                                                        \
  if (n_tile == N && mt == M && p.kc8 == 2 * K && (p.pool_rows > 0) == P && pair == R && taps == T) {              \
    *err = conv_launch_t<N, M, K, P, R, T>(c, grid, smem, st, pdl);                                                \
    return true;                                                                                                   \
  }
  FOR_EACH_CONV_CFG(X)
#undef X
  return false;
}

// ---------------------------------------------------------------------------------------------
// Process-wide plumbing for the launches that need their whole grid resident (stage chains, trunk launch)
// ---------------------------------------------------------------------------------------------
// Sticky error word in mapped host memory: a tile-dependency wait that timed out on the device sets it (conv_igemm.cuh:
// wait_tile_flag), the next API call reports FLOPE_ECUDA.  device = true returns the device-side alias.
uint32_t* host_err_word(bool device) {
  static uint32_t* h = nullptr;
  static uint32_t* d = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* hp = nullptr;
    if (cudaHostAlloc(&hp, sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return; }
    *static_cast<uint32_t*>(hp) = 0u;
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(hp); return; }
    h = static_cast<uint32_t*>(hp); d = static_cast<uint32_t*>(dp);
  });
  return device ? d : h;
}
bool take_device_timeout() {
  volatile uint32_t* h = host_err_word(false);
  if (h && *h) { *h = 0u; return true; }
  return false;
}

// The chain / trunk kernels deal tiles statically and spin on each other's completion flags: every CTA of such a launch
// must be resident, which holds as long as only ONE of them runs on the device at a time.  Engines are independent
// objects (the header allows calls on different engines from different threads and streams), so the library enforces
// it itself: a forward that contains such launches waits for the previous one on the same device (an event chain;
// the host mutex keeps wait -> launch -> record atomic across threads).  On one stream the wait is already satisfied
// and costs a microsecond of host time; per-layer launches (EnginePool's overlapping engines) never take the gate.
struct ResidencyGate {
  struct PerDevice { std::mutex mu; cudaEvent_t ev = nullptr; bool recorded = false; };
  static PerDevice& of(int dev) { static PerDevice g[64]; return g[dev & 63]; }
  PerDevice* g = nullptr;
  cudaStream_t st = nullptr;
  ResidencyGate(bool needed, int dev, cudaStream_t stream) {
    if (!needed) return;
    g = &of(dev); st = stream;
    g->mu.lock();
    if (!g->ev && cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); g->ev = nullptr; }
    if (g->ev && g->recorded) cudaStreamWaitEvent(st, g->ev, 0);
  }
  ~ResidencyGate() {
    if (!g) return;
    if (g->ev && cudaEventRecord(g->ev, st) == cudaSuccess) g->recorded = true;
    g->mu.unlock();
  }
};

}  // namespace

struct flope_engine {
  int device = 0, max_batch = 0, S = 0, num_sms = 148;
  bool weights_loaded = false;
  int launches = 0;
  std::vector<ActBuf> bufs;
  std::map<std::string, int> act_names;          // debug name -> buffer index
  std::vector<ConvLayer> layers;
  ConvLayer stem_pool;                           // stem conv with the max-pool fused into its epilogue
  int roi_item_floor = 14;
  int roi_item_auto = 1;                         // smaller items for launches that would not fill the persistent grid
  int roi_stream = 1;                            // streaming ROI kernels (roi3_kernel, roi_stream.cuh): the production path
  int roi_item_rows = 56;                        // output rows per work item of the streaming bilinear kernel
  int roi_item_rows8 = 128;                       // same for the streaming Lanczos4 kernel
  int roi_stage_kb = 14;                         // bytes per ring stage of the streaming kernels
  int roi_stages = 3;                            // ring depth
  int roi_ctas_per_sm = 0;                       // 0 = as many as fit
  bool chain_coop = false;                       // launch chains cooperatively (gang-scheduled): needed when several engines share a device
  unsigned long long* d_stamps = nullptr;        // phase stamps of the conv launches of one forward ("timeline" debug option)
  int stamp_launch = 0;
  bool fc_small = true;                          // fc on 128-row x 64-channel single-CTA tiles
  bool use_trunk = true;                         // layer1..layer4 as one launch when the plan allows (trunk_chain.cuh)
  int chain_dynamic = 0;                     // chains claim work items from an atomic counter (safe under partial residency)
  bool use_chain = true;                         // one persistent launch per ResNet stage (four convs) with per-tile completion flags
  std::vector<std::vector<int>> chains;          // layer indices of each stage
  uint32_t* d_roi_tab = nullptr;                 // Lanczos4 axis tables of the current ROI launch (grown on demand)
  size_t roi_tab_words = 0;
  int roi_axis_tab = 1;                          // Lanczos4: coefficient tables from a pre-kernel (0 = the producer warps compute them)
  unsigned int* d_roi_sched = nullptr;           // work counter of the streaming ROI kernels (self-resetting)
  int roi_dynamic = 1;                           // streaming ROI kernels claim items from the counter (0 = static round-robin)
  float4* d_split_ws = nullptr;                  // split-K partial accumulators of the trunk launch (latency tiles)
  int trunk_splitk = kMaxSplit;                  // most splits per tile (< 2: off)
  int trunk_split_stages = 15;                   // stages that may split (bit s)
  uint32_t* d_flags = nullptr;                   // completion counters of all chains, zeroed at the start of every forward
  size_t flags_per_chain = 0;                    // counters reserved per chain (kMaxChain layers x position tiles at max_batch)
  bool small_tiles = true;                       // latency-oriented tiles when max_batch is too small to fill the SMs
  bool use_pdl = true;                           // programmatic dependent launch between the backbone kernels
  bool use_pair = true;                          // CTA-pair (cta_group::2) conv kernels; flope_debug_set "pair" 0 selects the single-CTA ones
  bool fuse_pool = false;                        // stem conv + max-pool in one kernel (default whenever the crop side allows it)
  bool can_fuse_pool = false;
  int buf_x0 = -1, buf_stem = -1, buf_pool_in = -1, buf_pool = -1, buf_mp_out = -1;
  float* d_feat = nullptr;                       // (max_batch, 2048) fp32
  float* d_wrot = nullptr;                       // (9, 2048) fp32
  float* d_brot = nullptr;                       // (9,)
  float* d_r9 = nullptr;                         // (max_batch, 9) scratch
  int feat_dim = 2048;
  // CUDA graphs of the backbone (stem .. fc), one per batch size, captured on an engine-owned stream
  bool use_graph = true;
  cudaStream_t cap_stream = nullptr;
  struct GraphEntry { int n; int launches; cudaGraphExec_t exec; };
  std::vector<GraphEntry> graphs;
  // optional per-launch CUDA-event timing (bench.py roofline pass)
  bool profile = false;
  int profile_mode = 0;                          // 1: one event pair per launch; 2: one pair around the trunk's conv chain
  std::vector<std::string> prof_names;
  std::vector<cudaEvent_t> prof_ev;              // start/stop pairs, in launch order
};

namespace {

int add_buf(flope_engine* e, int C, int H, int W, int pad, bool parity, const char* dbg_name) {
  ActBuf b;
  b.g = make_geom(C, H, W, pad, e->max_batch);
  b.parity = parity;
  b.planes = (C / 8) * (parity ? 4 : 1);
  e->bufs.push_back(b);
  const int idx = (int)e->bufs.size() - 1;
  if (dbg_name) e->act_names[dbg_name] = idx;
  return idx;
}

void add_conv(flope_engine* e, const std::string& name, ConvKind kind, int cin, int cout, int in_buf, int out_buf,
              int res_buf, int relu, int out_mode, const std::string& wkey, const std::string& bnkey,
              const std::string& biaskey = "") {
  ConvLayer L;
  L.name = name; L.kind = kind; L.cin = cin; L.cout = cout;
  L.in_buf = in_buf; L.out_buf = out_buf; L.res_buf = res_buf;
  L.relu = relu; L.out_mode = out_mode;
  L.wkey = wkey; L.bnkey = bnkey; L.biaskey = biaskey;
  e->layers.push_back(L);
}

// Fill the group/tap tables and pick the tile configuration of one layer.
int plan_conv(flope_engine* e, ConvLayer& L) {
  const ActBuf& in = e->bufs[L.in_buf];
  ConvParams& p = L.p;
  std::memset(&p, 0, sizeof(p));
  const int Wp = in.g.Wp;
  p.in = in.d;
  p.in_plane = in.g.plane;
  p.in_base = in.g.base;
  p.Hp = in.g.Hp; p.Wp = in.g.Wp; p.H = in.g.H; p.W = in.g.W;
  L.group_taps.clear();
  L.group_cin.clear();
  int ntap_entries = 0;
  auto add_tap_entry = [&](int shift) { p.tap_shift[ntap_entries] = shift; return ntap_entries++; };
  switch (L.kind) {
    case K_CONV3: {
      p.kc8 = 8;
      p.n_groups = L.cin / 64;
      for (int t = 0; t < 9; ++t) add_tap_entry((t / 3 - 1) * Wp + (t % 3 - 1));
      for (int g = 0; g < p.n_groups; ++g) {
        p.group_plane[g] = g * 8; p.group_tapofs[g] = 0; p.group_ntaps[g] = 9;
        std::vector<TapDef> td;
        for (int t = 0; t < 9; ++t) td.push_back({t / 3, t % 3});
        L.group_taps.push_back(td);
        L.group_cin.push_back(g * 64);
      }
      p.halo_before = Wp + 1; p.halo_after = Wp + 1;
      p.first_group2 = p.n_groups;
      if (L.short_buf >= 0) {
        // out = relu(bn2(conv2(t)) + bn_d(conv1x1_s2(x))): with both BN scales folded into the weights the 1x1
        // stride-2 projection is just more K for the same accumulator.  x lives parity-split; its (0,0)
        // sub-grid (planes 0 .. cin2/8-1) is indexed by the same positions as t, shift 0 = tap entry 4.
        const ActBuf& sb = e->bufs[L.short_buf];
        if (sb.g.Hp != in.g.Hp || sb.g.Wp != in.g.Wp) return fail(FLOPE_EINVAL, "shortcut geometry mismatch");
        p.in2 = sb.d; p.in2_plane = sb.g.plane; p.in2_base = sb.g.base;
        for (int c = 0; c < L.cin2 / 64; ++c) {
          const int g = p.n_groups++;
          p.group_plane[g] = c * 8; p.group_tapofs[g] = 4; p.group_ntaps[g] = 1;
          L.group_taps.push_back({{1, 1}});
          L.group_cin.push_back(c * 64);
        }
      }
      break;
    }
    case K_CONV3_S2: {   // input is parity-split; its Geom is the half-resolution (= output) grid
      p.kc8 = 8;
      const int nch = L.cin / 64;
      p.n_groups = 4 * nch;
      int ofs[4], cnt[4];
      std::vector<TapDef> tds[4];
      for (int q = 0; q < 4; ++q) {
        const int ph = q >> 1, pw = q & 1;
        ofs[q] = ntap_entries; cnt[q] = 0;
        for (int r = 0; r < 3; ++r) {
          if (((r + 1) & 1) != ph) continue;          // input row 2i+r-1 has parity (r+1)&1
          for (int s = 0; s < 3; ++s) {
            if (((s + 1) & 1) != pw) continue;
            add_tap_entry((r == 0 ? -1 : 0) * Wp + (s == 0 ? -1 : 0));
            tds[q].push_back({r, s});
            ++cnt[q];
          }
        }
      }
      for (int q = 0; q < 4; ++q)
        for (int c = 0; c < nch; ++c) {
          const int g = q * nch + c;
          p.group_plane[g] = q * (L.cin / 8) + c * 8;
          p.group_tapofs[g] = ofs[q]; p.group_ntaps[g] = cnt[q];
          L.group_taps.push_back(tds[q]);
          L.group_cin.push_back(c * 64);
        }
      p.halo_before = Wp + 1; p.halo_after = Wp + 1;   // halo_after is not needed by these taps: kept equal to the 3x3
      break;                                          // convs' so that a stage's layers share one ring geometry (chains)
    }
    case K_DOWN1_S2: {
      p.kc8 = 8;
      p.n_groups = L.cin / 64;
      add_tap_entry(0);
      for (int g = 0; g < p.n_groups; ++g) {
        p.group_plane[g] = g * 8; p.group_tapofs[g] = 0; p.group_ntaps[g] = 1;   // parity (0,0) planes come first
        L.group_taps.push_back({{1, 1}});   // the 1x1 kernel is stored as (0,0); handled in the packer
        L.group_cin.push_back(g * 64);
      }
      p.halo_before = 0; p.halo_after = 0;
      break;
    }
    case K_STEM: {
      p.kc8 = 2;
      p.n_groups = 1;
      std::vector<TapDef> td;
      for (int t = 0; t < 16; ++t) {
        add_tap_entry((t / 4 - 2) * Wp + (t % 4 - 2));
        td.push_back({t / 4, t % 4});
      }
      p.group_plane[0] = 0; p.group_tapofs[0] = 0; p.group_ntaps[0] = 16;
      L.group_taps.push_back(td);
      L.group_cin.push_back(0);
      p.halo_before = 2 * Wp + 2; p.halo_after = Wp + 1;
      break;
    }
    case K_FC: {
      p.kc8 = 8;
      p.n_groups = L.cin / 64;
      add_tap_entry(0);
      for (int g = 0; g < p.n_groups; ++g) {
        p.group_plane[g] = g * 8; p.group_tapofs[g] = 0; p.group_ntaps[g] = 1;
        L.group_taps.push_back({{0, 0}});
        L.group_cin.push_back(g * 64);
      }
      p.halo_before = 0; p.halo_after = 0;
      break;
    }
  }
  if (L.kind != K_CONV3) p.first_group2 = p.n_groups;
  if (p.n_groups > kMaxGroups || ntap_entries > kMaxTaps) return fail(FLOPE_EINVAL, "conv plan exceeds table sizes");
  p.taps_total = 0;
  for (int g = 0; g < p.n_groups; ++g) p.taps_total += p.group_ntaps[g];

  // ---- output side ----
  p.relu = L.relu;
  p.out_mode = L.out_mode;
  p.Cout = L.cout;
  if (L.out_mode == OUT_F32_ROWS) {
    p.out = e->d_feat;
  } else {
    const ActBuf& ob = e->bufs[L.out_buf];
    p.out = ob.d;
    p.out_plane = ob.g.plane; p.out_base = ob.g.base; p.out_Hp = ob.g.Hp; p.out_Wp = ob.g.Wp;
  }
  if (L.res_buf >= 0) {
    const ActBuf& rb = e->bufs[L.res_buf];
    p.res = rb.d;
    p.res_plane = rb.g.plane; p.res_base = rb.g.base; p.res_Hp = rb.g.Hp; p.res_Wp = rb.g.Wp;
  }

  // ---- tile configuration: persistent kernel, one CTA per SM, accumulator double-buffered in TMEM ----
  L.pair = e->use_pair && L.kind != K_FC;   // fc: M = batch rows only, a handful of tiles - stays single-CTA
  L.n_tile = L.cout >= 128 ? 128 : 64;
  if (L.pair && L.cout >= 256) L.n_tile = 256; // M = 256 x N = 256 MMAs: 8 KB of operand reads per SM per 128 tensor cycles
  L.mt = 256 / L.n_tile;                       // 2 stages x 256 columns = all 512 TMEM columns
  if (L.kind == K_FC && e->fc_small) {
    // fc: M = crops only.  With 256 x 128 tiles a 256-crop batch is 16 CTAs that each stream 384 KB of operands and
    // store a 128 KB fp32 tile (17 us, a third of it epilogue); 128 x 64 tiles make it 64 short CTAs.
    L.n_tile = 64; L.mt = 1;
  }
  if (L.pair && !L.pool && L.kind != K_STEM && e->small_tiles) {
    // small batches (streaming: a handful of flowers per frame): with the throughput tiles a layer is a few work
    // items whose serial MMA chain (up to 288 K-steps at N = 256) is the whole latency.  128-position x 64-channel
    // tiles give 4-16x more items and chains of 36-288 short MMAs.
    const long long npos = (long long)e->max_batch * p.Hp * p.Wp;
    const long long items = (npos + 256 * L.mt - 1) / (256 * L.mt) * (L.cout / L.n_tile);
    if (items < e->num_sms / 4) { L.n_tile = 64; L.mt = 1; }
  }
  const int nb_rows = L.pair ? L.n_tile / 2 : L.n_tile;   // weight rows per CTA and tile
  size_t pool_smem = 0;
  int span = L.mt * 128;                       // positions of the halo tile one tile covers (without the halo)
  if (L.pool) {
    // sub-tile = one conv row (lane = column), tile = four rows, work item = a quarter of a crop
    if (p.Wp > 128 || p.H % 16) return fail(FLOPE_EINVAL, "fused stem pooling needs S/2+2 <= 128 and S % 32 == 0");
    // a work item is a band of pool_rows conv rows (a quarter of the crop; 8 rows when the batch is too small to
    // fill the pairs with quarters - a shorter serial chain per CTA at the price of one carry row per 8)
    p.pool_rows = p.H / 4;
    if (e->small_tiles && e->max_batch * 4 < e->num_sms / 2 && p.H % 8 == 0) p.pool_rows = 8;
    p.pool_split = p.H / p.pool_rows;
    span = 3 * p.Wp + 128;
    pool_smem = kPoolXchBytes;
  }
  const int halo = p.halo_before + p.halo_after;
  // latency tiles: the trunk launch keeps kLatencyTapsPerSlot weight tiles in every ring slot (trunk_chain.cuh); the
  // per-layer kernels use the first tile's worth of each slot
  const int b_tiles_per_slot = (L.pair && L.n_tile == 64 && L.mt == 1 && L.kind != K_FC) ? kLatencyTapsPerSlot : 1;
  auto smem_of = [&](int n_a, int n_b) {
    return (size_t)1024 + (size_t)kMaxChain * L.cout * sizeof(float) + (size_t)n_a * p.kc8 * (span + halo) * 16 +
           (size_t)n_b * b_tiles_per_slot * p.kc8 * nb_rows * 16 + pool_smem;
  };
  // weight tiles are consumed every MT*kc8/2 MMAs, so several must be in flight to cover L2 latency;
  // halo tiles are consumed once per group: two or three slots are enough.  Rings run across tiles.
  int best_a = 0, best_b = 0;
  for (int n_b : {8, 6, 4, 3, 2}) {
    if (p.kc8 == 2) n_b = kMaxBSlots;          // stem: 2 KB weight tiles
    int n_a = p.kc8 == 2 ? 4 : 3;
    while (n_a >= 1 && smem_of(n_a, n_b) > (size_t)kMaxSmem) --n_a;
    if (n_a >= 2 || (n_a == 1 && n_b == 2)) { best_a = n_a; best_b = n_b; break; }
  }
  if (!best_a) return fail(FLOPE_EINVAL, "conv layer " + L.name + " does not fit in shared memory");
  p.n_a_slots = best_a; p.n_b_slots = best_b;
  p.b_resident = (L.cout == L.n_tile && p.taps_total <= best_b) ? 1 : 0;   // e.g. stem: all 16 weight tiles stay in smem
  if (p.b_resident) p.n_b_slots = p.taps_total;                            // slot t <-> weight tile t, for every tile
  L.smem = smem_of(best_a, best_b);
  return FLOPE_OK;
}

int build_network(flope_engine* e) {
  const int S = e->S;
  const int s2 = S / 2, s4 = S / 4;
  e->buf_x0 = add_buf(e, 16, s2, s2, 2, false, "x0");
  e->buf_stem = add_buf(e, 64, s2, s2, 1, false, "stem");
  int cur = add_buf(e, 64, s4, s4, 1, false, "maxpool");
  e->buf_mp_out = cur;
  add_conv(e, "conv1", K_STEM, 16, 64, e->buf_x0, e->buf_stem, -1, 1, OUT_PLAIN, "base.conv1.weight", "base.bn1");
  if (s2 + 2 <= 128) {                         // fused stem + max-pool (crop side <= 252); larger crops keep two kernels
    ConvLayer L = e->layers.back();
    L.name = "conv1+maxpool"; L.out_buf = e->buf_mp_out; L.pool = true;
    e->stem_pool = L;
    e->can_fuse_pool = true;
    e->fuse_pool = true;
  }
  int C = 64, side = s4;
  for (int stage = 1; stage <= 4; ++stage) {
    const std::string ln = "base.layer" + std::to_string(stage);
    const std::string dn = "layer" + std::to_string(stage);
    int bufB = add_buf(e, C, side, side, 1, false, nullptr);
    int blk0_out;
    if (stage == 1) {
      blk0_out = add_buf(e, C, side, side, 1, false, (dn + ".0").c_str());
      add_conv(e, dn + ".0.conv1", K_CONV3, C, C, cur, bufB, -1, 1, OUT_PLAIN, ln + ".0.conv1.weight", ln + ".0.bn1");
      add_conv(e, dn + ".0.conv2", K_CONV3, C, C, bufB, blk0_out, cur, 1, OUT_PLAIN, ln + ".0.conv2.weight", ln + ".0.bn2");
    } else {
      // `cur` is the parity-split output of the previous stage, geometry = this stage's grid
      const int cin = C / 2;
      blk0_out = add_buf(e, C, side, side, 1, false, (dn + ".0").c_str());
      add_conv(e, dn + ".0.conv1", K_CONV3_S2, cin, C, cur, bufB, -1, 1, OUT_PLAIN, ln + ".0.conv1.weight", ln + ".0.bn1");
      // conv2 + the 1x1/s2 projection shortcut (downsample.0 / downsample.1) in one accumulator
      add_conv(e, dn + ".0.conv2+downsample", K_CONV3, C, C, bufB, blk0_out, -1, 1, OUT_PLAIN, ln + ".0.conv2.weight", ln + ".0.bn2");
      e->layers.back().short_buf = cur;
      e->layers.back().cin2 = cin;
      e->layers.back().wkey2 = ln + ".0.downsample.0.weight";
      e->layers.back().bnkey2 = ln + ".0.downsample.1";
    }
    int blk1_out;
    int mode;
    if (stage < 4) {
      blk1_out = add_buf(e, C, side / 2, side / 2, 1, true, (dn + ".1").c_str());
      mode = OUT_PARITY;
    } else {
      blk1_out = add_buf(e, C, side, side, 1, false, (dn + ".1").c_str());
      mode = OUT_PLAIN;
    }
    add_conv(e, dn + ".1.conv1", K_CONV3, C, C, blk0_out, bufB, -1, 1, OUT_PLAIN, ln + ".1.conv1.weight", ln + ".1.bn1");
    add_conv(e, dn + ".1.conv2", K_CONV3, C, C, bufB, blk1_out, blk0_out, 1, mode, ln + ".1.conv2.weight", ln + ".1.bn2");
    {
      const int last = (int)e->layers.size() - 1;        // the stage's four convs form one chain
      e->chains.push_back({last - 3, last - 2, last - 1, last});
    }
    cur = blk1_out;
    if (stage < 4) { C *= 2; side /= 2; }
  }
  e->buf_pool_in = cur;
  e->buf_pool = add_buf(e, 512, 1, 1, 0, false, nullptr);
  add_conv(e, "fc", K_FC, 512, e->feat_dim, e->buf_pool, -1, -1, 1, OUT_F32_ROWS, "base.fc.0.weight", "", "base.fc.0.bias");
  return FLOPE_OK;
}

void pack_weights_host(const ConvLayer& L, const float* w, const std::vector<float>& scale, const float* w2,
                       const std::vector<float>& scale2, std::vector<__nv_bfloat16>& out) {
  const ConvParams& p = L.p;
  const int NT = L.n_tile;
  const int n_tiles = L.cout / NT;
  const size_t tile_elems = (size_t)p.kc8 * NT * 8;
  // element index inside one packed weight tile: [k8][NT][8], or for the pair kernel [rank][k8][NT/2][8]
  // (each CTA of a pair copies its half of the rows with one bulk copy)
  const int half = NT / 2;
  auto at = [&](int k8, int n, int j) -> size_t {
    if (!L.pair) return ((size_t)k8 * NT + n) * 8 + j;
    return (((size_t)(n / half) * p.kc8 + k8) * half + (n % half)) * 8 + j;
  };
  out.assign((size_t)n_tiles * p.taps_total * tile_elems, __float2bfloat16_rn(0.f));
  for (int nt = 0; nt < n_tiles; ++nt) {
    size_t tile = (size_t)nt * p.taps_total;
    for (int g = 0; g < p.n_groups; ++g) {
      for (size_t t = 0; t < L.group_taps[g].size(); ++t, ++tile) {
        const TapDef td = L.group_taps[g][t];
        __nv_bfloat16* dst = out.data() + tile * tile_elems;
        for (int k8 = 0; k8 < p.kc8; ++k8)
          for (int n = 0; n < NT; ++n)
            for (int j = 0; j < 8; ++j) {
              const int co = nt * NT + n;
              float v = 0.f;
              if (g >= p.first_group2) {          // projection shortcut: (Cout, Cin2) 1x1 weights, their own BN scale
                const int ci = L.group_cin[g] + k8 * 8 + j;
                dst[at(k8, n, j)] = __float2bfloat16_rn(w2[(size_t)co * L.cin2 + ci] * scale2[co]);
                continue;
              }
              switch (L.kind) {
                case K_CONV3:
                case K_CONV3_S2: {
                  const int ci = L.group_cin[g] + k8 * 8 + j;
                  v = w[(((size_t)co * L.cin + ci) * 3 + td.r) * 3 + td.s];
                  break;
                }
                case K_DOWN1_S2:
                case K_FC: {
                  const int ci = L.group_cin[g] + k8 * 8 + j;
                  v = w[(size_t)co * L.cin + ci];
                  break;
                }
                case K_STEM: {
                  // k = by*8 + bx*4 + c ; w8 = 7x7 kernel zero-padded to 8x8 at the top/left
                  const int by = k8, bx = j >> 2, c = j & 3;
                  const int i8 = 2 * td.r + by, j8 = 2 * td.s + bx;
                  if (c < 3 && i8 >= 1 && j8 >= 1) v = w[(((size_t)co * 3 + c) * 7 + (i8 - 1)) * 7 + (j8 - 1)];
                  break;
                }
              }
              dst[at(k8, n, j)] = __float2bfloat16_rn(v * scale[co]);   // BN scale folded in before the bf16 rounding
            }
      }
    }
  }
}

struct ProfScope {                               // records a start/stop event pair around one launch (mode 1) or a chain (mode 2)
  flope_engine* e; cudaStream_t st; bool on;
  ProfScope(flope_engine* e_, const std::string& name, cudaStream_t st_, int mode = 1)
      : e(e_), st(st_), on(e_->profile && e_->profile_mode == mode) {
    if (!on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    e->prof_names.push_back(name); e->prof_ev.push_back(a); e->prof_ev.push_back(b);
    cudaEventRecord(a, st);
  }
  ~ProfScope() { if (on) cudaEventRecord(e->prof_ev.back(), st); }
};

// One launch of conv_igemm_kernel over `count` consecutive layers (a chain; count == 1 for a single layer).
constexpr int kStampLaunches = 32;              // conv launches of one forward the timeline keeps
constexpr size_t kChainHeader = 2048;           // uint32 words ahead of a chain's tile flags (claim counter + hand-over slots)
static_assert(kChainHeader / 2 >= 74 * kClaimRing, "hand-over slots for every CTA");

// Kernel arguments of `count` consecutive layers run as one chain (count == 1: a single layer).
int build_chain(flope_engine* e, ConvLayer* const* Ls, int count, int n, uint32_t* flags, ConvChain& c) {
  ConvLayer& L0 = *Ls[0];
  std::memset(&c, 0, sizeof(c));
  c.n_layers = count;
  const int TM = L0.mt * 128 * (L0.pair ? 2 : 1);           // positions per (pair) tile
  for (int i = 0; i < count; ++i) {
    ConvLayer& L = *Ls[i];
    ConvParams& p = c.L[i];
    p = L.p;
    p.n_positions = n * p.Hp * p.Wp;
    p.wgt = L.d_w; p.bias = L.d_bias;
    p.n_n_tiles = L.cout / L.n_tile;
    // work items: (pair) tiles, or for the pooled stem row bands (two per pair)
    p.n_work = L.pool ? (n * p.pool_split + (L.pair ? 1 : 0)) / (L.pair ? 2 : 1) : (p.n_positions + TM - 1) / TM * p.n_n_tiles;
    p.res_layer = -1;
    if (L.res_buf >= 0)
      for (int j = 0; j < i; ++j)
        if (Ls[j]->out_buf == L.res_buf) p.res_layer = j;  // the residual is produced inside this launch
    if (L.n_tile != L0.n_tile || L.mt != L0.mt || L.pair != L0.pair || p.kc8 != L0.p.kc8 || p.halo_before != L0.p.halo_before ||
        p.halo_after != L0.p.halo_after || p.n_a_slots != L0.p.n_a_slots || p.n_b_slots != L0.p.n_b_slots || L.cout != L0.cout ||
        p.Hp != L0.p.Hp || p.Wp != L0.p.Wp)
      return fail(FLOPE_EINVAL, "layers of a chain must share tile configuration and geometry: " + L.name);
  }
  c.n_m_tiles = c.L[0].n_work / c.L[0].n_n_tiles;
  // chain scratch: [0] claim counter, [kChainHeader/2 ..) claim hand-over slots, [kChainHeader ..) tile flags
  c.flags = count > 1 ? flags + kChainHeader : nullptr;
  c.counter = (count > 1 && e->chain_dynamic) ? flags : nullptr;
  c.claim_static = e->chain_dynamic == 2;
  c.claims = (count > 1 && e->chain_dynamic) ? flags + kChainHeader / 2 : nullptr;
  c.expected = (uint32_t)(c.L[0].n_n_tiles * (L0.pair ? 2 : 1));
  c.fail = e->d_flags ? e->d_flags + 1 : nullptr;      // one word for every chain of the engine (header of the first chain)
  c.host_err = host_err_word(true);
  return FLOPE_OK;
}

// One launch of conv_igemm_kernel over `count` consecutive layers (a chain; count == 1 for a single layer).
int run_convs(flope_engine* e, ConvLayer* const* Ls, int count, int n, cudaStream_t st, uint32_t* flags) {
  ConvLayer& L0 = *Ls[0];
  ProfScope ps(e, count == 1 ? "conv:" + L0.name : "conv:" + L0.name.substr(0, L0.name.find('.')) + " (chain of " + std::to_string(count) + ")", st);
  ConvChain c;
  int rc;
  if ((rc = build_chain(e, Ls, count, n, flags, c))) return rc;
  const int tiles = c.L[0].n_work * count;
  dim3 grid((unsigned)(L0.pair ? 2 * std::min(tiles, e->num_sms / 2) : std::min(tiles, e->num_sms)));
  if (e->d_stamps && e->stamp_launch < kStampLaunches && grid.x <= 148) c.stamps = e->d_stamps + (size_t)e->stamp_launch++ * 148 * kStampWords;
  cudaError_t ce = cudaSuccess;
  const int taps = L0.kind == K_STEM ? 16 : 0;   // the stem's 4x4 window is issued a row of taps at a time
  // bit 0: programmatic dependent launch; bit 1: cooperative launch (chains of engines that share the device)
  const int launch_mode = (e->use_pdl ? 1 : 0) | ((count > 1 && e->chain_coop) ? 2 : 0);
  if (!conv_launch(L0.n_tile, L0.mt, L0.pair, taps, c, grid, L0.smem, st, launch_mode, &ce)) return fail(FLOPE_EINVAL, "no kernel instance for " + L0.name);
  if (ce != cudaSuccess) return fail(FLOPE_ECUDA, "launch of " + L0.name + ": " + cudaGetErrorString(ce));
  ++e->launches;
  return FLOPE_OK;
}

// Shared memory of the trunk launch: every stage is laid out exactly like its per-stage chain, so the largest plan.
size_t trunk_smem(const flope_engine* e) {
  size_t m = 0;
  for (const auto& ch : e->chains) m = std::max(m, e->layers[ch[0]].smem);
  return m;
}

// layer1 .. layer4 as one launch (trunk_chain.cuh) when the plan has the shapes that kernel is instantiated for:
// returns the mask of stages on latency tiles (64x1), or -1 when the trunk kernel does not apply.
int trunk_shape_mask(const flope_engine* e) {
  if (!e->use_trunk || !e->use_chain || !e->use_pair || e->chain_dynamic || e->chain_coop || e->chains.size() != kTrunkStages) return -1;
  static const int shape[kTrunkStages][2] = {{64, 4}, {128, 2}, {256, 1}, {256, 1}};
  int mask = 0;
  for (int s = 0; s < kTrunkStages; ++s) {
    if (e->chains[s].size() != (size_t)kMaxChain) return -1;
    const ConvLayer& first = e->layers[e->chains[s][0]];
    const bool small = first.n_tile == 64 && first.mt == 1;
    if (small) mask |= 1 << s;
    for (int li : e->chains[s]) {
      const ConvLayer& L = e->layers[li];
      if (!L.pair || L.pool || L.p.kc8 != 8 || L.p.b_resident) return -1;
      if (small ? (L.n_tile != 64 || L.mt != 1) : (L.n_tile != shape[s][0] || L.mt != shape[s][1])) return -1;
    }
    if (s > 0 && first.kind != K_CONV3_S2) return -1;
  }
  return (mask == 0 || mask == 8 || mask == 12 || mask == 14 || mask == 15) ? mask : -1;
}

// Split-K for a stage on latency tiles whose layers would occupy less than half of the CTA pairs (ConvChain::k_splits):
// as many splits as keep every item of a layer in one round of the pairs, each a contiguous range of K groups with about
// the same number of weight tiles.
constexpr size_t kSplitWsPerStage = (size_t)74 * 2 * 16 * 128;          // float4s: 74 partial items of 2 x 128 rows x 64 channels
void plan_splits(const flope_engine* e, int s, ConvLayer* const* Ls, ConvChain& c) {
  c.k_splits = 1;
  if (e->trunk_splitk < 2 || !((e->trunk_split_stages >> s) & 1) || !e->d_split_ws || Ls[0]->n_tile != 64 || Ls[0]->mt != 1) return;
  const int pairs = std::min(74, e->num_sms / 2);
  // decided on the engine's max_batch, not on this call's n: the K partition - hence the rounding - of a crop's result
  // must not depend on how many crops came with it
  const int n_work = (int)(((long long)e->max_batch * c.L[0].Hp * c.L[0].Wp + 255) / 256) * c.L[0].n_n_tiles;
  int ks = std::min(std::min(kMaxSplit, e->trunk_splitk), pairs / std::max(1, n_work));
  for (int l = 0; l < c.n_layers; ++l) ks = std::min(ks, c.L[l].n_groups);
  // measured on the streaming configuration (8 crops): layer4 (16 tiles, 3-4 splits) 133.6 -> 126 us for the trunk launch;
  // layer3 (28 tiles, 2 splits) costs 4 us instead - the store / wait / reload of the partials outweighs half a K loop
  if (ks < 3 || (size_t)n_work * (ks - 1) * 2 * 16 * 128 > kSplitWsPerStage) return;
  for (int l = 0; l < c.n_layers; ++l) {
    const ConvParams& q = c.L[l];
    int g = 0, taps = 0;
    c.split_group[l][0] = 0; c.split_tap[l][0] = 0;
    for (int sp = 1; sp < ks; ++sp) {
      // advance to the group boundary nearest to sp / ks of the taps, leaving a group for every later split
      const int target = q.taps_total * sp / ks;
      do { taps += q.group_ntaps[g]; ++g; } while (taps < target && g < q.n_groups - (ks - sp));
      c.split_group[l][sp] = (unsigned char)g; c.split_tap[l][sp] = (unsigned short)taps;
    }
    c.split_group[l][ks] = (unsigned char)q.n_groups; c.split_tap[l][ks] = (unsigned short)q.taps_total;
  }
  c.k_splits = ks;
  c.split_flags = c.flags + (size_t)c.n_layers * c.n_m_tiles;
  c.split_ws = e->d_split_ws + (size_t)s * kSplitWsPerStage;
}

int run_trunk(flope_engine* e, int n, cudaStream_t st) {
  ProfScope ps(e, "conv:layer1-4 (one launch, 16 convs)", st);
  TrunkParams tp;                                 // ~11 KB of kernel arguments
  std::memset(&tp, 0, sizeof(tp));
  long long items = 0;
  int rc;
  for (int s = 0; s < kTrunkStages; ++s) {
    ConvLayer* Ls[kMaxChain];
    for (int i = 0; i < kMaxChain; ++i) Ls[i] = &e->layers[e->chains[s][i]];
    if ((rc = build_chain(e, Ls, kMaxChain, n, e->d_flags + s * e->flags_per_chain, tp.st[s]))) return rc;
    tp.tile_pos[s] = Ls[0]->mt * 256;
    plan_splits(e, s, Ls, tp.st[s]);
    items += (long long)tp.st[s].n_layers * tp.st[s].L[0].n_work * std::max(1, tp.st[s].k_splits);
  }
  const size_t smem = trunk_smem(e);
  dim3 grid((unsigned)(2 * std::min<long long>(items, e->num_sms / 2)));
  if (e->d_stamps && e->stamp_launch < kStampLaunches && grid.x <= 148) tp.stamps = e->d_stamps + (size_t)e->stamp_launch++ * 148 * kStampWords;
  cudaError_t ce = cudaErrorInvalidValue;
  const int pdl = e->use_pdl ? 1 : 0;
  switch (trunk_shape_mask(e)) {
    case 0: ce = launch_k(trunk_chain_kernel<0>, grid, dim3(kConvThreads), smem, st, 2, pdl, tp); break;
    case 8: ce = launch_k(trunk_chain_kernel<8>, grid, dim3(kConvThreads), smem, st, 2, pdl, tp); break;
    case 12: ce = launch_k(trunk_chain_kernel<12>, grid, dim3(kConvThreads), smem, st, 2, pdl, tp); break;
    case 14: ce = launch_k(trunk_chain_kernel<14>, grid, dim3(kConvThreads), smem, st, 2, pdl, tp); break;
    case 15: ce = launch_k(trunk_chain_kernel<15>, grid, dim3(kConvThreads), smem, st, 2, pdl, tp); break;
  }
  if (ce != cudaSuccess) return fail(FLOPE_ECUDA, std::string("launch of the trunk chain: ") + cudaGetErrorString(ce));
  ++e->launches;
  return FLOPE_OK;
}

int run_conv(flope_engine* e, ConvLayer& L, int n, cudaStream_t st) {
  ConvLayer* one[1] = {&L};
  return run_convs(e, one, 1, n, st, nullptr);
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  return (int)std::min<long long>(std::max<long long>(g, 1), 148LL * 16);
}

// Backbone + fc on the stem input already present in buf_x0; leaves features in d_feat.
int run_backbone_launches(flope_engine* e, int n, cudaStream_t st) {
  int rc;
  size_t li = 0;
  // mode 2: the stem .. layer4 conv_igemm launches (the trunk) as they run in production - back to back, programmatic
  // dependent launch overlapping each prologue with its predecessor's tail - between ONE pair of events
  if (e->use_chain && e->d_flags) CUDA_TRY(cudaMemsetAsync(e->d_flags, 0, e->chains.size() * e->flags_per_chain * sizeof(uint32_t), st));
  if (e->d_stamps) {
    e->stamp_launch = 0;
    CUDA_TRY(cudaMemsetAsync(e->d_stamps, 0, (size_t)kStampLaunches * 148 * kStampWords * sizeof(unsigned long long), st));
  }
  std::unique_ptr<ProfScope> trunk(new ProfScope(e, "trunk", st, 2));
  if (e->fuse_pool) {
    ++li;
    if ((rc = run_conv(e, e->stem_pool, n, st))) return rc;   // stem conv + BN + ReLU + max-pool in one kernel
  } else {
    if ((rc = run_conv(e, e->layers[li++], n, st))) return rc;   // stem
    const ActBuf& a = e->bufs[e->buf_stem];
    const ActBuf& b = e->bufs[e->buf_mp_out];
    const long long total = (long long)(a.g.C / 8) * n * b.g.H * b.g.W;
    ProfScope ps(e, "maxpool", st);
    launch_k(maxpool3x3s2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, 1, e->use_pdl, a.d, a.g, b.d, b.g, n);
    ++e->launches;
  }
  if (trunk_shape_mask(e) >= 0) {
    if ((rc = run_trunk(e, n, st))) return rc;
    li = e->layers.size() - 1;
  } else if (e->use_chain && !e->chains.empty()) {
    // one launch per stage: its four convs as a layer-major stream of tiles, tile-level dependencies through flags
    for (size_t ci = 0; ci < e->chains.size(); ++ci) {
      ConvLayer* Ls[kMaxChain];
      const int cnt = (int)e->chains[ci].size();
      for (int i = 0; i < cnt; ++i) Ls[i] = &e->layers[e->chains[ci][i]];
      if ((rc = run_convs(e, Ls, cnt, n, st, e->d_flags + ci * e->flags_per_chain))) return rc;
    }
    li = e->layers.size() - 1;
  } else {
    for (; li + 1 < e->layers.size(); ++li)
      if ((rc = run_conv(e, e->layers[li], n, st))) return rc;
  }
  trunk.reset();
  {
    const ActBuf& a = e->bufs[e->buf_pool_in];
    const ActBuf& b = e->bufs[e->buf_pool];
    const int total_warps = (a.g.C / 8) * n;
    ProfScope ps(e, "avgpool", st);
    launch_k(avgpool_kernel, dim3((total_warps + 7) / 8), dim3(256), 0, st, 1, e->use_pdl, a.d, a.g, b.d, b.g, n);
    ++e->launches;
  }
  if ((rc = run_conv(e, e->layers[li], n, st))) return rc;         // fc
  if (e->cap_stream == nullptr || st != e->cap_stream) CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

void drop_graphs(flope_engine* e) {
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

// The ~23 backbone launches of one batch size are captured once into a CUDA graph (engine-internal
// buffers only, so every kernel argument is fixed for a given n) and replayed with one launch.
int run_backbone(flope_engine* e, int n, cudaStream_t st) {
  if (take_device_timeout())
    return fail(FLOPE_ECUDA, "a tile-dependency wait timed out in an earlier launch (its CTAs were not all resident: the SMs were held "
                             "by other work for seconds); the results of that call are invalid");
  ResidencyGate gate(e->use_chain && !e->chain_coop && !e->chain_dynamic, e->device, st);
  if (!e->use_graph || e->profile) return run_backbone_launches(e, n, st);
  for (auto& g : e->graphs)
    if (g.n == n) {
      CUDA_TRY(cudaGraphLaunch(g.exec, st));
      e->launches += g.launches;
      return FLOPE_OK;
    }
  if (!e->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  const int before = e->launches;
  CUDA_TRY(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  int rc = run_backbone_launches(e, n, e->cap_stream);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
  const int launches = e->launches - before;
  e->launches = before;
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess) return fail(FLOPE_ECUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return fail(FLOPE_ECUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(ce));
  if (e->graphs.size() >= 16) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
  e->graphs.push_back({n, launches, exec});
  CUDA_TRY(cudaGraphLaunch(exec, st));
  e->launches += launches;
  return FLOPE_OK;
}

int run_head(flope_engine* e, const float* feat, const float* r9_in, const float* R_in, int n, float* r9_out,
             float* R_out, double* Ryaw_out, cudaStream_t st) {
  ProfScope ps(e, "pose_head", st);
  launch_k(pose_head_kernel, dim3(n), dim3(kHeadThreads), 0, st, 1, e->use_pdl && feat != nullptr,
           feat, e->feat_dim, (const float*)e->d_wrot, (const float*)e->d_brot, r9_in, n, r9_out, R_out, Ryaw_out, R_in);
  ++e->launches;
  CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

// streaming ROI kernels (roi_stream.cuh): persistent grid, CTAs per SM from the occupancy calculator
template <int TAPS, bool HAS_MASK, int FMT, int MAXT, int MINB>
cudaError_t roi3_launch(const Roi3Params& rp, int num_sms, int ctas_per_sm, int block, size_t smem, cudaStream_t st) {
  static int occ[64] = {};                        // per device: opt in to > 48 KB of dynamic shared memory once
  static size_t occ_smem[64] = {};
  static int occ_block[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!occ[dev] || occ_smem[dev] != smem || occ_block[dev] != block) {
    cudaError_t ce = cudaFuncSetAttribute(roi3_kernel<TAPS, HAS_MASK, FMT, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (ce != cudaSuccess) return ce;
    int n = 0;
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, roi3_kernel<TAPS, HAS_MASK, FMT, MAXT, MINB>, block, smem);
    if (ce != cudaSuccess) return ce;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    occ[dev] = n; occ_smem[dev] = smem; occ_block[dev] = block;
  }
  const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ[dev]) : occ[dev];
  const int grid = std::max(1, std::min(rp.n_items, num_sms * per_sm));
  roi3_kernel<TAPS, HAS_MASK, FMT, MAXT, MINB><<<grid, block, smem, st>>>(rp);
  return cudaGetLastError();
}

int run_roi(flope_engine* e, const uint8_t* d_frames, int n_frames, int H, int W, int64_t frame_stride,
            const uint8_t* d_masks, const int32_t* d_boxes, int n, int S, int interp, void* d_out, int out_fmt,
            cudaStream_t st) {
  RoiParams rp{};
  rp.frames = d_frames; rp.frame_stride = frame_stride; rp.masks = d_masks; rp.mask_stride = (long long)H * W;
  rp.H = H; rp.W = W; rp.boxes = d_boxes; rp.n = n; rp.S = S; rp.out_fmt = out_fmt;
  if (out_fmt == FLOPE_OUT_ENGINE) {
    rp.out = e->bufs[e->buf_x0].d;
    rp.g = e->bufs[e->buf_x0].g;
  } else {
    rp.out = d_out;
  }
  const bool has_mask = d_masks != nullptr;
  const bool lanczos = interp == FLOPE_INTERP_LANCZOS4;
  ProfScope ps(e, lanczos ? "roi_crop:lanczos4" : "roi_crop:linear", st);
  // ---- streaming kernels: 16-byte phase of a row segment independent of the row, one output column per thread ----
  const int cw = (S <= 256 && S % 32 == 0) ? S : (S <= 512 && S % 64 == 0) ? S / 2 : 0;
  if (e->roi_stream && cw && W % 16 == 0 && n_frames >= 1 && n >= 1 && (out_fmt != FLOPE_OUT_ENGINE || rp.g.plane * 16 < (1LL << 31))) {
    Roi3Params q{};
    q.frames = d_frames; q.frame_stride = frame_stride; q.masks = d_masks; q.mask_stride = (long long)H * W; q.W = W;
    q.boxes = d_boxes; q.n = n; q.S = S; q.out_fmt = out_fmt; q.out = rp.out; q.g = rp.g;
    q.cols_per_item = cw;
    q.col_blocks = S / cw;
    q.rows_per_item = std::max(1, std::min(std::min(kR3MaxItemRows, S), lanczos ? e->roi_item_rows8 : e->roi_item_rows));
    if (e->roi_item_auto) {
      // small launches (a streamed frame with a few flowers, one 256-crop step): halve the items until every resident CTA
      // has about three to claim - an item re-reads 1 (Lanczos4: 7) source rows of its neighbour, hence the floors
      const long long want = 3LL * e->num_sms * (lanczos ? 3 : 4);
      const int floor_rows = lanczos ? 32 : e->roi_item_floor;
      while (q.rows_per_item / 2 >= floor_rows && (long long)n * ((S + q.rows_per_item - 1) / q.rows_per_item) * q.col_blocks < want)
        q.rows_per_item /= 2;
    }
    q.items_per_crop = (S + q.rows_per_item - 1) / q.rows_per_item * q.col_blocks;
    const long long n_items = (long long)n * q.items_per_crop;
    q.n_items = (int)n_items;
    const int side = std::min(H, W);                 // a square in-frame box is at most this wide
    const int max_pitch = ((15 + 3 * side + 15) & ~15) + ((15 + side + 15) & ~15) + (lanczos ? 2 * kR3PadM + kR3PadL + kR3PadR : 0);
    q.stage_bytes = (std::max(e->roi_stage_kb * 1024, (lanczos ? 1 : 2) * max_pitch) + 127) & ~127;
    q.n_stages = std::max(2, std::min(kR3MaxStages, e->roi_stages));
    q.sched = e->roi_dynamic ? e->d_roi_sched : nullptr;
    if (lanczos && e->roi_axis_tab) {
      // the Lanczos4 coefficient tables of every crop and axis, computed by the whole GPU instead of the producer warps
      const size_t words = (size_t)n * 2 * S * 8;
      if (words > e->roi_tab_words) {
        if (e->d_roi_tab) CUDA_TRY(cudaFree(e->d_roi_tab));          // (synchronises: nothing in flight uses the old table)
        e->d_roi_tab = nullptr; e->roi_tab_words = 0;
        CUDA_TRY(cudaMalloc(&e->d_roi_tab, words * sizeof(uint32_t)));
        e->roi_tab_words = words;
      }
      const long long total = 2LL * n * S;
      roi3_axis_tables_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_boxes, n, S, e->d_roi_tab);
      ++e->launches;
      q.axis_tab = e->d_roi_tab;
    }
    q.xtab_slot = cw * (lanczos ? kR3XtabEntry8 : kR3XtabEntry2);
    q.ring_off = (r3_xtab_off(lanczos ? 8 : 2) + 2 * q.xtab_slot + 127) & ~127;
    const size_t smem = (size_t)q.ring_off + (size_t)q.n_stages * q.stage_bytes + kR3RingTail;
    if (smem <= (size_t)kMaxSmem && n_items < (1LL << 30)) {
      const int block = cw + 32;
      const int key = (lanczos ? 4 : 0) | (has_mask ? 2 : 0) | (out_fmt == FLOPE_OUT_ENGINE ? 1 : 0);
      cudaError_t ce;
      // up to 224 columns per item: 256 threads, four CTAs per SM (<= 64 registers; Lanczos4: three); wider items: 288 threads
#define ROI3_CASE(K, T, M, F) case K: ce = block <= 256 ? roi3_launch<T, M, F, 256, (T == 2 ? 4 : 3)>(q, e->num_sms, e->roi_ctas_per_sm, block, smem, st) \
                                                       : roi3_launch<T, M, F, 288, (T == 2 ? 4 : 3)>(q, e->num_sms, e->roi_ctas_per_sm, block, smem, st); break;
      switch (key) {
        ROI3_CASE(0, 2, false, 0) ROI3_CASE(1, 2, false, 1) ROI3_CASE(2, 2, true, 0) ROI3_CASE(3, 2, true, 1)
        ROI3_CASE(4, 8, false, 0) ROI3_CASE(5, 8, false, 1) ROI3_CASE(6, 8, true, 0)
        ROI3_CASE(7, 8, true, 1)
        default: ce = cudaErrorInvalidValue; break;
      }
#undef ROI3_CASE
      if (ce != cudaSuccess) return fail(FLOPE_ECUDA, std::string("ROI kernel launch: ") + cudaGetErrorString(ce));
      ++e->launches;
      return FLOPE_OK;
    }
  }
  // ---- generic fallback: one thread per output column, taps read from global memory ----
  rp.rows_per_strip = S >= 448 ? 128 : 112;
  const int block = S >= 256 ? 256 : ((S + 31) / 32 * 32);
  dim3 grid((S + block - 1) / block, (S + rp.rows_per_strip - 1) / rp.rows_per_strip, n);
  if (lanczos) {
    if (has_mask) roi_crop_kernel<8, true><<<grid, block, 0, st>>>(rp);
    else roi_crop_kernel<8, false><<<grid, block, 0, st>>>(rp);
  } else {
    if (has_mask) roi_crop_kernel<2, true><<<grid, block, 0, st>>>(rp);
    else roi_crop_kernel<2, false><<<grid, block, 0, st>>>(rp);
  }
  ++e->launches;
  CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

__global__ void normalise_lut_kernel(float* out) {
  const int i = threadIdx.x, m = blockIdx.x;
  out[m * 256 + i] = normalise_u8(i, m);
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int flope_version(void) { return 100; }
const char* flope_last_error(void) { return g_err.c_str(); }

int flope_engine_create(flope_engine** out, int device, int max_batch, int crop_hw) {
  if (!out) return fail(FLOPE_EINVAL, "out is NULL");
  *out = nullptr;
  if (max_batch < 1 || crop_hw < 32 || crop_hw % 32 != 0 || crop_hw > 1024)
    return fail(FLOPE_EINVAL, "max_batch must be >= 1 and crop_hw a multiple of 32 in [32,1024]");
  if ((long long)max_batch * crop_hw * crop_hw >= (1LL << 32))
    return fail(FLOPE_EINVAL, "max_batch * crop_hw^2 must stay below 2^32 (32-bit pixel indices in the ingest kernel)");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(FLOPE_EINVAL, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  // the library carries sm_100a code only, and arch-specific targets do not run on other 10.x parts
  if (prop.major != 10 || prop.minor != 0) return fail(FLOPE_ECUDA, "flope_b200 kernels are built for sm_100a (B200, compute capability 10.0) only");
  if (prop.multiProcessorCount > 148) return fail(FLOPE_ECUDA, "more than 148 SMs: the claim slots and stamp buffers are sized for B200");
  CUDA_TRY(conv_set_all_attrs());
  flope_engine* e = new flope_engine();
  e->device = device; e->max_batch = max_batch; e->S = crop_hw;
  e->num_sms = prop.multiProcessorCount;
  int rc = build_network(e);
  // from here on every early return (CUDA_TRY included) releases what has been allocated so far
  std::unique_ptr<flope_engine, void (*)(flope_engine*)> guard(e, flope_engine_destroy);
  if (rc) return rc;
  for (ActBuf& b : e->bufs) {
    if (cudaMalloc(&b.d, b.bytes()) != cudaSuccess) { cudaGetLastError(); return fail(FLOPE_ENOMEM, "activation buffer allocation failed"); }
    CUDA_TRY(cudaMemset(b.d, 0, b.bytes()));
  }
  CUDA_TRY(cudaMalloc(&e->d_feat, (size_t)(e->max_batch + kMaxTM) * e->feat_dim * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_wrot, (size_t)9 * e->feat_dim * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_brot, 9 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_roi_sched, 2 * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(e->d_roi_sched, 0, 2 * sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&e->d_r9, (size_t)e->max_batch * 9 * sizeof(float)));
  for (ConvLayer& L : e->layers)
    if ((rc = plan_conv(e, L))) return rc;
  if (e->can_fuse_pool && (rc = plan_conv(e, e->stem_pool))) return rc;
  {
    // completion counters of the chains: kMaxChain layers x position tiles at max_batch (the smallest tile is 128 positions)
    size_t per = 0;
    for (const auto& ch : e->chains) {
      const ConvParams& p = e->layers[ch[0]].p;
      per = std::max(per, kChainHeader + (size_t)kMaxChain * ((size_t)e->max_batch * p.Hp * p.Wp / 128 + 2) + (size_t)kMaxChain * 128);   // + split-K counters
    }
    e->flags_per_chain = per;
    if (per) CUDA_TRY(cudaMalloc(&e->d_flags, e->chains.size() * per * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&e->d_split_ws, kTrunkStages * kSplitWsPerStage * sizeof(float4)));     // 19 MB: split-K partials of the latency tiles
  }
  CUDA_TRY(cudaDeviceSynchronize());
  *out = guard.release();
  return FLOPE_OK;
}

void flope_engine_destroy(flope_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  drop_graphs(e);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (ActBuf& b : e->bufs) cudaFree(b.d);
  for (ConvLayer& L : e->layers) { cudaFree(L.d_w); cudaFree(L.d_bias); }
  cudaFree(e->stem_pool.d_w); cudaFree(e->stem_pool.d_bias);
  cudaFree(e->d_flags);
  cudaFree(e->d_split_ws);
  cudaFree(e->d_roi_sched);
  cudaFree(e->d_roi_tab);
  cudaFree(e->d_stamps);
  cudaFree(e->d_feat); cudaFree(e->d_wrot); cudaFree(e->d_brot); cudaFree(e->d_r9);
  delete e;
}

int flope_engine_load_weights(flope_engine* e, const flope_tensor_desc* tensors, int n) {
  if (!e || !tensors) return fail(FLOPE_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(e->device));
  drop_graphs(e);                              // captured kernel arguments point at the old weight buffers
  std::map<std::string, const flope_tensor_desc*> sd;
  for (int i = 0; i < n; ++i)
    if (tensors[i].name && tensors[i].data) sd[tensors[i].name] = &tensors[i];
  auto need = [&](const std::string& k, int64_t numel) -> const float* {
    auto it = sd.find(k);
    if (it == sd.end()) return nullptr;
    int64_t ne = 1;
    for (int d = 0; d < it->second->ndim; ++d) ne *= it->second->shape[d];
    return ne == numel ? it->second->data : nullptr;
  };
  std::vector<ConvLayer*> all;
  for (ConvLayer& L : e->layers) all.push_back(&L);
  if (e->can_fuse_pool) all.push_back(&e->stem_pool);
  for (ConvLayer* Lp : all) {
    ConvLayer& L = *Lp;
    int64_t wn = 0;
    switch (L.kind) {
      case K_CONV3: case K_CONV3_S2: wn = (int64_t)L.cout * L.cin * 9; break;
      case K_DOWN1_S2: case K_FC: wn = (int64_t)L.cout * L.cin; break;
      case K_STEM: wn = (int64_t)L.cout * 3 * 49; break;
    }
    const float* w = need(L.wkey, wn);
    if (!w) return fail(FLOPE_EINVAL, "state_dict entry missing or mis-shaped: " + L.wkey);
    std::vector<float> scale(L.cout, 1.f), bias(L.cout, 0.f);
    if (!L.bnkey.empty()) {
      const float* gw = need(L.bnkey + ".weight", L.cout);
      const float* gb = need(L.bnkey + ".bias", L.cout);
      const float* mu = need(L.bnkey + ".running_mean", L.cout);
      const float* var = need(L.bnkey + ".running_var", L.cout);
      if (!gw || !gb || !mu || !var) return fail(FLOPE_EINVAL, "BatchNorm entries missing for " + L.bnkey);
      for (int c = 0; c < L.cout; ++c) {     // eval-mode fold, eps = 1e-5 (torchvision BatchNorm2d default)
        const float s = gw[c] / std::sqrt(var[c] + 1e-5f);
        scale[c] = s;
        bias[c] = gb[c] - mu[c] * s;
      }
    } else {
      const float* b = need(L.biaskey, L.cout);
      if (!b) return fail(FLOPE_EINVAL, "bias missing: " + L.biaskey);
      for (int c = 0; c < L.cout; ++c) bias[c] = b[c];
    }
    const float* w2 = nullptr;
    std::vector<float> scale2(L.cout, 1.f);
    if (L.short_buf >= 0) {
      w2 = need(L.wkey2, (int64_t)L.cout * L.cin2);
      const float* gw = need(L.bnkey2 + ".weight", L.cout);
      const float* gb = need(L.bnkey2 + ".bias", L.cout);
      const float* mu = need(L.bnkey2 + ".running_mean", L.cout);
      const float* var = need(L.bnkey2 + ".running_var", L.cout);
      if (!w2 || !gw || !gb || !mu || !var) return fail(FLOPE_EINVAL, "projection shortcut entries missing: " + L.wkey2);
      for (int c = 0; c < L.cout; ++c) {
        const float s2 = gw[c] / std::sqrt(var[c] + 1e-5f);
        scale2[c] = s2;
        bias[c] += gb[c] - mu[c] * s2;        // one fp32 bias for the summed branches
      }
    }
    std::vector<__nv_bfloat16> packed;
    pack_weights_host(L, w, scale, w2, scale2, packed);
    cudaFree(L.d_w); cudaFree(L.d_bias);
    L.d_w = nullptr; L.d_bias = nullptr;
    CUDA_TRY(cudaMalloc(&L.d_w, packed.size() * sizeof(__nv_bfloat16)));
    CUDA_TRY(cudaMalloc(&L.d_bias, L.cout * sizeof(float)));
    CUDA_TRY(cudaMemcpy(L.d_w, packed.data(), packed.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(L.d_bias, bias.data(), L.cout * sizeof(float), cudaMemcpyHostToDevice));
  }
  const float* wr = need("fc_rot.weight", (int64_t)9 * e->feat_dim);
  const float* br = need("fc_rot.bias", 9);
  if (!wr || !br) return fail(FLOPE_EINVAL, "fc_rot entries missing");
  CUDA_TRY(cudaMemcpy(e->d_wrot, wr, (size_t)9 * e->feat_dim * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(e->d_brot, br, 9 * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaDeviceSynchronize());
  e->weights_loaded = true;
  return FLOPE_OK;
}

int flope_squarify_filter(const int32_t* boxes, int n, int H, int W, int32_t* out_sq, uint8_t* keep) {
  if (n < 0 || (n > 0 && (!boxes || !out_sq || !keep))) return fail(FLOPE_EINVAL, "NULL argument");
  for (int i = 0; i < n; ++i) {
    // squarify_bb (mvg.py:324-343) on integers: the short side grows by diff, the extra pixel of an
    // odd diff going to the min side; int() truncation is exact because every operand is integral
    int64_t xmin = boxes[4 * i], ymin = boxes[4 * i + 1], xmax = boxes[4 * i + 2], ymax = boxes[4 * i + 3];
    const int64_t xr = xmax - xmin, yr = ymax - ymin;
    const int64_t diff = xr > yr ? xr - yr : yr - xr;
    const int64_t lo = (diff % 2 == 0) ? diff / 2 : (diff + 1) / 2;
    const int64_t hi = (diff % 2 == 0) ? diff / 2 : (diff - 1) / 2;
    if (xr > yr) { ymin -= lo; ymax += hi; }
    else if (xr < yr) { xmin -= lo; xmax += hi; }
    out_sq[4 * i] = (int32_t)xmin; out_sq[4 * i + 1] = (int32_t)ymin;
    out_sq[4 * i + 2] = (int32_t)xmax; out_sq[4 * i + 3] = (int32_t)ymax;
    keep[i] = !(xmin < 0 || ymin < 0 || xmax > W || ymax > H);   // bb_in_frame (mvg.py:345-351)
  }
  return FLOPE_OK;
}

int flope_pack_boxes(const uint8_t* img, int H, int W, int ch, const int32_t* boxes, int n, int slot_h, int slot_w, uint8_t* out) {
  if (n < 0 || ch < 1 || (n > 0 && (!img || !boxes || !out))) return fail(FLOPE_EINVAL, "bad argument");
  for (int i = 0; i < n; ++i) {
    const int x0 = boxes[4 * i], y0 = boxes[4 * i + 1], x1 = boxes[4 * i + 2], y1 = boxes[4 * i + 3];
    if (x0 < 0 || y0 < 0 || x1 > W || y1 > H || x1 <= x0 || y1 <= y0 || x1 - x0 > slot_w || y1 - y0 > slot_h)
      return fail(FLOPE_EINVAL, "box outside the image or larger than the slot");
    const size_t row = (size_t)(x1 - x0) * ch;
    uint8_t* dst = out + (size_t)i * slot_h * slot_w * ch;
    for (int y = y0; y < y1; ++y) std::memcpy(dst + (size_t)(y - y0) * slot_w * ch, img + ((size_t)y * W + x0) * ch, row);
  }
  return FLOPE_OK;
}

int flope_roi_crop(flope_engine* e, const uint8_t* d_frames, int n_frames, int H, int W, int64_t frame_stride,
                   const uint8_t* d_masks, const int32_t* d_boxes, int n, int S, int interp, void* d_out, int out_fmt,
                   void* stream) {
  if (!e || !d_frames || !d_boxes) return fail(FLOPE_EINVAL, "NULL argument");
  if (n == 0) return FLOPE_OK;
  if (n < 0 || S < 2 || S % 2) return fail(FLOPE_EINVAL, "bad n or S");
  if (interp != FLOPE_INTERP_LINEAR && interp != FLOPE_INTERP_LANCZOS4) return fail(FLOPE_EINVAL, "bad interp");
  if (out_fmt == FLOPE_OUT_ENGINE) {
    if (S != e->S || n > e->max_batch) return fail(FLOPE_EINVAL, "engine-format crops need S == crop_hw and n <= max_batch");
  } else if (out_fmt != FLOPE_OUT_F32_NCHW || !d_out) {
    return fail(FLOPE_EINVAL, "bad out_fmt / d_out");
  }
  CUDA_TRY(cudaSetDevice(e->device));
  e->launches = 0;
  return run_roi(e, d_frames, n_frames, H, W, frame_stride, d_masks, d_boxes, n, S, interp, d_out, out_fmt,
                 (cudaStream_t)stream);
}

int flope_ingest_crops(flope_engine* e, const float* d_in, int n, void* stream) {
  if (!e || !d_in) return fail(FLOPE_EINVAL, "NULL argument");
  if (n < 0 || n > e->max_batch) return fail(FLOPE_EINVAL, "bad n (0 .. max_batch)");
  if (n == 0) return FLOPE_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const ActBuf& x0 = e->bufs[e->buf_x0];
  const long long total = (long long)n * e->S * (e->S / 4);
  e->launches = 0;
  {
    ProfScope ps(e, "ingest", st);
    ingest_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, st>>>(d_in, n, e->S, x0.d, x0.g);
    ++e->launches;
  }
  CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

int flope_posenet_forward(flope_engine* e, const float* d_in, int n, float* d_r9, void* stream) {
  if (!e || !d_r9) return fail(FLOPE_EINVAL, "NULL argument");
  if (!e->weights_loaded) return fail(FLOPE_ESTATE, "weights not loaded");
  if (n == 0) return FLOPE_OK;
  if (n < 0 || (!d_in && n > e->max_batch)) return fail(FLOPE_EINVAL, "bad n");
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  e->launches = 0;
  for (int done = 0; done < n; done += e->max_batch) {
    const int nb = std::min(e->max_batch, n - done);
    if (d_in) {
      const ActBuf& x0 = e->bufs[e->buf_x0];
      const long long total = (long long)nb * e->S * (e->S / 4);
      ProfScope ps(e, "ingest", st);
      ingest_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, st>>>(d_in + (size_t)done * 3 * e->S * e->S, nb, e->S, x0.d, x0.g);
      ++e->launches;
    }
    int rc = run_backbone(e, nb, st);
    if (rc) return rc;
    if ((rc = run_head(e, e->d_feat, nullptr, nullptr, nb, d_r9 + (size_t)done * 9, nullptr, nullptr, st))) return rc;
  }
  return FLOPE_OK;
}

int flope_pose_head(flope_engine* e, const float* d_r9, int n, float* d_R, double* d_R_yaw, void* stream) {
  if (!e || !d_r9) return fail(FLOPE_EINVAL, "NULL argument");
  if (n == 0) return FLOPE_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  e->launches = 0;
  return run_head(e, nullptr, d_r9, nullptr, n, nullptr, d_R, d_R_yaw, (cudaStream_t)stream);
}

int flope_nullify_yaw(flope_engine* e, const float* d_R_in, int n, double* d_R_yaw, void* stream) {
  if (!e || !d_R_in || !d_R_yaw) return fail(FLOPE_EINVAL, "NULL argument");
  if (n == 0) return FLOPE_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  e->launches = 0;
  return run_head(e, nullptr, nullptr, d_R_in, n, nullptr, nullptr, d_R_yaw, (cudaStream_t)stream);
}

int flope_infer_frames(flope_engine* e, const uint8_t* d_frames, int n_frames, int H, int W, int64_t frame_stride,
                       const uint8_t* d_masks, const int32_t* d_boxes, int n, int interp, float* d_r9, float* d_R,
                       double* d_R_yaw, void* stream) {
  if (!e || !d_frames || !d_boxes) return fail(FLOPE_EINVAL, "NULL argument");
  if (!e->weights_loaded) return fail(FLOPE_ESTATE, "weights not loaded");
  if (n == 0) return FLOPE_OK;
  if (n < 0) return fail(FLOPE_EINVAL, "bad n");
  if (interp != FLOPE_INTERP_LINEAR && interp != FLOPE_INTERP_LANCZOS4) return fail(FLOPE_EINVAL, "bad interp");
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  e->launches = 0;
  for (int done = 0; done < n; done += e->max_batch) {
    const int nb = std::min(e->max_batch, n - done);
    int rc = run_roi(e, d_frames, n_frames, H, W, frame_stride, d_masks, d_boxes + (size_t)done * 5, nb, e->S, interp,
                     nullptr, FLOPE_OUT_ENGINE, st);
    if (rc) return rc;
    if ((rc = run_backbone(e, nb, st))) return rc;
    if ((rc = run_head(e, e->d_feat, nullptr, nullptr, nb, d_r9 ? d_r9 + (size_t)done * 9 : nullptr,
                       d_R ? d_R + (size_t)done * 9 : nullptr, d_R_yaw ? d_R_yaw + (size_t)done * 9 : nullptr, st)))
      return rc;
  }
  return FLOPE_OK;
}

int flope_depth_values(int device, const void* d_depth, int depth_dtype, float depth_div, const uint8_t* d_mask, int H, int W,
                       const int32_t* d_boxes, int n, float near_plane, float far_plane, int erode_k, uint8_t* d_scratch,
                       double* d_val, int32_t* d_count, void* stream) {
  if (!d_depth || !d_mask || !d_scratch) return fail(FLOPE_EINVAL, "NULL argument");
  if (n < 0 || (n > 0 && (!d_boxes || !d_val || !d_count))) return fail(FLOPE_EINVAL, "NULL argument");
  if (H < 1 || W < 1 || erode_k < 1 || erode_k > kErodeMaxK) return fail(FLOPE_EINVAL, "bad frame size or erosion size (1..31)");
  if (depth_dtype != 0 && depth_dtype != 1) return fail(FLOPE_EINVAL, "depth_dtype must be 0 (float32 metres) or 1 (uint16 raw)");
  if (depth_dtype == 1 && !(depth_div > 0.f)) return fail(FLOPE_EINVAL, "depth_div must be positive");
  CUDA_TRY(cudaSetDevice(device));
  DepthParams dp{};
  dp.depth = d_depth; dp.dtype = depth_dtype; dp.div = depth_div; dp.mask = d_mask; dp.H = H; dp.W = W;
  dp.near_plane = near_plane; dp.far_plane = far_plane; dp.k = erode_k; dp.eroded = d_scratch;
  {
    // cv2.getStructuringElement(MORPH_ELLIPSE, (k,k)): row i is ones on [c - dx, c + dx + 1) with
    // dx = round_half_even(c * sqrt((r*r - dy*dy) / (r*r))), dy = i - r, r = c = k/2 (checked against cv2 in the tests)
    const int r = erode_k / 2, c = erode_k / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < erode_k; ++i) {
      const int dy = i - r;
      int j1 = 0, j2 = 0;
      if (std::abs(dy) <= r) {
        const int dx = (int)std::nearbyint(c * std::sqrt((r * r - dy * dy) * inv_r2));
        j1 = std::max(c - dx, 0);
        j2 = std::min(c + dx + 1, erode_k);
      }
      dp.j1[i] = j1; dp.j2[i] = j2;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((W + kErodeTileW - 1) / kErodeTileW, (H + kErodeTileH - 1) / kErodeTileH);
  erode_valid_kernel<<<grid, 256, 0, st>>>(dp);
  if (n > 0) {
    // (sum, count) accumulators of the row-stripe CTAs: n x 16 bytes from the stream-ordered pool
    unsigned long long* acc = nullptr;
    CUDA_TRY(cudaMallocAsync(&acc, (size_t)n * 2 * sizeof(unsigned long long), st));
    if (cudaError_t ce = cudaMemsetAsync(acc, 0, (size_t)n * 2 * sizeof(unsigned long long), st)) {
      cudaFreeAsync(acc, st);
      return fail(FLOPE_ECUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(ce));
    }
    box_depth_kernel<<<dim3(n, kBoxSplit), 256, 0, st>>>(dp, d_boxes, acc);
    box_depth_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(acc, n, d_val, d_count);
    CUDA_TRY(cudaFreeAsync(acc, st));
  }
  CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

int flope_yolo_mask(int device, const float* d_masks, int n, int h, int w, uint8_t* d_small, uint8_t* d_out, int H, int W,
                    void* d_tables, void* stream) {
  if (!d_small || !d_out || !d_tables || (n > 0 && !d_masks)) return fail(FLOPE_EINVAL, "NULL argument");
  if (n < 0 || h < 1 || w < 1 || H < 1 || W < 1) return fail(FLOPE_EINVAL, "bad sizes");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  // d_tables: (W + H) int32 source indices followed by (W + H) x 2 int16 coefficients
  int* xofs = reinterpret_cast<int*>(d_tables);
  int* yofs = xofs + W;
  short* xcoef = reinterpret_cast<short*>(yofs + H);
  short* ycoef = xcoef + 2 * (size_t)W;
  const long long hw = (long long)h * w;
  merge_instance_masks_kernel<<<grid_for(hw, 256), 256, 0, st>>>(d_masks, n, hw, d_small);
  linear_table_kernel<<<(W + 127) / 128, 128, 0, st>>>(W, w, 0, xofs, xcoef);
  linear_table_kernel<<<(H + 127) / 128, 128, 0, st>>>(H, h, 1, yofs, ycoef);
  dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8);
  resize_linear_u8_kernel<<<grid, blk, 0, st>>>(d_small, h, w, d_out, H, W, xofs, xcoef, yofs, ycoef);
  CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

int flope_engine_last_launches(const flope_engine* e) { return e ? e->launches : 0; }

int flope_engine_profile(flope_engine* e, int enable) {
  if (!e) return fail(FLOPE_EINVAL, "NULL argument");
  for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
  e->prof_ev.clear(); e->prof_names.clear();
  e->profile = enable != 0;
  e->profile_mode = enable;
  return FLOPE_OK;
}

int flope_engine_profile_read(flope_engine* e, char* names, int names_len, float* ms, int max_entries) {
  if (!e || !names || !ms) return fail(FLOPE_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const int n = (int)e->prof_names.size();
  if (n > max_entries) return fail(FLOPE_EINVAL, "profile buffer too small");
  std::string joined;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, e->prof_ev[2 * i], e->prof_ev[2 * i + 1]));
    ms[i] = t;
    joined += e->prof_names[i];
    joined += '\n';
  }
  if ((int)joined.size() + 1 > names_len) return fail(FLOPE_EINVAL, "names buffer too small");
  std::memcpy(names, joined.c_str(), joined.size() + 1);
  return n;
}

int64_t flope_debug_activation(flope_engine* e, const char* name, int n, float* d_out, void* stream) {
  if (!e || !name || !d_out) return fail(FLOPE_EINVAL, "NULL argument");
  auto it = e->act_names.find(name);
  if (it == e->act_names.end()) return fail(FLOPE_EINVAL, std::string("unknown activation ") + name);
  const ActBuf& b = e->bufs[it->second];
  const int H = b.parity ? 2 * b.g.H : b.g.H, W = b.parity ? 2 * b.g.W : b.g.W;
  const long long total = (long long)n * b.g.C * H * W;
  unpack_to_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(b.d, b.g, b.parity ? 1 : 0, n, d_out);
  if (cudaGetLastError() != cudaSuccess) return fail(FLOPE_ECUDA, "unpack launch failed");
  return (int64_t)b.g.C * H * W;
}

int flope_debug_normalise_lut(float* d_out, void* stream) {
  if (!d_out) return fail(FLOPE_EINVAL, "NULL argument");
  normalise_lut_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(d_out);
  CUDA_TRY(cudaGetLastError());
  return FLOPE_OK;
}

int flope_debug_timeline(flope_engine* e, unsigned long long* out, int max_launches) {
  if (!e || !out) return fail(FLOPE_EINVAL, "null argument");
  if (!e->d_stamps) return fail(FLOPE_EINVAL, "timeline is off: flope_debug_set(e, \"timeline\", 1) first");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const int n = std::min(std::min(max_launches, e->stamp_launch), kStampLaunches);
  CUDA_TRY(cudaMemcpy(out, e->d_stamps, (size_t)n * 148 * kStampWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return n;
}

int flope_engine_set_schedule(flope_engine* e, int schedule) {
  if (!e) return fail(FLOPE_EINVAL, "NULL argument");
  if (schedule < FLOPE_SCHED_PERSISTENT || schedule > FLOPE_SCHED_COOPERATIVE) return fail(FLOPE_EINVAL, "unknown schedule");
  e->use_chain = schedule != FLOPE_SCHED_PER_LAYER;
  e->use_trunk = schedule == FLOPE_SCHED_PERSISTENT;
  e->chain_dynamic = schedule == FLOPE_SCHED_DYNAMIC ? 1 : 0;
  e->chain_coop = schedule == FLOPE_SCHED_COOPERATIVE;
  drop_graphs(e);
  return FLOPE_OK;
}

int flope_debug_set(flope_engine* e, const char* key, int value) {
  if (!e || !key) return fail(FLOPE_EINVAL, "NULL argument");
  if (!std::strcmp(key, "use_graph")) { e->use_graph = value != 0; return FLOPE_OK; }
  if (!std::strcmp(key, "trunk_splitk")) { e->trunk_splitk = value == 1 ? kMaxSplit : value; drop_graphs(e); return FLOPE_OK; }   // 0: off, 1: default
  if (!std::strcmp(key, "trunk_split_stages")) { e->trunk_split_stages = value; drop_graphs(e); return FLOPE_OK; }
  if (!std::strcmp(key, "roi_stream")) { e->roi_stream = value; return FLOPE_OK; }
  if (!std::strcmp(key, "roi_item_auto")) { e->roi_item_auto = value; return FLOPE_OK; }
  if (!std::strcmp(key, "roi_item_floor")) { e->roi_item_floor = std::max(1, value); return FLOPE_OK; }
  if (!std::strcmp(key, "roi_item_rows") || !std::strcmp(key, "roi_item_rows8")) {
    if (value < 1 || value > kR3MaxItemRows) return fail(FLOPE_EINVAL, "roi_item_rows must be in [1,128]");
    (key[13] ? e->roi_item_rows8 : e->roi_item_rows) = value;
    e->roi_item_auto = 0;                       // an explicit size is taken as given
    return FLOPE_OK;
  }
  if (!std::strcmp(key, "roi_stage_kb")) {
    if (value < 1 || value > 64) return fail(FLOPE_EINVAL, "roi_stage_kb must be in [1,64]");
    e->roi_stage_kb = value;
    return FLOPE_OK;
  }
  if (!std::strcmp(key, "roi_stages")) {
    if (value < 2 || value > kR3MaxStages) return fail(FLOPE_EINVAL, "roi_stages must be in [2,8]");
    e->roi_stages = value;
    return FLOPE_OK;
  }
  if (!std::strcmp(key, "roi_axis_tab")) { e->roi_axis_tab = value; return FLOPE_OK; }
  if (!std::strcmp(key, "roi_dynamic")) { e->roi_dynamic = value; return FLOPE_OK; }
  if (!std::strcmp(key, "roi_ctas_per_sm")) { e->roi_ctas_per_sm = value; return FLOPE_OK; }
  if (!std::strcmp(key, "chain_coop")) { e->chain_coop = value != 0; drop_graphs(e); return FLOPE_OK; }
  if (!std::strcmp(key, "timeline")) {
    drop_graphs(e);
    if (value && !e->d_stamps) CUDA_TRY(cudaMalloc(&e->d_stamps, (size_t)kStampLaunches * 148 * kStampWords * sizeof(unsigned long long)));
    if (!value && e->d_stamps) { cudaFree(e->d_stamps); e->d_stamps = nullptr; }
    return FLOPE_OK;
  }
  if (!std::strcmp(key, "trunk")) { e->use_trunk = value != 0; drop_graphs(e); return FLOPE_OK; }
  if (!std::strcmp(key, "chain_dynamic")) { e->chain_dynamic = value; drop_graphs(e); return FLOPE_OK; }
  if (!std::strcmp(key, "chain")) { e->use_chain = value != 0; drop_graphs(e); return FLOPE_OK; }
  if (!std::strcmp(key, "pdl")) { e->use_pdl = value != 0; drop_graphs(e); return FLOPE_OK; }
  if (!std::strcmp(key, "pair") || !std::strcmp(key, "small_tiles") || !std::strcmp(key, "fc_small")) {   // re-plans every layer; the packed weights depend on it: reload them
    if (key[0] == 'p') e->use_pair = value != 0; else if (key[0] == 's') e->small_tiles = value != 0; else e->fc_small = value != 0;
    drop_graphs(e);
    e->weights_loaded = false;
    for (ConvLayer& L : e->layers) { int rc = plan_conv(e, L); if (rc) return rc; }
    if (e->can_fuse_pool) { int rc = plan_conv(e, e->stem_pool); if (rc) return rc; }
    return FLOPE_OK;
  }
  if (!std::strcmp(key, "fuse_pool")) {
    if (value && !e->can_fuse_pool) return fail(FLOPE_EINVAL, "fused stem pooling supports crop sides up to 252");
    e->fuse_pool = value != 0;
    drop_graphs(e);
    return FLOPE_OK;
  }
  return fail(FLOPE_EINVAL, std::string("unknown debug key ") + key);
}

}  // extern "C"

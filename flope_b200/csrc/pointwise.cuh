// flope_b200: small HBM-bound kernels around the backbone (sm_100a).
//   ingest_nchw_f32 : (B,3,S,S) fp32 [0,1] -> space-to-depth blocked-pixel bf16 (the stem's input)
//   maxpool3x3s2    : torchvision ResNet stem maxpool (3x3, stride 2, pad 1) on blocked-pixel bf16
//   avgpool         : AdaptiveAvgPool2d(1) (sunflower/models/posenet.py:12) -> blocked "pixel = crop" bf16
//   unpack_to_nchw  : blocked-pixel bf16 -> NCHW fp32 (tests / debugging only)
#pragma once
#include "common.cuh"

namespace flope {

// Space-to-depth stem input: grid (S/2 x S/2), 2 planes.  Plane `by` (= y&1) pixel (y>>1, x>>1)
// holds 8 bf16: [bx=0: c0 c1 c2 0 | bx=1: c0 c1 c2 0].
__global__ void ingest_nchw_f32_kernel(const float* __restrict__ x, int n, int S, __nv_bfloat16* __restrict__ out,
                                       Geom g) {
  // one thread = 4 consecutive pixels of one row: three 16-byte loads (one per channel plane),
  // one 32-byte run of two s2d pixels out
  griddep_wait();
  griddep_launch();
  // 32-bit index arithmetic (n*S*S/4 < 2^31, checked by the host): the 64-bit div/mod sequences of the first version
  // made this copy kernel issue-bound (~400 instructions per 48 bytes read)
  const uint32_t S4 = (uint32_t)S >> 2;
  const uint32_t total = (uint32_t)n * (uint32_t)S * S4;
  const long long plane_sz = (long long)S * S;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / S4;                 // row index b*S + yy
    const int x4 = (int)(i - r * S4);
    const int b = (int)(r / (uint32_t)S);
    const int yy = (int)(r - (uint32_t)b * (uint32_t)S);
    const float* px = x + ((long long)b * 3 * S + yy) * S + 4 * x4;
    const float4 c0 = __ldcs(reinterpret_cast<const float4*>(px));
    const float4 c1 = __ldcs(reinterpret_cast<const float4*>(px + plane_sz));
    const float4 c2 = __ldcs(reinterpret_cast<const float4*>(px + 2 * plane_sz));
    uint4 o0, o1;
    o0.x = pack_bf16x2(c0.x, c1.x); o0.y = pack_bf16x2(c2.x, 0.f);
    o0.z = pack_bf16x2(c0.y, c1.y); o0.w = pack_bf16x2(c2.y, 0.f);
    o1.x = pack_bf16x2(c0.z, c1.z); o1.y = pack_bf16x2(c2.z, 0.f);
    o1.z = pack_bf16x2(c0.w, c1.w); o1.w = pack_bf16x2(c2.w, 0.f);
    const long long pos = g.base + geom_pos(g, b, yy >> 1, 2 * x4);
    uint4* dst = reinterpret_cast<uint4*>(out + ((long long)(yy & 1) * g.plane + pos) * 8);
    dst[0] = o0;
    dst[1] = o1;
  }
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Inputs are post-ReLU (>= 0) and padded positions hold zeros, so zero padding == -inf padding.
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, Geom gi, __nv_bfloat16* __restrict__ out,
                                    Geom go, int n) {
  griddep_wait();
  griddep_launch();
  const int C8 = gi.C >> 3;
  const long long total = (long long)C8 * n * go.H * go.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % go.W);
    const int oh = (int)((i / go.W) % go.H);
    const int b = (int)((i / ((long long)go.W * go.H)) % n);
    const int c8 = (int)(i / ((long long)go.W * go.H * n));
    const __nv_bfloat16* src = in + ((long long)c8 * gi.plane + gi.base) * 8;
    uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        // (2*oh+dy, 2*ow+dx): -1 lands on the shared zero pad row/column of the blocked layout
        const long long pp = geom_pos(gi, b, 2 * oh, 2 * ow) + dy * gi.Wp + dx;
        const uint4 v = *reinterpret_cast<const uint4*>(src + pp * 8);
        m.x = bf16x2_max(m.x, v.x); m.y = bf16x2_max(m.y, v.y);
        m.z = bf16x2_max(m.z, v.z); m.w = bf16x2_max(m.w, v.w);
      }
    }
    __nv_bfloat16* dst = out + ((long long)c8 * go.plane + go.base + geom_pos(go, b, oh, ow)) * 8;
    *reinterpret_cast<uint4*>(dst) = m;
  }
}

// One warp per (c8 plane, crop): the crop's Hp*Wp positions of a plane are contiguous and its padding positions
// hold zeros (the layout's invariant - every conv relies on it), so the lanes stride over ALL of them with 16-byte
// loads and no index arithmetic; fp32 partial sums; a reduce-scatter over the lanes (each xor step halves the
// channels a lane still carries: 4+2+1+1+1 shuffles instead of 8x5); lane 0 writes the bf16 means into the "one
// pixel per crop" blocked tensor the fc GEMM reads (plane c8, position = crop index).
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ in, Geom gi, __nv_bfloat16* __restrict__ out,
                               Geom go, int n) {
  griddep_wait();
  griddep_launch();
  const int C8 = gi.C >> 3;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= C8 * n) return;
  const int b = wid % n;
  const int c8 = wid / n;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int img = gi.Hp * gi.Wp;
  const uint4* src = reinterpret_cast<const uint4*>(in + ((long long)c8 * gi.plane + gi.base + (long long)b * img) * 8);
  for (int i = lane; i < img; i += 32) {
    const uint4 v = src[i];
    s[0] += bf16_lo(v.x); s[1] += bf16_hi(v.x); s[2] += bf16_lo(v.y); s[3] += bf16_hi(v.y);
    s[4] += bf16_lo(v.z); s[5] += bf16_hi(v.z); s[6] += bf16_lo(v.w); s[7] += bf16_hi(v.w);
  }
  // lanes with bit 4 set keep channels 4..7, the others 0..3; then bit 3 splits those four into two, bit 2 into one
  float t4[4], t2[2], t1;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float give = hi ? s[j] : s[4 + j], keep = hi ? s[4 + j] : s[j];
      t4[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float give = hi ? t4[j] : t4[2 + j], keep = hi ? t4[2 + j] : t4[j];
      t2[j] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
  }
  {
    const bool hi = lane & 4;
    const float give = hi ? t2[0] : t2[1], keep = hi ? t2[1] : t2[0];
    t1 = keep + __shfl_xor_sync(0xffffffffu, give, 4);
  }
  t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
  t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
  // lane 4*ch (ch = 0..7) now holds the sum of channel ch: 4*bit4 + 2*bit3 + bit2 of the lane index
  if (!(lane & 3)) {
    const int ch = lane >> 2;
    out[((long long)c8 * go.plane + go.base + b) * 8 + ch] = __float2bfloat16_rn(t1 * (1.0f / (float)(gi.H * gi.W)));
  }
}

// parity = 0: plain tensor.  parity = 1: parity-split tensor whose Geom describes the half-res grid
// and whose logical size is (2H x 2W); plane index = ((h&1)*2 + (w&1)) * C/8 + c/8.
__global__ void unpack_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, Geom g, int parity, int n,
                                      float* __restrict__ out) {
  const int H = parity ? 2 * g.H : g.H, W = parity ? 2 * g.W : g.W;
  const long long total = (long long)n * g.C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % g.C);
    const int b = (int)(i / ((long long)W * H * g.C));
    long long plane = c >> 3, pos;
    if (parity) {
      plane += (((h & 1) << 1) | (w & 1)) * (g.C >> 3);
      pos = g.base + geom_pos(g, b, h >> 1, w >> 1);
    } else {
      pos = g.base + geom_pos(g, b, h, w);
    }
    out[i] = __bfloat162float(in[(plane * g.plane + pos) * 8 + (c & 7)]);
  }
}

}  // namespace flope

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
  const long long total = (long long)n * S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % S);
    const int yy = (int)((i / S) % S);
    const int b = (int)(i / ((long long)S * S));
    const float* px = x + ((long long)b * 3 * S + yy) * S + xx;
    const long long plane_sz = (long long)S * S;
    const float c0 = px[0], c1 = px[plane_sz], c2 = px[2 * plane_sz];
    uint2 o;
    o.x = pack_bf16x2(c0, c1);
    o.y = pack_bf16x2(c2, 0.f);
    const long long pos = g.base + geom_pos(g, b, yy >> 1, xx >> 1);
    __nv_bfloat16* dst = out + ((long long)(yy & 1) * g.plane + pos) * 8 + (xx & 1) * 4;
    *reinterpret_cast<uint2*>(dst) = o;
  }
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Inputs are post-ReLU (>= 0) and padded positions hold zeros, so zero padding == -inf padding.
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, Geom gi, __nv_bfloat16* __restrict__ out,
                                    Geom go, int n) {
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

// One thread per (c8 plane, crop): mean of H*W pixels in fp32, written as bf16 into the
// "one pixel per crop" blocked tensor the fc GEMM reads (plane c8, position = crop index).
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ in, Geom gi, __nv_bfloat16* __restrict__ out,
                               Geom go, int n) {
  const int C8 = gi.C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C8 * n) return;
  const int b = i % n;
  const int c8 = i / n;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __nv_bfloat16* src = in + ((long long)c8 * gi.plane + gi.base) * 8;
  for (int h = 0; h < gi.H; ++h) {
    for (int w = 0; w < gi.W; ++w) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + geom_pos(gi, b, h, w) * 8);
      s[0] += bf16_lo(v.x); s[1] += bf16_hi(v.x); s[2] += bf16_lo(v.y); s[3] += bf16_hi(v.y);
      s[4] += bf16_lo(v.z); s[5] += bf16_hi(v.z); s[6] += bf16_lo(v.w); s[7] += bf16_hi(v.w);
    }
  }
  const float inv = 1.0f / (float)(gi.H * gi.W);
  uint4 o;
  o.x = pack_bf16x2(s[0] * inv, s[1] * inv); o.y = pack_bf16x2(s[2] * inv, s[3] * inv);
  o.z = pack_bf16x2(s[4] * inv, s[5] * inv); o.w = pack_bf16x2(s[6] * inv, s[7] * inv);
  *reinterpret_cast<uint4*>(out + ((long long)c8 * go.plane + go.base + b) * 8) = o;
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

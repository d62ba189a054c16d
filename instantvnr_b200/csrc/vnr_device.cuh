// vnr_device.cuh -- device-side model description and the hash-grid encode of one
// (sample, level), shared by the decode, marcher and training kernels.
//
// Arithmetic follows the reference's kernel_grid (tcnn encodings/grid.h:120-243),
// grid_index/fast_hash (:64-99) and pos_fract (common_device.h:405-412): fp32 position
// and weights, fp16 table values, trilinear sum ACCUMULATED IN HALF in corner order
// idx = 0..7 (bit d of idx selects the +1 neighbour along dim d).
// The library is compiled with -fmad=false; every fused multiply-add is explicit.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace vnr {

constexpr int kMaxLevels = 16;
constexpr int kWidth = 64;        // n_neurons of the fully-fused MLP (example-model.json:28)
constexpr int kOutPad = 16;       // padded output rows (fully_fused_mlp.cu:677)
constexpr int kTile = 128;        // samples per tensor-core tile (UMMA M)
constexpr int kMaxHidden = 8;

struct LevelDesc {
  uint32_t offset;      // first entry of the level (in entries)
  uint32_t size;        // entries in the level ("hashmap_size")
  uint32_t res;         // grid resolution
  uint32_t res2;        // res*res (uint32 wrap)
  float scale;          // exp2f(l*log2(pls))*base - 1
  uint32_t hashed;      // 1: fast_hash, 0: dense index
  uint32_t mask;        // size-1 if size is a power of two else 0 (use %)
  uint32_t pad_;
};

struct DecoderDesc {
  int n_levels;
  int n_feat;           // F in {1,2,4,8}
  int enc_dims;         // L*F
  int enc_pad;          // padded to a multiple of 16 (<= 64)
  int n_hidden;         // hidden layers (each a ReLU matmul into 64 neurons)
  uint32_t n_mlp;       // number of MLP params (grid table starts there)
  uint32_t n_grid;
  uint32_t pad_;
  LevelDesc lv[kMaxLevels];
};

__device__ __forceinline__ uint32_t level_index(const LevelDesc& lv, uint32_t x, uint32_t y, uint32_t z) {
  uint32_t idx;
  if (lv.hashed) idx = x ^ (y * 2654435761u) ^ (z * 805459861u);
  else           idx = x + y * lv.res + z * lv.res2;
  return lv.mask ? (idx & lv.mask) : (idx % lv.size);
}

struct CornerSetup {
  uint32_t gx, gy, gz;
  float wx, wy, wz;
};

__device__ __forceinline__ CornerSetup corner_setup(const LevelDesc& lv, float x, float y, float z) {
  CornerSetup c;
  float px = __fmaf_rn(x, lv.scale, 0.5f), py = __fmaf_rn(y, lv.scale, 0.5f), pz = __fmaf_rn(z, lv.scale, 0.5f);
  float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  c.gx = (uint32_t)(int)fx; c.gy = (uint32_t)(int)fy; c.gz = (uint32_t)(int)fz;
  c.wx = px - fx; c.wy = py - fy; c.wz = pz - fz;
  return c;
}

__device__ __forceinline__ float corner_weight(const CornerSetup& c, int idx) {
  float w = (idx & 1) ? c.wx : 1.f - c.wx;
  w *= (idx & 2) ? c.wy : 1.f - c.wy;
  w *= (idx & 4) ? c.wz : 1.f - c.wz;
  return w;
}

__device__ __forceinline__ uint32_t corner_index(const LevelDesc& lv, const CornerSetup& c, int idx) {
  return level_index(lv, c.gx + (idx & 1), c.gy + ((idx >> 1) & 1), c.gz + ((idx >> 2) & 1));
}

__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// result += (half)(w * (float)v)   for a packed pair
__device__ __forceinline__ __half2 acc_pair(__half2 acc, uint32_t v, float w) {
  float2 f = __half22float2(u32_as_h2(v));
  return __hadd2(acc, __floats2half2_rn(w * f.x, w * f.y));
}

// Encode one level with F = 8: returns the 8 halves as one 16-byte vector.
__device__ __forceinline__ uint4 encode_level_f8(const LevelDesc& lv, const __half* __restrict__ grid, float x, float y, float z) {
  const CornerSetup c = corner_setup(lv, x, y, z);
  const uint4* __restrict__ tab = reinterpret_cast<const uint4*>(grid) + lv.offset;
  uint4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __ldg(tab + corner_index(lv, c, i));
  __half2 a0 = __float2half2_rn(0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float w = corner_weight(c, i);
    a0 = acc_pair(a0, v[i].x, w); a1 = acc_pair(a1, v[i].y, w);
    a2 = acc_pair(a2, v[i].z, w); a3 = acc_pair(a3, v[i].w, w);
  }
  return make_uint4(h2_as_u32(a0), h2_as_u32(a1), h2_as_u32(a2), h2_as_u32(a3));
}

// F = 4: 8 bytes
__device__ __forceinline__ uint2 encode_level_f4(const LevelDesc& lv, const __half* __restrict__ grid, float x, float y, float z) {
  const CornerSetup c = corner_setup(lv, x, y, z);
  const uint2* __restrict__ tab = reinterpret_cast<const uint2*>(grid) + lv.offset;
  uint2 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __ldg(tab + corner_index(lv, c, i));
  __half2 a0 = __float2half2_rn(0.f), a1 = a0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float w = corner_weight(c, i);
    a0 = acc_pair(a0, v[i].x, w); a1 = acc_pair(a1, v[i].y, w);
  }
  return make_uint2(h2_as_u32(a0), h2_as_u32(a1));
}

// F = 2: 4 bytes
__device__ __forceinline__ uint32_t encode_level_f2(const LevelDesc& lv, const __half* __restrict__ grid, float x, float y, float z) {
  const CornerSetup c = corner_setup(lv, x, y, z);
  const uint32_t* __restrict__ tab = reinterpret_cast<const uint32_t*>(grid) + lv.offset;
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __ldg(tab + corner_index(lv, c, i));
  __half2 a0 = __float2half2_rn(0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) a0 = acc_pair(a0, v[i], corner_weight(c, i));
  return h2_as_u32(a0);
}

// F = 1: 2 bytes
__device__ __forceinline__ __half encode_level_f1(const LevelDesc& lv, const __half* __restrict__ grid, float x, float y, float z) {
  const CornerSetup c = corner_setup(lv, x, y, z);
  const __half* __restrict__ tab = grid + lv.offset;
  __half v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __ldg(tab + corner_index(lv, c, i));
  __half a = __float2half_rn(0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) a = __hadd(a, __float2half_rn(corner_weight(c, i) * __half2float(v[i])));
  return a;
}

}  // namespace vnr

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
constexpr int kMaxPeers = 8;       // GPUs of one NVSwitch box
constexpr int kMaxDevices = 16;    // per-device caches of per-function attributes

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

// Table indices of the 8 corners of a cell.  The level descriptor is warp-uniform, so the
// hashed/dense and power-of-two/modulo choices are real (uniform) branches, and the per-dimension
// terms are shared between corners: (g+1)*P == g*P + P (mod 2^32).
__device__ __forceinline__ void level_indices(const LevelDesc& lv, const CornerSetup& c, uint32_t (&idx)[8]) {
  const uint32_t xa[2] = {c.gx, c.gx + 1u};
  if (lv.hashed) {
    const uint32_t y0 = c.gy * 2654435761u, z0 = c.gz * 805459861u;
    const uint32_t yb[2] = {y0, y0 + 2654435761u}, zc[2] = {z0, z0 + 805459861u};
#pragma unroll
    for (int i = 0; i < 8; ++i) idx[i] = xa[i & 1] ^ yb[(i >> 1) & 1] ^ zc[i >> 2];
  } else {
    const uint32_t y0 = c.gy * lv.res, z0 = c.gz * lv.res2;
    const uint32_t yb[2] = {y0, y0 + lv.res}, zc[2] = {z0, z0 + lv.res2};
#pragma unroll
    for (int i = 0; i < 8; ++i) idx[i] = xa[i & 1] + yb[(i >> 1) & 1] + zc[i >> 2];
  }
  if (lv.mask) {
#pragma unroll
    for (int i = 0; i < 8; ++i) idx[i] &= lv.mask;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) idx[i] %= lv.size;
  }
}

// The 8 trilinear weights, ((wx * wy) * wz) as the sequential `weight *= ...` of the reference.
__device__ __forceinline__ void corner_weights(float wx, float wy, float wz, float (&w)[8]) {
  const float ax[2] = {1.f - wx, wx}, ay[2] = {1.f - wy, wy}, az[2] = {1.f - wz, wz};
  float xy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) xy[i] = ax[i & 1] * ay[i >> 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = xy[i & 3] * az[i >> 2];
}

// Table loads.  LD 0: ld.global.nc (allocates an L1 line per miss); 1: ld.global.nc.L1::no_allocate; 2: ld.global.cg (L2 only).
// A CTA that takes most of the SM's shared memory leaves a small L1 (228 KB - smem): with LD 0 the lines of the misses in
// flight are bounded by it.
template <int LD, typename T> __device__ __forceinline__ T table_load(const T* p) {
  if constexpr (LD == 0) return __ldg(p);
  else if constexpr (sizeof(T) == 16) {
    uint4 v;
    if constexpr (LD == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return *reinterpret_cast<T*>(&v);
  } else return __ldcg(p);
}

template <int F> struct FeatVec;
template <> struct FeatVec<8> { typedef uint4 type; };
template <> struct FeatVec<4> { typedef uint2 type; };
template <> struct FeatVec<2> { typedef uint32_t type; };
template <> struct FeatVec<1> { typedef unsigned short type; };

// One (sample, level) gather split in two halves so that callers can keep the loads of the next
// level in flight while the current one is reduced: issue() computes the corner indices and
// starts the 8 vector loads, finish() does the fp16-accumulated trilinear sum.
template <int F, int LD = 0>
struct LevelGather {
  typedef typename FeatVec<F>::type T;
  T v[8];
  float wx, wy, wz;

  __device__ __forceinline__ void issue(const LevelDesc& lv, const __half* __restrict__ grid, float x, float y, float z) {
    const CornerSetup c = corner_setup(lv, x, y, z);
    wx = c.wx; wy = c.wy; wz = c.wz;
    uint32_t idx[8];
    level_indices(lv, c, idx);
    const T* __restrict__ tab = reinterpret_cast<const T*>(grid) + lv.offset;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = table_load<LD>(tab + idx[i]);
  }

  __device__ __forceinline__ T finish() const {
    float w[8];
    corner_weights(wx, wy, wz, w);
    if constexpr (F == 8) {
      __half2 a0 = __float2half2_rn(0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a0 = acc_pair(a0, v[i].x, w[i]); a1 = acc_pair(a1, v[i].y, w[i]);
        a2 = acc_pair(a2, v[i].z, w[i]); a3 = acc_pair(a3, v[i].w, w[i]);
      }
      return make_uint4(h2_as_u32(a0), h2_as_u32(a1), h2_as_u32(a2), h2_as_u32(a3));
    } else if constexpr (F == 4) {
      __half2 a0 = __float2half2_rn(0.f), a1 = a0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { a0 = acc_pair(a0, v[i].x, w[i]); a1 = acc_pair(a1, v[i].y, w[i]); }
      return make_uint2(h2_as_u32(a0), h2_as_u32(a1));
    } else if constexpr (F == 2) {
      __half2 a0 = __float2half2_rn(0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) a0 = acc_pair(a0, v[i], w[i]);
      return h2_as_u32(a0);
    } else {
      __half a = __float2half_rn(0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) a = __hadd(a, __float2half_rn(w[i] * __half2float(__ushort_as_half(v[i]))));
      return __half_as_ushort(a);
    }
  }
};

// byte offset of feature column `col` (F halves wide) inside a 128-byte swizzled tile row
template <int F>
__device__ __forceinline__ uint32_t feat_offset(uint32_t level, uint32_t row_sw) {
  constexpr uint32_t per_chunk = 8u / F;                      // levels per 16-byte chunk
  return (((level / per_chunk) ^ row_sw) << 4) + (level % per_chunk) * (2u * F);
}

// Gather levels [l0, l1) of one sample into its swizzled tile row, software-pipelined: the
// loads of level l+1 are in flight while level l is reduced.
template <int F, int LD = 0>
__device__ __forceinline__ void encode_levels(uint8_t* rowp, uint32_t row_sw, const DecoderDesc& d, const __half* __restrict__ grid,
                                              float x, float y, float z, int l0, int l1) {
  typedef typename FeatVec<F>::type T;
  if (l0 >= l1) return;
  LevelGather<F, LD> cur, nxt;
  cur.issue(d.lv[l0], grid, x, y, z);
  for (int l = l0; l < l1; ++l) {
    if (l + 1 < l1) nxt.issue(d.lv[l + 1], grid, x, y, z);
    *reinterpret_cast<T*>(rowp + feat_offset<F>((uint32_t)l, row_sw)) = cur.finish();
    cur = nxt;
  }
}

// Same without the cross-level software pipeline (one level's 8 loads in flight): fewer live registers, for
// kernels whose gather threads are capped low and whose gather is not the bound (training).
template <int F, int LD = 0>
__device__ __forceinline__ void encode_levels_lean(uint8_t* rowp, uint32_t row_sw, const DecoderDesc& d, const __half* __restrict__ grid,
                                                   float x, float y, float z, int l0, int l1) {
  typedef typename FeatVec<F>::type T;
  for (int l = l0; l < l1; ++l) {
    LevelGather<F, LD> g;
    g.issue(d.lv[l], grid, x, y, z);
    *reinterpret_cast<T*>(rowp + feat_offset<F>((uint32_t)l, row_sw)) = g.finish();
  }
}

// Gather all levels of one sample into row `row` of a 128B-swizzled A tile (128 rows of 64 halves) and zero the
// padding features up to enc_pad (tcnn pads the encoding to a multiple of 16: grid.h:616-620).
template <int F, bool PIPELINED = true, int LD = 0>
__device__ __forceinline__ void encode_row(uint8_t* a_smem, const DecoderDesc& d, const __half* __restrict__ grid, float x, float y, float z, uint32_t row) {
  uint8_t* rowp = a_smem + row * 128u;
  const uint32_t sw = (row & 7u);
  if constexpr (PIPELINED) encode_levels<F, LD>(rowp, sw, d, grid, x, y, z, 0, d.n_levels);
  else encode_levels_lean<F, LD>(rowp, sw, d, grid, x, y, z, 0, d.n_levels);
  for (int k = d.enc_dims; k < d.enc_pad; ++k)
    *reinterpret_cast<__half*>(rowp + ((((uint32_t)k >> 3) ^ sw) << 4) + ((uint32_t)k & 7u) * 2u) = __float2half_rn(0.f);
}

}  // namespace vnr

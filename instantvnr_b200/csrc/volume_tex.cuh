// volume_tex.cuh -- software restatement of the 3-D linear-filtered texture fetch the reference issues on its volume
// textures (tex3D<float>, normalized coordinates, clamp addressing; createCudaTexture is un-vendored OVR code).
// Linear filtering uses the texture unit's 1.8 fixed-point weights (CUDA C Programming Guide, "Linear Filtering"), so the
// result is bit-identical to the oracle's tex3d_linear and within one weight quantum of the hardware unit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vnr {

__device__ __forceinline__ float tex_frac(float xb, float fl) {
  const float fr = xb - fl;
  return floorf(__fmaf_rn(fr, 256.f, 0.5f)) * (1.f / 256.f);
}

__device__ __forceinline__ float sample_volume_linear(const float* __restrict__ vol, int3 dims, float u, float v, float w) {
  const float cx = __fmaf_rn(u, (float)dims.x, -0.5f), cy = __fmaf_rn(v, (float)dims.y, -0.5f), cz = __fmaf_rn(w, (float)dims.z, -0.5f);
  const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
  const float ax = tex_frac(cx, fx), ay = tex_frac(cy, fy), az = tex_frac(cz, fz);
  const int x0 = min(max((int)fx, 0), dims.x - 1), x1 = min(max((int)fx + 1, 0), dims.x - 1);
  const int y0 = min(max((int)fy, 0), dims.y - 1), y1 = min(max((int)fy + 1, 0), dims.y - 1);
  const int z0 = min(max((int)fz, 0), dims.z - 1), z1 = min(max((int)fz + 1, 0), dims.z - 1);
  const size_t sx = 1, sy = (size_t)dims.x, sz = (size_t)dims.x * dims.y;
  auto at = [&](int x, int y, int z) { return __ldg(vol + x * sx + y * sy + z * sz); };
  auto lerp = [](float t, float p, float q) { return __fmaf_rn(t, q, (1.f - t) * p); };
  const float c00 = lerp(ax, at(x0, y0, z0), at(x1, y0, z0));
  const float c10 = lerp(ax, at(x0, y1, z0), at(x1, y1, z0));
  const float c01 = lerp(ax, at(x0, y0, z1), at(x1, y0, z1));
  const float c11 = lerp(ax, at(x0, y1, z1), at(x1, y1, z1));
  return lerp(az, lerp(ay, c00, c10), lerp(ay, c01, c11));
}

// sampleVolume (raytracing.h:105-110): `p = p * (1 - rdims) + 0.5 * rdims; tex3D(p)` -- with Array3DScalar::rdims = 0: the
// member is default-initialised (core/array.h:43) and nothing in the reference ever assigns it (set_volume, object.cpp:362-383,
// sets dims / data / type only), so the lookup is the plain normalised-coordinate texture fetch.  Confirmed against the
// reference's own marcher compiled in place (oracle/ref_marcher).
__device__ __forceinline__ float sample_volume(const float* __restrict__ vol, int3 dims, float x, float y, float z) {
  return sample_volume_linear(vol, dims, x, y, z);
}

}  // namespace vnr

// macrocell.cu -- macrocell value-range and max-opacity grids (space skipping).
//
// Follows core/macrocell.cu: update_single_macrocell :11-40 (ranges stored offset by -1 / +1
// so that zero-initialised memory works with atomicMin/Max), update_macrocell_explicit :42-73
// (online, from every training batch), update_macrocell_implicit :75-111 (offline, from the
// ground-truth volume; one launch here instead of one per z-slice :223-229),
// macrocell_max_opacity_kernel :153-193.  MACROCELL_SIZE_MIP = 4 (CMakeLists.txt:68).
#include "volume.h"
#include "train.h"

namespace vnr {

constexpr int kMcMip = 4;
constexpr int kMcSize = 1 << kMcMip;

// float atomics through integer atomics (core/instantvnr_types.h:185-199)
__device__ __forceinline__ void atomic_min_f(float* addr, float value) {
  if (!signbit(value)) atomicMin((int*)addr, __float_as_int(value));
  else atomicMax((unsigned int*)addr, __float_as_uint(value));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float value) {
  if (!signbit(value)) atomicMax((int*)addr, __float_as_int(value));
  else atomicMin((unsigned int*)addr, __float_as_uint(value));
}

__device__ __forceinline__ void update_single(int x, int y, int z, int3 md, float* __restrict__ mc, float value) {
  const int cx = x >> kMcMip, cy = y >> kMcMip, cz = z >> kMcMip;
  if (cx < 0 || cx >= md.x || cy < 0 || cy >= md.y || cz < 0 || cz >= md.z) return;
  const uint32_t idx = cx + cy * md.x + cz * md.y * md.x;
  atomic_min_f(mc + 2 * idx, value - 1.f);
  atomic_max_f(mc + 2 * idx + 1, value + 1.f);
}

__device__ __forceinline__ void update_voxel(uint32_t x, uint32_t y, uint32_t z, int3 md, float* __restrict__ mc, float value) {
  const int sx = (x % kMcSize) == 0 ? -1 : (x % kMcSize) == (kMcSize - 1) ? 1 : 0;
  const int sy = (y % kMcSize) == 0 ? -1 : (y % kMcSize) == (kMcSize - 1) ? 1 : 0;
  const int sz = (z % kMcSize) == 0 ? -1 : (z % kMcSize) == (kMcSize - 1) ? 1 : 0;
  const int X = (int)x, Y = (int)y, Z = (int)z;
  update_single(X, Y, Z, md, mc, value);
  if (sx) update_single(X + sx, Y, Z, md, mc, value);
  if (sy) update_single(X, Y + sy, Z, md, mc, value);
  if (sx && sy) update_single(X + sx, Y + sy, Z, md, mc, value);
  if (sz) {
    update_single(X, Y, Z + sz, md, mc, value);
    if (sx) update_single(X + sx, Y, Z + sz, md, mc, value);
    if (sy) update_single(X, Y + sy, Z + sz, md, mc, value);
    if (sx && sy) update_single(X + sx, Y + sy, Z + sz, md, mc, value);
  }
}

__global__ void macrocell_explicit_kernel(uint32_t n, const float* __restrict__ coords, const float* __restrict__ values, int3 dims, int3 md, float* __restrict__ mc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float cx = coords[3 * (size_t)i], cy = coords[3 * (size_t)i + 1], cz = coords[3 * (size_t)i + 2];
  const uint32_t x = min(max((uint32_t)floorf(cx * dims.x), 0u), (uint32_t)(dims.x - 1));
  const uint32_t y = min(max((uint32_t)floorf(cy * dims.y), 0u), (uint32_t)(dims.y - 1));
  const uint32_t z = min(max((uint32_t)floorf(cz * dims.z), 0u), (uint32_t)(dims.z - 1));
  update_voxel(x, y, z, md, mc, values[i]);
}

// One thread per voxel; a block-level min/max pre-reduction is not needed for correctness and
// the kernel runs once per ground-truth upload.
__global__ void macrocell_implicit_kernel(uint64_t n_vox, const float* __restrict__ volume, int3 dims, int3 md, float* __restrict__ mc) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_vox) return;
  const uint64_t stride = (uint64_t)dims.x * dims.y;
  const uint32_t x = (uint32_t)(idx % dims.x), y = (uint32_t)((idx % stride) / dims.x), z = (uint32_t)(idx / stride);
  update_voxel(x, y, z, md, mc, volume[idx]);
}

__global__ void macrocell_max_opacity_kernel(uint32_t n_cells, const float* __restrict__ alphas, int n_alpha, float lo, float hi, float rcp,
                                             const float2* __restrict__ range, float* __restrict__ out) {
  extern __shared__ float s_alpha[];
  for (int k = threadIdx.x; k < n_alpha; k += blockDim.x) s_alpha[k] = alphas[k];
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  float2 r = range[i];
  r.x += 1.f; r.y -= 1.f;
  const float lower = (fminf(fmaxf(r.x, lo), hi) - lo) * rcp;
  const float upper = (fminf(fmaxf(r.y, lo), hi) - lo) * rcp;
  // uint32_t i_lower = floorf(fmaf(lower, n-1, 0.5f)) - 1: the float -> uint32 conversion saturates at 0
  const float fl = floorf(__fmaf_rn(lower, (float)(n_alpha - 1), 0.5f)) - 1.f;
  uint32_t il = fl <= 0.f ? 0u : (uint32_t)fl;
  uint32_t iu = (uint32_t)(floorf(__fmaf_rn(upper, (float)(n_alpha - 1), 0.5f)) + 1.f);
  il = min(il, (uint32_t)(n_alpha - 1));
  iu = min(iu, (uint32_t)(n_alpha - 1));
  float op = 0.f;
  for (uint32_t k = il; k <= iu; ++k) op = fmaxf(op, s_alpha[k]);
  out[i] = op;
}

void macrocell_preload_kernels() {
  cudaFuncAttributes fa;
  VNR_CUDA(cudaFuncGetAttributes(&fa, macrocell_explicit_kernel));
  VNR_CUDA(cudaFuncGetAttributes(&fa, macrocell_max_opacity_kernel));
}

void macrocell_update_explicit(Volume* v, const float* d_xyz, const float* d_values, size_t n, cudaStream_t s) {
  if (!n) return;
  const int3 dims = make_int3(v->dims[0], v->dims[1], v->dims[2]), md = make_int3(v->mc_dims[0], v->mc_dims[1], v->mc_dims[2]);
  macrocell_explicit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((uint32_t)n, d_xyz, d_values, dims, md, v->mc_range.p);
  VNR_CUDA(cudaGetLastError());
}

void macrocell_update_implicit(Volume* v, cudaStream_t s) {
  if (!v->have_gt) throw StateError("no ground-truth volume set");
  const int3 dims = make_int3(v->dims[0], v->dims[1], v->dims[2]), md = make_int3(v->mc_dims[0], v->mc_dims[1], v->mc_dims[2]);
  const uint64_t n = (uint64_t)dims.x * dims.y * dims.z;
  macrocell_implicit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, v->gt.p, dims, md, v->mc_range.p);
  VNR_CUDA(cudaGetLastError());
}

void macrocell_update_max_opacity(Volume* v, cudaStream_t s) {
  wait_for_frames(v, s);                 // frames in flight still walk the current max-opacity grid
  if (v->n_alpha <= 0) return;                       // macrocell.cu:245
  const uint32_t n = (uint32_t)v->cells();
  const float rcp = 1.f / (v->tfn_hi - v->tfn_lo);
  macrocell_max_opacity_kernel<<<(n + 255) / 256, 256, v->n_alpha * sizeof(float), s>>>(n, v->tfn_alpha.p, v->n_alpha, v->tfn_lo, v->tfn_hi, rcp,
                                                                                         reinterpret_cast<const float2*>(v->mc_range.p), v->mc_maxop.p);
  VNR_CUDA(cudaGetLastError());
}

}  // namespace vnr

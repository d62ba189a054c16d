// probe.cu -- memory-system microbenchmarks that give the decode / training rooflines their denominators.
//
// Measurement taps only (no reference counterpart, nothing of the product path runs through them).  They are
// deliberately INDEPENDENT of the product's gather code (LevelGather / scatter_level): plain grid-stride kernels at
// full occupancy that issue
//   * random 16-byte read-only loads (ld.global.nc.v4) over a buffer of a given size -- 46.7 MB (the example model's
//     table, L2-resident on B200) gives `l2_gather_gbs`, 306.8 MB (T = 2^22, larger than the 126 MB L2) gives
//     `hbm_gather_gbs`;
//   * random 16-byte fp16 vector reductions (red.global.add.noftz.v4.f16x2), the instruction of the hash-grid
//     backward (train.cu), which gives the scatter ceiling of the training step.
// Addresses come from a counter-based integer hash, so consecutive loads of a thread are independent (8 in flight).
#include <cstdint>
#include <cstring>

#include "../../include/vnr_c.h"
#include "vnr_host.h"

namespace vnr {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {          // murmur3 finalizer
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// every thread: `per_thread` loads in batches of 8 independent ones; n_vec = number of 16-byte vectors in the table
__global__ void __launch_bounds__(256) probe_loads_kernel(const uint4* __restrict__ table, uint32_t n_vec, uint32_t per_thread, uint32_t seed,
                                                          uint32_t* __restrict__ sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t fold = 0;
  uint32_t c = mix32(t * 0x9e3779b9u + seed);
  for (uint32_t k = 0; k < per_thread; k += 8) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c = c * 1664525u + 1013904223u;
      const uint32_t idx = (uint32_t)(((uint64_t)mix32(c) * n_vec) >> 32);
      v[i] = ldg_nc_v4(table + idx);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) fold ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (fold == 0x12345678u) sink[t & 1023u] = fold;               // practically never: keeps the loads alive
}

__global__ void __launch_bounds__(256) probe_reds_kernel(__half* __restrict__ table, uint32_t n_vec, uint32_t per_thread, uint32_t seed) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c = mix32(t * 0x9e3779b9u + seed);
  const uint32_t one = 0x04000400u;                               // two small fp16 values (2^-14): sums stay finite
  for (uint32_t k = 0; k < per_thread; ++k) {
    c = c * 1664525u + 1013904223u;
    const uint32_t idx = (uint32_t)(((uint64_t)mix32(c) * n_vec) >> 32);
    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %1, %1, %1};" ::"l"(table + (size_t)idx * 8), "r"(one) : "memory");
  }
}

// random 32-byte loads (one 256-bit LDG per 32-byte sector): is the L1 bound per REQUEST or per byte?  If this kernel moves as many
// requests per second as probe_loads_kernel, fetching the two x-neighbours of a cell with one load pays.
__global__ void __launch_bounds__(256) probe_loads32_kernel(const uint4* __restrict__ table, uint32_t n_vec, uint32_t per_thread, uint32_t seed,
                                                            uint32_t* __restrict__ sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t fold = 0;
  uint32_t c = mix32(t * 0x9e3779b9u + seed);
  const uint32_t n_pair = n_vec / 2;
  for (uint32_t k = 0; k < per_thread; k += 4) {
    uint32_t v[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      c = c * 1664525u + 1013904223u;
      const uint32_t idx = (uint32_t)(((uint64_t)mix32(c) * n_pair) >> 32);
      asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]), "=r"(v[i][4]), "=r"(v[i][5]), "=r"(v[i][6]), "=r"(v[i][7])
                   : "l"(table + 2 * (size_t)idx));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) fold ^= v[i][q];
  }
  if (fold == 0x12345678u) sink[t & 1023u] = fold;
}

// the same random loads from ONE resident CTA per SM (grid = number of SMs) whose dynamic shared memory shrinks the L1: how the
// gather rate of a warp-specialised persistent CTA depends on its thread count and on the L1 left over by its shared memory
__global__ void probe_loads_cta_kernel(const uint4* __restrict__ table, uint32_t n_vec, uint32_t per_thread, uint32_t seed, uint32_t* __restrict__ sink) {
  extern __shared__ uint8_t probe_smem[];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (per_thread == 0xFFFFFFFFu) probe_smem[threadIdx.x] = 1;     // never: keeps the allocation referenced
  uint32_t fold = 0;
  uint32_t c = mix32(t * 0x9e3779b9u + seed);
  for (uint32_t k = 0; k < per_thread; k += 8) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c = c * 1664525u + 1013904223u;
      const uint32_t idx = (uint32_t)(((uint64_t)mix32(c) * n_vec) >> 32);
      v[i] = ldg_nc_v4(table + idx);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) fold ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (fold == 0x12345678u) sink[t & 1023u] = fold;
}

// random 16-byte loads with the LEVEL STRUCTURE of a hash-grid decode on uniform random coordinates: every "sample" does 8 loads
// per level at random entries of THAT level's table (small coarse levels stay L2- and partly L1-resident, levels larger than the
// L2 miss in proportion), nothing else.  The like-for-like ceiling of the decode's gather for a model whose table exceeds the
// L2 (T = 2^22), independent of the product's index arithmetic.
struct ProbeLevels { uint32_t n_levels; uint32_t offset[16]; uint32_t size[16]; };      // in 16-byte entries
__global__ void __launch_bounds__(256) probe_loads_levels_kernel(const uint4* __restrict__ table, ProbeLevels lv, uint32_t n_samples, uint32_t seed,
                                                                 uint32_t* __restrict__ sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_samples) return;
  uint32_t fold = 0;
  uint32_t c = mix32(t * 0x9e3779b9u + seed);
  for (uint32_t l = 0; l < lv.n_levels; ++l) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c = c * 1664525u + 1013904223u;
      const uint32_t idx = lv.offset[l] + (uint32_t)(((uint64_t)mix32(c) * lv.size[l]) >> 32);
      v[i] = ldg_nc_v4(table + idx);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) fold ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (fold == 0x12345678u) sink[t & 1023u] = fold;
}

// loads and reductions TOGETHER (even blocks: random 16-byte loads over `table`, odd blocks: random fp16x8 reductions over `table2`):
// what the training step's gather and scatter cost when they share the memory system
__global__ void __launch_bounds__(256) probe_mixed_kernel(const uint4* __restrict__ table, __half* __restrict__ table2, uint32_t n_vec, uint32_t per_thread,
                                                          uint32_t seed, uint32_t* __restrict__ sink) {
  const uint32_t t = (blockIdx.x >> 1) * blockDim.x + threadIdx.x;
  uint32_t c = mix32(t * 0x9e3779b9u + seed + (blockIdx.x & 1u) * 77u);
  if (blockIdx.x & 1u) {
    const uint32_t one = 0x04000400u;
    for (uint32_t k = 0; k < per_thread; ++k) {
      c = c * 1664525u + 1013904223u;
      const uint32_t idx = (uint32_t)(((uint64_t)mix32(c) * n_vec) >> 32);
      asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %1, %1, %1};" ::"l"(table2 + (size_t)idx * 8), "r"(one) : "memory");
    }
    return;
  }
  uint32_t fold = 0;
  for (uint32_t k = 0; k < per_thread; k += 8) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c = c * 1664525u + 1013904223u;
      const uint32_t idx = (uint32_t)(((uint64_t)mix32(c) * n_vec) >> 32);
      v[i] = ldg_nc_v4(table + idx);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) fold ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (fold == 0x12345678u) sink[t & 1023u] = fold;
}

// stores of finished pixels into PINNED HOST memory (the zero-copy frame download): float4 per thread over a `width`-pixel-wide
// image.  pattern 0: scanline (a warp writes 512 contiguous bytes), 1: 8 x 4 pixel tiles per warp (four 128-byte runs, the
// marcher's ray order), 2: scanline with one 32-byte store per thread (two pixels), 3: tiles of 16 x 2 pixels (two 256-byte runs)
__global__ void __launch_bounds__(128) probe_host_store_kernel(float4* __restrict__ dst, uint32_t width, uint32_t height, int pattern) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t npix = width * height;
  const float4 v = make_float4((float)t, 1.f, 2.f, 3.f);
  if (pattern == 2) {
    if (2 * t + 1 >= npix) return;
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(dst + 2 * (size_t)t), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    return;
  }
  if (t >= npix) return;
  uint32_t pix = t;
  if (pattern == 1 || pattern == 3) {
    const uint32_t tw = pattern == 1 ? 8u : 16u, th = 32u / tw;
    const uint32_t warp = t >> 5, lane = t & 31u, tiles_x = width / tw;
    const uint32_t x = (warp % tiles_x) * tw + lane % tw, y = (warp / tiles_x) * th + lane / tw;
    if (y >= height) return;
    pix = y * width + x;
  }
  dst[pix] = v;
}

// streaming copy (read + write) for reference next to MEASURED_PEAKS.json's figure
__global__ void __launch_bounds__(256) probe_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n_vec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace vnr

using namespace vnr;

#define VNR_EXPORT extern "C" __attribute__((visibility("default")))

// kind 0: random 16-byte loads, 1: random 16-byte fp16x8 reductions, 2: streaming copy of table_bytes (n_ops ignored), 3: random
// 32-byte loads (n_ops loads of 32 bytes), 4..7: stores of a table_bytes frame (float4 pixels, 1024 wide) into pinned HOST memory in
// pattern kind - 4 (see probe_host_store_kernel).
// Runs `repeats` timed launches after one warm-up and returns the fastest (ms_best) and the mean (ms_mean) launch time.
VNR_EXPORT int vnr_probe_memory(int kind, size_t table_bytes, size_t n_ops, int repeats, float* ms_best, float* ms_mean) {
  if (kind < 0 || kind > 400 || table_bytes < 4096 || repeats < 1 || !ms_best) return VNR_ERR_INVALID;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return VNR_ERR_CUDA; }
  void* table = nullptr; void* aux = nullptr;
  cudaStream_t s = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = VNR_OK;
  auto ok = [&](cudaError_t e) { if (e != cudaSuccess) { cudaGetLastError(); rc = VNR_ERR_CUDA; } return e == cudaSuccess; };
  const size_t n_vec = table_bytes / 16;
  do {
    const bool host_table = kind >= 4 && kind <= 7;
    if (host_table) { if (!ok(cudaMallocHost(&table, n_vec * 16))) break; }
    else if (!ok(cudaMalloc(&table, n_vec * 16))) break;
    if (!ok(cudaMalloc(&aux, kind == 2 ? n_vec * 16 : 4096))) break;
    void* table2 = nullptr;                 // kind 8: the reductions' own table (the gradient buffer next to the parameter table)
    if (kind == 8) { if (!ok(cudaMalloc(&table2, n_vec * 16)) || !ok(cudaMemset(table2, 0, n_vec * 16))) break; }
    if (host_table) memset(table, 0, n_vec * 16);
    else if (!ok(cudaMemset(table, 0, n_vec * 16))) break;
    if (!ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking))) break;
    if (!ok(cudaEventCreate(&e0)) || !ok(cudaEventCreate(&e1))) break;
    const uint32_t per_thread = 64;                                 // the example model's loads / reductions per sample
    const size_t threads = (n_ops + per_thread - 1) / per_thread;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    float best = 1e30f, sum = 0.f;
    for (int it = 0; it <= repeats && rc == VNR_OK; ++it) {
      if (kind == 1 && !ok(cudaMemsetAsync(table, 0, n_vec * 16, s))) break;
      ok(cudaEventRecord(e0, s));
      if (kind == 0) probe_loads_kernel<<<grid, 256, 0, s>>>((const uint4*)table, (uint32_t)n_vec, per_thread, 17u + it, (uint32_t*)aux);
      else if (kind == 1) probe_reds_kernel<<<grid, 256, 0, s>>>((__half*)table, (uint32_t)n_vec, per_thread, 17u + it);
      else if (kind == 2) probe_copy_kernel<<<148 * 16, 256, 0, s>>>((const uint4*)table, (uint4*)aux, n_vec);
      else if (kind == 3) probe_loads32_kernel<<<grid, 256, 0, s>>>((const uint4*)table, (uint32_t)n_vec, per_thread, 17u + it, (uint32_t*)aux);
      else if (kind >= 9) {
        // kind 9 + 4 * i + j: threads = 256 << i (i = 0..2), dynamic shared memory = {0, 100, 160, 200} KB [j]; one CTA per SM
        // kind 100 + kb: 256 threads, kb KB of dynamic shared memory
        const int i = kind >= 100 ? 0 : (kind - 9) / 4, j = kind >= 100 ? 0 : (kind - 9) % 4;
        const unsigned threads_cta = 256u << i;
        const size_t smem = kind >= 100 ? (size_t)(kind - 100) << 10 : (size_t[]){0, 100 << 10, 160 << 10, 200 << 10}[j];
        if (i > 2 || smem > (227u << 10)) { rc = VNR_ERR_INVALID; break; }
        ok(cudaFuncSetAttribute(probe_loads_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10));
        const uint32_t pt = (uint32_t)(n_ops / ((size_t)num_sms() * threads_cta)) & ~7u;
        probe_loads_cta_kernel<<<num_sms(), threads_cta, smem, s>>>((const uint4*)table, (uint32_t)n_vec, pt, 17u + it, (uint32_t*)aux);
      }
      else if (kind == 8) probe_mixed_kernel<<<2 * grid, 256, 0, s>>>((const uint4*)table, (__half*)table2, (uint32_t)n_vec, per_thread, 17u + it, (uint32_t*)aux);
      else { const uint32_t w = 1024, h = (uint32_t)(n_vec / w); probe_host_store_kernel<<<(w * h + 127) / 128, 128, 0, s>>>((float4*)table, w, h, kind - 4); }
      ok(cudaGetLastError());
      ok(cudaEventRecord(e1, s));
      if (!ok(cudaEventSynchronize(e1))) break;
      float ms = 0.f;
      ok(cudaEventElapsedTime(&ms, e0, e1));
      if (it == 0) continue;                                        // warm-up (first touch of the table)
      best = ms < best ? ms : best; sum += ms;
    }
    *ms_best = best;
    if (ms_mean) *ms_mean = sum / (float)repeats;
    if (table2) cudaFree(table2);
  } while (0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (s) cudaStreamDestroy(s);
  if (table) { if (kind >= 4 && kind <= 7) cudaFreeHost(table); else cudaFree(table); }
  if (aux) cudaFree(aux);
  return rc;
}

// n_samples x (8 random 16-byte loads per level) over a table laid out as the levels of a hash grid (level_entries[l] entries of
// 16 bytes each, back to back): fastest and mean time of `repeats` launches after one warm-up
VNR_EXPORT int vnr_probe_levels(const uint32_t* level_entries, int n_levels, size_t n_samples, int repeats, float* ms_best, float* ms_mean) {
  if (!level_entries || n_levels < 1 || n_levels > 16 || repeats < 1 || !ms_best || n_samples < 1 || n_samples > 0xFFFFFFFFull) return VNR_ERR_INVALID;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return VNR_ERR_CUDA; }
  ProbeLevels lv; lv.n_levels = (uint32_t)n_levels;
  size_t total = 0;
  for (int l = 0; l < n_levels; ++l) { if (!level_entries[l]) return VNR_ERR_INVALID; lv.offset[l] = (uint32_t)total; lv.size[l] = level_entries[l]; total += level_entries[l]; }
  void* table = nullptr; void* aux = nullptr; cudaStream_t s = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = VNR_OK;
  auto ok = [&](cudaError_t e) { if (e != cudaSuccess) { cudaGetLastError(); rc = VNR_ERR_CUDA; } return e == cudaSuccess; };
  do {
    if (!ok(cudaMalloc(&table, total * 16)) || !ok(cudaMemset(table, 0, total * 16)) || !ok(cudaMalloc(&aux, 4096))) break;
    if (!ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) || !ok(cudaEventCreate(&e0)) || !ok(cudaEventCreate(&e1))) break;
    float best = 1e30f, sum = 0.f;
    for (int it = 0; it <= repeats && rc == VNR_OK; ++it) {
      ok(cudaEventRecord(e0, s));
      probe_loads_levels_kernel<<<(unsigned)((n_samples + 255) / 256), 256, 0, s>>>((const uint4*)table, lv, (uint32_t)n_samples, 23u + it, (uint32_t*)aux);
      ok(cudaGetLastError());
      ok(cudaEventRecord(e1, s));
      if (!ok(cudaEventSynchronize(e1))) break;
      float ms = 0.f; ok(cudaEventElapsedTime(&ms, e0, e1));
      if (it == 0) continue;
      best = ms < best ? ms : best; sum += ms;
    }
    *ms_best = best; if (ms_mean) *ms_mean = sum / (float)repeats;
  } while (0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (s) cudaStreamDestroy(s);
  if (table) cudaFree(table);
  if (aux) cudaFree(aux);
  return rc;
}

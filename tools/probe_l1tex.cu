// probe_l1tex.cu -- what does a 16-byte gather cost in the L1TEX data pipe?  Standalone measurement tool
// (nvcc -arch=sm_100a -O3 -o probe_l1tex probe_l1tex.cu).  Every variant issues the same number of 16-byte
// loads per thread from an L2-resident 46.7 MB table; only the address pattern inside a warp instruction
// differs.  Prints loads/s; run under ncu with --metrics l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,
// l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum to get wavefronts and sectors per load.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// MODE 0: every lane an independent random entry
// MODE 1: lane pairs share a 32-byte sector (entry 2k, 2k+1), pairs random
// MODE 2: lane quads share 64 bytes
// MODE 3: 8 lanes share one 128-byte line
// MODE 4: every lane random, but the lane's second load of an iteration pair hits the same sector as its first
// MODE 5: lane pairs share a sector half of the time (the hash-grid x-pair statistics), else random
// MODE 6: like 0 with 8-byte loads (two per 16 bytes)
template <int MODE>
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ tab, uint32_t mask, int iters, uint32_t* __restrict__ out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t fold = 0;
  for (int it = 0; it < iters; it += 8) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t key = (uint32_t)(it + k) * 0x9e3779b9u;
      uint32_t idx;
      if (MODE == 0 || MODE == 6) idx = mix(tid ^ key);
      else if (MODE == 1) idx = (mix((tid >> 1) ^ key) << 1) | (lane & 1);
      else if (MODE == 2) idx = (mix((tid >> 2) ^ key) << 2) | (lane & 3);
      else if (MODE == 3) idx = (mix((tid >> 3) ^ key) << 3) | (lane & 7);
      else if (MODE == 4) idx = (mix(tid ^ ((uint32_t)(it + (k & ~1)) * 0x9e3779b9u)) & ~1u) | (k & 1);
      else { const uint32_t h = mix((tid >> 1) ^ key); idx = (h & 0x80000000u) ? ((h << 1) | (lane & 1)) : mix(tid ^ key); }
      if (MODE == 6) {
        const uint2* t2 = reinterpret_cast<const uint2*>(tab);
        const uint2 a = __ldg(t2 + ((idx & mask) << 1)), b = __ldg(t2 + ((mix(idx) & mask) << 1));
        v[k] = make_uint4(a.x, a.y, b.x, b.y);
      } else {
        v[k] = __ldg(tab + (idx & mask));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) fold ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  }
  out[tid] = fold;
}

template <int MODE>
static void run(const char* name, const uint4* tab, uint32_t mask, uint32_t* out) {
  const int blocks = 148 * 8 * 4, iters = 512;
  gather<MODE><<<blocks, 256>>>(tab, mask, 64, out);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  gather<MODE><<<blocks, 256>>>(tab, mask, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double loads = (double)blocks * 256 * iters * (MODE == 6 ? 2 : 1);
  printf("%-44s %8.3f ms  %7.2f G loads/s  (%6.2f G 16-byte units/s)\n", name, ms, loads / ms / 1e6, (double)blocks * 256 * iters / ms / 1e6);
}

int main() {
  const uint32_t n = 1u << 21;            // 2 Mi entries x 16 B = 32 MiB (L2 resident)
  uint4* tab; uint32_t* out;
  cudaMalloc(&tab, (size_t)n * 16); cudaMemset(tab, 1, (size_t)n * 16);
  cudaMalloc(&out, (size_t)148 * 8 * 4 * 256 * 4);
  run<0>("0 lanes independent", tab, n - 1, out);
  run<1>("1 lane pairs share a sector", tab, n - 1, out);
  run<2>("2 lane quads share 64 B", tab, n - 1, out);
  run<3>("3 eight lanes share a line", tab, n - 1, out);
  run<4>("4 same-lane consecutive loads share a sector", tab, n - 1, out);
  run<5>("5 lane pairs share a sector half the time", tab, n - 1, out);
  run<6>("6 independent 8-byte loads", tab, n - 1, out);
  return 0;
}

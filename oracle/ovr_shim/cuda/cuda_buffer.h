// ovr_shim/cuda/cuda_buffer.h -- stand-in for OVR's cuda/cuda_buffer.h + cuda/cuda_math.h + cuda/texture.h + util kernels, as far
// as the reference's marcher sources use them: CUDA_CHECK / CUDA_SYNC_CHECK, CUDABuffer (a cudaMalloc'ed byte buffer with
// resize / d_pointer), util::linear_kernel / bilinear_kernel (one thread per element, 1-D / 2-D launch) and the texture helpers
// array.h names.  TEST INFRASTRUCTURE (oracle/ref_marcher); nothing in the product includes this.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include <sstream>
#include <iostream>
#include <type_traits>

#define CUDA_CHECK(call)                                                                                          \
  do {                                                                                                            \
    cudaError_t rc_ = (call);                                                                                     \
    if (rc_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(rc_) + " at " #call); \
  } while (0)
#define CUDA_CHECK_NOEXCEPT(call) (void)(call)
#define CUDA_SYNC_CHECK()                                                                                         \
  do {                                                                                                            \
    cudaDeviceSynchronize();                                                                                      \
    cudaError_t rc_ = cudaGetLastError();                                                                         \
    if (rc_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(rc_));      \
  } while (0)

#define GDT_TERMINAL_RED ""
#define GDT_TERMINAL_RESET ""
#define GDT_TERMINAL_GREEN ""
#define GDT_TERMINAL_YELLOW ""

namespace util {
inline size_t& total_n_bytes_allocated() { static size_t n = 0; return n; }
inline std::string prettyBytes(size_t n) { std::ostringstream s; s << n << " B"; return s.str(); }

template <typename T> inline T div_round_up(T a, T b) { return (a + b - 1) / b; }
// one thread per element; the kernel's first parameter(s) are the element count(s)
template <typename... Types, typename... Args>
inline void linear_kernel(void (*kernel)(uint32_t, Types...), uint32_t shmem, cudaStream_t stream, uint32_t n, Args... args) {
  if (n == 0) return;
  const uint32_t threads = 128, blocks = (n + threads - 1) / threads;
  kernel<<<blocks, threads, shmem, stream>>>(n, (Types)args...);
}
template <typename... Types, typename... Args>
inline void bilinear_kernel(void (*kernel)(uint32_t, uint32_t, Types...), uint32_t shmem, cudaStream_t stream, uint32_t w, uint32_t h, Args... args) {
  if (w == 0 || h == 0) return;
  const dim3 threads(16, 8), blocks((w + 15) / 16, (h + 7) / 8);
  kernel<<<blocks, threads, shmem, stream>>>(w, h, (Types)args...);
}
}  // namespace util

struct CUDABuffer {
  void* d_ptr = nullptr; size_t sizeInBytes = 0; bool external = false;
  void set_external(CUDABuffer& o) { free(); d_ptr = o.d_ptr; sizeInBytes = o.sizeInBytes; external = true; }
  void memset(int value, cudaStream_t s) { if (d_ptr) CUDA_CHECK(cudaMemsetAsync(d_ptr, value, sizeInBytes, s)); }
  size_t size() const { return sizeInBytes; }
  void* d_pointer() const { return d_ptr; }
  void alloc(size_t n, cudaStream_t = 0) { free(); if (n) CUDA_CHECK(cudaMalloc(&d_ptr, n)); sizeInBytes = n; }
  void resize(size_t n, cudaStream_t s = 0) { if (n != sizeInBytes) alloc(n, s); }
  void free(cudaStream_t = 0) { if (d_ptr && !external) cudaFree(d_ptr); d_ptr = nullptr; sizeInBytes = 0; external = false; }
  void nullify(cudaStream_t s = 0) { if (d_ptr) CUDA_CHECK(cudaMemsetAsync(d_ptr, 0, sizeInBytes, s)); }
  template <typename T> void upload(const T* t, size_t count) { CUDA_CHECK(cudaMemcpy(d_ptr, t, count * sizeof(T), cudaMemcpyHostToDevice)); }
  template <typename T> void download(T* t, size_t count) { CUDA_CHECK(cudaMemcpy(t, d_ptr, count * sizeof(T), cudaMemcpyDeviceToHost)); }
  template <typename T> void alloc_and_upload(const std::vector<T>& v) { alloc(v.size() * sizeof(T)); upload(v.data(), v.size()); }
};

// texture helpers named by core/array.h (normalized coordinates, clamp addressing: what the reference's lookups assume,
// raytracing.h:71-81,105-110)
template <typename T>
inline cudaTextureObject_t createCudaTexture(cudaArray_t array, cudaTextureReadMode read_mode, cudaTextureFilterMode filter,
                                             cudaTextureFilterMode = cudaFilterModeLinear) {
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = array;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = filter; td.readMode = read_mode; td.normalizedCoords = 1;
  cudaTextureObject_t tex = 0;
  CUDA_CHECK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  return tex;
}
template <typename T> inline cudaArray_t allocateCudaArray1D(const T* data, size_t n) {
  cudaChannelFormatDesc desc = cudaCreateChannelDesc<T>();
  cudaArray_t a = nullptr;
  CUDA_CHECK(cudaMallocArray(&a, &desc, n, 0));
  CUDA_CHECK(cudaMemcpy2DToArray(a, 0, 0, data, n * sizeof(T), n * sizeof(T), 1, cudaMemcpyHostToDevice));
  return a;
}
template <typename T> inline void fillCudaArray1D(cudaArray_t a, const T* data, size_t n) {
  CUDA_CHECK(cudaMemcpy2DToArray(a, 0, 0, data, n * sizeof(T), n * sizeof(T), 1, cudaMemcpyHostToDevice));
}

// Can the GPU pull slabs of a file straight out of the page cache?  mmap the file, cudaHostRegister the mapping, read it from a
// kernel (zero-copy loads over PCIe) -- the refresh path of the out-of-core sampler without a CPU copy.  (GPU box; measurement only)
//   nvcc -O3 -arch=sm_100a -o tools/_build/probe_mmap tools/probe_mmap.cu && tools/_build/probe_mmap /tmp/file.raw
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__global__ void pull_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const uint32_t* __restrict__ slab_first, uint32_t vec_per_slab, uint32_t n_slabs) {
  const size_t total = (size_t)vec_per_slab * n_slabs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t s = (uint32_t)(i / vec_per_slab), k = (uint32_t)(i % vec_per_slab);
    dst[i] = src[(size_t)slab_first[s] + k];
  }
}

__global__ void spin_kernel(long long cycles) {
  extern __shared__ char sm[];
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) { }
  if (cycles < 0) sm[0] = 1;
}

int main(int ac, char** av) {
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10);
  const char* path = ac > 1 ? av[1] : "/tmp/vnr_probe_mmap.raw";
  const size_t bytes = (size_t)1 << 30;
  int fd = open(path, O_RDONLY);
  if (fd < 0 || lseek(fd, 0, SEEK_END) < (off_t)bytes) {
    if (fd >= 0) close(fd);
    fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    std::vector<char> buf(1 << 24, 7);
    for (size_t o = 0; o < bytes; o += buf.size()) if (write(fd, buf.data(), buf.size()) != (ssize_t)buf.size()) { perror("write"); return 1; }
    close(fd); fd = open(path, O_RDONLY);
  }
  const uint32_t n_slabs = 1024, slab_bytes = 104448, vec = slab_bytes / 16;
  std::vector<uint32_t> first(n_slabs);
  srand(1);
  for (auto& f : first) f = (uint32_t)(((size_t)rand() % ((bytes - slab_bytes) / 4096)) * 4096 / 16);
  uint32_t* d_first; uint4* d_dst;
  cudaMalloc(&d_first, n_slabs * 4); cudaMemcpy(d_first, first.data(), n_slabs * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&d_dst, (size_t)n_slabs * slab_bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct { const char* name; int prot, mflags; unsigned rflags; } modes[] = {
    {"MAP_SHARED  + Mapped|ReadOnly", PROT_READ, MAP_SHARED, cudaHostRegisterMapped | cudaHostRegisterReadOnly},
    {"MAP_PRIVATE + Mapped|ReadOnly", PROT_READ, MAP_PRIVATE, cudaHostRegisterMapped | cudaHostRegisterReadOnly},
    {"MAP_PRIVATE rw + Mapped", PROT_READ | PROT_WRITE, MAP_PRIVATE, cudaHostRegisterMapped},
    {"MAP_SHARED|POPULATE + Mapped|ReadOnly", PROT_READ, MAP_SHARED | MAP_POPULATE, cudaHostRegisterMapped | cudaHostRegisterReadOnly},
  };
  for (auto& m : modes) {
    void* p = mmap(nullptr, bytes, m.prot, m.mflags, fd, 0);
    if (p == MAP_FAILED) { printf("%-40s mmap failed\n", m.name); continue; }
    cudaError_t e = cudaHostRegister(p, bytes, m.rflags);
    if (e != cudaSuccess) { printf("%-40s cudaHostRegister: %s\n", m.name, cudaGetErrorString(e)); cudaGetLastError(); munmap(p, bytes); continue; }
    void* dp = nullptr; e = cudaHostGetDevicePointer(&dp, p, 0);
    if (e != cudaSuccess) { printf("%-40s cudaHostGetDevicePointer: %s\n", m.name, cudaGetErrorString(e)); cudaGetLastError(); cudaHostUnregister(p); munmap(p, bytes); continue; }
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
      cudaEventRecord(e0);
      pull_kernel<<<148 * 8, 256>>>((const uint4*)dp, d_dst, d_first, vec, n_slabs);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    e = cudaGetLastError();
    char chk[16]; cudaMemcpy(chk, d_dst, 16, cudaMemcpyDeviceToHost);
    printf("%-40s ok: %u slabs x %u B pulled in %.3f ms = %.1f GB/s (%s, first byte %d)\n", m.name, n_slabs, slab_bytes, best, (double)n_slabs * slab_bytes / best / 1e6,
           cudaGetErrorString(e), (int)chk[0]);
    // the same slabs through the copy engines: ONE cudaMemcpyBatchAsync of 3 slices per slab (no SM involved: can it run under a
    // kernel that fills every SM?)
    {
      const size_t slice = slab_bytes / 3, n_copies = (size_t)n_slabs * 3;
      std::vector<void*> dsts(n_copies), srcs(n_copies); std::vector<size_t> sizes(n_copies, slice);
      for (size_t i = 0; i < n_copies; ++i) {
        srcs[i] = (char*)p + (size_t)first[i / 3] * 16 + (i % 3) * ((size_t)1 << 20);      // slices of a slab are a plane apart in the file
        if ((char*)srcs[i] + slice > (char*)p + bytes) srcs[i] = (char*)p + (size_t)first[i / 3] * 16;
        dsts[i] = (char*)d_dst + i * slice;
      }
      cudaMemcpyAttributes at = {}; at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
      size_t idx0 = 0, fail = 0;
      cudaStream_t cs, ks; cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking);
      for (int busy = 0; busy < 2; ++busy) {
        float bestb = 1e30f; cudaError_t eb = cudaSuccess;
        for (int it = 0; it < 4; ++it) {
          if (busy) spin_kernel<<<148, 1024, 200 << 10, ks>>>(6000000);          // ~3 ms on every SM
          cudaEventRecord(e0, cs);
          eb = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), n_copies, &at, &idx0, 1, &fail, cs);
          cudaEventRecord(e1, cs); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < bestb) bestb = ms;
          cudaDeviceSynchronize();
        }
        printf("%-40s cudaMemcpyBatchAsync of %zu x %zu B%s: %s, %.3f ms = %.1f GB/s\n", m.name, n_copies, slice, busy ? " under a kernel on every SM" : "", cudaGetErrorString(eb), bestb,
               (double)n_copies * slice / bestb / 1e6);
        cudaGetLastError();
      }
      cudaStreamDestroy(cs); cudaStreamDestroy(ks);
    }
    cudaHostUnregister(p); munmap(p, bytes);
  }
  // reference point: pinned staging + one cudaMemcpyAsync of the same bytes
  void* h; cudaMallocHost(&h, (size_t)n_slabs * slab_bytes);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) { cudaEventRecord(e0); cudaMemcpyAsync(d_dst, h, (size_t)n_slabs * slab_bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  printf("pinned staging, one cudaMemcpyAsync:      %.3f ms = %.1f GB/s\n", best, (double)n_slabs * slab_bytes / best / 1e6);
  return 0;
}

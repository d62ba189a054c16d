// ingest.cu -- ground-truth volumes from raw files, streamed into HBM.
//
// Replaces what the reference needs three samplers for (core/samplers/neural_sampler.cpp):
// StaticSampler::load (:223-288: read the whole file, min/max, convert_volume to normalised floats :176-210,
// upload), and the OutOfCoreSampler / VirtualMemorySampler (:488-1191) that exist because a 24 GB GPU and the
// host RAM cannot hold a large volume.  On B200 the whole volume is resident in HBM (1024^3 float = 4 GiB,
// 2048^3 = 32 GiB of 180 GB), so "out of core" becomes a streaming LOAD: a reader thread fills pinned
// double buffers with pread(), the device converts each chunk ((float)v - vmin) / (vmax - vmin) clamped to
// [0,1] (convert_volume) into the resident float volume, and training samples in-core afterwards with the
// static sampler's arithmetic.  When no value range is given the file is streamed twice (min/max, then
// convert), as the reference does in memory.
#include <condition_variable>
#include <cstdio>
#include <fcntl.h>
#include <mutex>
#include <thread>
#include <unistd.h>

#include "train.h"
#include "volume.h"

namespace vnr {

enum RawType { RAW_U8 = 0, RAW_I8, RAW_U16, RAW_I16, RAW_U32, RAW_I32, RAW_U64, RAW_I64, RAW_F32, RAW_F32x2, RAW_F32x3, RAW_F32x4, RAW_F64 };   // core/mathdef.h:51-65

static size_t raw_size(int t) {
  switch (t) {
    case RAW_U8: case RAW_I8: return 1;
    case RAW_U16: case RAW_I16: return 2;
    case RAW_U32: case RAW_I32: case RAW_F32: return 4;
    case RAW_F64: return 8;
    default: throw UnsupportedError("unsupported voxel type (scalar uint8/int8/uint16/int16/uint32/int32/float/double only)");
  }
}

template <typename T> __device__ __forceinline__ T load_raw(const uint8_t* p, size_t i, bool swap) {
  T v;
  uint8_t b[sizeof(T)];
#pragma unroll
  for (int k = 0; k < (int)sizeof(T); ++k) b[k] = p[i * sizeof(T) + (swap ? sizeof(T) - 1 - k : k)];
  memcpy(&v, b, sizeof(T));
  return v;
}

template <typename T>
__global__ void raw_minmax_kernel(const uint8_t* __restrict__ raw, size_t n, bool swap, float* __restrict__ minmax) {
  float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = (float)load_raw<T>(raw, i, swap);
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) {
    // ordered-int trick for float atomics of either sign
    if (mn >= 0.f) atomicMin((int*)&minmax[0], __float_as_int(mn)); else atomicMax((unsigned*)&minmax[0], __float_as_uint(mn));
    if (mx >= 0.f) atomicMax((int*)&minmax[1], __float_as_int(mx)); else atomicMin((unsigned*)&minmax[1], __float_as_uint(mx));
  }
}

template <typename T>
__global__ void raw_convert_kernel(const uint8_t* __restrict__ raw, size_t n, bool swap, float vmin, float vmax, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __fdiv_rn((float)load_raw<T>(raw, i, swap) - vmin, vmax - vmin);      // convert_volume :176-210
  out[i] = fminf(fmaxf(v, 0.f), 1.f);
}

template <typename T>
static void launch_chunk(bool minmax_pass, const uint8_t* d_raw, size_t n, bool swap, float vmin, float vmax, float* d_out, float* d_minmax, cudaStream_t s) {
  if (minmax_pass) raw_minmax_kernel<T><<<1184, 256, 0, s>>>(d_raw, n, swap, d_minmax);
  else raw_convert_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_raw, n, swap, vmin, vmax, d_out);
}

static void dispatch_chunk(int type, bool minmax_pass, const uint8_t* d_raw, size_t n, bool swap, float vmin, float vmax, float* d_out, float* d_minmax, cudaStream_t s) {
  switch (type) {
    case RAW_U8: launch_chunk<uint8_t>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_I8: launch_chunk<int8_t>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_U16: launch_chunk<uint16_t>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_I16: launch_chunk<int16_t>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_U32: launch_chunk<uint32_t>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_I32: launch_chunk<int32_t>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_F32: launch_chunk<float>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    case RAW_F64: launch_chunk<double>(minmax_pass, d_raw, n, swap, vmin, vmax, d_out, d_minmax, s); break;
    default: throw UnsupportedError("unsupported voxel type");
  }
  VNR_CUDA(cudaGetLastError());
}

// One streaming pass over the file: reader thread -> pinned double buffers -> H2D -> kernel.
static void stream_file(int fd, uint64_t offset, size_t count, int type, bool swap, bool minmax_pass, float vmin, float vmax,
                        float* d_out, float* d_minmax, cudaStream_t s) {
  const size_t esz = raw_size(type);
  const size_t chunk_elems = std::min<size_t>(count, ((size_t)64 << 20) / esz);
  uint8_t* h_buf[2] = {nullptr, nullptr};
  uint8_t* d_buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  std::string error;
  try {
    for (int k = 0; k < 2; ++k) {
      VNR_CUDA(cudaMallocHost((void**)&h_buf[k], chunk_elems * esz));
      VNR_CUDA(cudaMalloc((void**)&d_buf[k], chunk_elems * esz));
      VNR_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    }
    const size_t n_chunks = (count + chunk_elems - 1) / chunk_elems;
    std::mutex m; std::condition_variable cv;
    int filled[2] = {-1, -1};        // chunk index held by each buffer (-1: free)
    bool failed = false;
    std::thread reader([&] {
      for (size_t c = 0; c < n_chunks; ++c) {
        const int k = (int)(c & 1);
        { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return filled[k] < 0 || failed; }); if (failed) return; }
        const size_t first = c * chunk_elems, n = std::min(chunk_elems, count - first);
        size_t got = 0; const size_t want = n * esz;
        while (got < want) {
          const ssize_t r = pread(fd, h_buf[k] + got, want - got, (off_t)(offset + first * esz + got));
          if (r <= 0) { std::lock_guard<std::mutex> lk(m); failed = true; error = "volume file is shorter than dims * voxel size"; cv.notify_all(); return; }
          got += (size_t)r;
        }
        { std::lock_guard<std::mutex> lk(m); filled[k] = (int)c; }
        cv.notify_all();
      }
    });
    try {
      for (size_t c = 0; c < n_chunks; ++c) {
        const int k = (int)(c & 1);
        { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return filled[k] == (int)c || failed; }); if (failed) break; }
        const size_t first = c * chunk_elems, n = std::min(chunk_elems, count - first);
        VNR_CUDA(cudaMemcpyAsync(d_buf[k], h_buf[k], n * esz, cudaMemcpyHostToDevice, s));
        dispatch_chunk(type, minmax_pass, d_buf[k], n, swap, vmin, vmax, d_out ? d_out + first : nullptr, d_minmax, s);
        VNR_CUDA(cudaEventRecord(done[k], s));
        // the pinned buffer may be refilled once its copy has completed (the kernel reads the device copy)
        VNR_CUDA(cudaEventSynchronize(done[k]));
        { std::lock_guard<std::mutex> lk(m); filled[k] = -1; }
        cv.notify_all();
      }
    } catch (...) {
      { std::lock_guard<std::mutex> lk(m); failed = true; }
      cv.notify_all(); reader.join(); throw;
    }
    reader.join();
    if (failed) throw InvalidError(error.empty() ? "reading the volume file failed" : error);
  } catch (...) {
    for (int k = 0; k < 2; ++k) { if (h_buf[k]) cudaFreeHost(h_buf[k]); if (d_buf[k]) cudaFree(d_buf[k]); if (done[k]) cudaEventDestroy(done[k]); }
    throw;
  }
  for (int k = 0; k < 2; ++k) { cudaFreeHost(h_buf[k]); cudaFree(d_buf[k]); cudaEventDestroy(done[k]); }
}

void load_groundtruth_file(Volume* v, const char* path, int type, uint64_t offset, bool big_endian, float vmin, float vmax, float* range_out) {
  raw_size(type);
  const size_t count = (size_t)v->dims[0] * v->dims[1] * v->dims[2];
  const int fd = open(path, O_RDONLY);
  if (fd < 0) throw InvalidError(std::string("cannot open volume file ") + path);
  try {
    cudaStream_t s = v->stream;
    if (!(vmax > vmin)) {        // range not given: compute it from the data (StaticSampler::load :248-262)
      DevBuf<float> mm; mm.alloc(2);
      const float init[2] = {3.402823466e+38f, -3.402823466e+38f};
      VNR_CUDA(cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, s));
      stream_file(fd, offset, count, type, big_endian, true, 0.f, 1.f, nullptr, mm.p, s);
      float h[2];
      VNR_CUDA(cudaMemcpyAsync(h, mm.p, sizeof h, cudaMemcpyDeviceToHost, s));
      VNR_CUDA(cudaStreamSynchronize(s));
      vmin = h[0]; vmax = h[1];
      if (!(vmax > vmin)) throw InvalidError("volume has an empty value range");
    }
    v->gt.alloc(count);
    stream_file(fd, offset, count, type, big_endian, false, vmin, vmax, v->gt.p, nullptr, s);
    VNR_CUDA(cudaStreamSynchronize(s));
    v->have_gt = true;
    if (range_out) { range_out[0] = vmin; range_out[1] = vmax; }
  } catch (...) { close(fd); throw; }
  close(fd);
}

}  // namespace vnr

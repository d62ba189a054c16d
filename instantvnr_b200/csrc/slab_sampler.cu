// slab_sampler.cu -- the out-of-core training sampler: a pool of random slabs of the volume file resident in HBM,
// refreshed every step, sampled on the device.
//
// Replaces OutOfCoreSampler::sample (core/samplers/neural_sampler.cpp:1065-1120) and its RandomBuffer (:488-668).
// The reference keeps NUM_BLOCKS (65 536) slabs in HOST memory -- a slab is the full x-row x ceil(32 KiB / row bytes)
// y-rows x 1 z-slice, plus one ghost row / slice on every side -- replaces NUM_CONCURRENT_BLOCKS (1024) of them per
// step through libaio, and evaluates every training sample on the CPU with TBB (pick a random slab, a random voxel of
// it, a random point in that voxel's cell, trilinear interpolation of the values normalised BEFORE interpolating),
// then uploads 16 B per sample.  Here the pool lives in HBM in the file's own scalar type (the 64 x 1024 default slabs
// of a 1024^3 float volume are 7.9 GB of the 180 GB); a step uploads the refreshed slabs (pread workers -> pinned
// double buffer -> two cudaMemcpyAsync) and ONE kernel draws and interpolates the batch where it is consumed.  The
// per-sample arithmetic (index selection, coordinate, trilinear_vkl :302-329) is restated exactly; the uniforms come
// from the sampler's pcg32 stream (five per sample) instead of the reference's unseeded std::mt19937.
//
// Refresh path, round 2: the GPU PULLS the refreshed slabs out of the page cache itself.  The file is mmap'ed (private,
// writable -- the one combination cudaHostRegister accepts for file-backed pages here) and registered as mapped pinned
// memory; a step's refresh is one kernel whose zero-copy loads fetch the 1024 random slabs over PCIe (107 MB in 2.1 ms =
// 51 GB/s, tools/probe_mmap.cu) with no CPU copy and no staging buffer.  The pread path read the same bytes out of the page
// cache into pinned staging with up to 16 host threads PER RANK: host-bound, and slower the more ranks shared the host (1 / 2 /
// 8 GPUs: 358 / 219 / 66 steps/s at 1024^3).  It remains the fallback when the registration fails (a file larger than what can
// be pinned) or with VNR_OOC_PULL=0.
#include <atomic>
#include <cstdlib>
#include <fcntl.h>
#include <memory>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>

#include "train.h"
#include "volume.h"

namespace vnr {

enum { SLAB_U8 = 0, SLAB_I8, SLAB_U16, SLAB_I16, SLAB_U32, SLAB_I32, SLAB_F32 = 8, SLAB_F64 = 12 };   // ValueType, core/mathdef.h:51-65

static size_t slab_elem_size(int t) {
  switch (t) {
    case SLAB_U8: case SLAB_I8: return 1;
    case SLAB_U16: case SLAB_I16: return 2;
    case SLAB_U32: case SLAB_I32: case SLAB_F32: return 4;
    case SLAB_F64: return 8;
    default: throw UnsupportedError("unsupported voxel type (scalar uint8/int8/uint16/int16/uint32/int32/float/double only)");
  }
}

// RandomBuffer::Block (:497-503), reduced to what the sampling kernel reads
struct SlabDesc {
  uint64_t first_voxel;      // flattened file index of the slab's first (non-ghost) voxel
  uint32_t length;           // voxels in the slab proper
  int gy0, gz0, gny, gnz;    // ghost-extended bounds: lower y / z and extent in y / z (x is always the full row)
};

struct SlabSampler {
  int fd = -1;
  uint64_t file_offset = 0;
  int type = SLAB_F32; size_t elem = 4;
  int dims[3] = {0, 0, 0};
  float vmin = 0.f, vmax = 1.f;
  int block_rows = 1;                    // block_dims.y (:541)
  int nby = 1, nbz = 1;                  // block_index_space (:548-550)
  size_t slot_bytes = 0;                 // block_size_aligned (:560)
  uint32_t n_slots = 0, n_refresh = 0;   // NUM_BLOCKS, NUM_CONCURRENT_BLOCKS
  DevBuf<uint8_t> pool;
  DevBuf<SlabDesc> table;
  std::vector<SlabDesc> h_table;
  uint8_t* h_stage[2] = {nullptr, nullptr};
  SlabDesc* h_desc[2] = {nullptr, nullptr};
  cudaEvent_t copied[2] = {nullptr, nullptr};
  int cur = 0;
  bool pending = false; uint32_t pending_first = 0;
  std::vector<std::thread> workers;
  std::atomic<int> io_error{0};
  Pcg32 host_rng;                        // slab selection (the reference: std::mt19937, neural_sampler.cu:62-75)
  int rank = 0;                          // data-parallel rank: selects the pcg32 stream of host_rng
  uint64_t bytes_uploaded = 0;
  // pull mode: the file, mapped and registered; its device-side address
  void* map = nullptr; size_t map_bytes = 0; const uint8_t* map_dev = nullptr;
  // ... and its own stream: the pull of the NEXT refresh starts as soon as a batch has been drawn and runs under the training
  // step of that batch (which does not touch the pool); the next draw waits for it
  cudaStream_t pull_stream = nullptr; cudaEvent_t ev_drawn = nullptr, ev_pulled = nullptr; bool pull_in_flight = false;
  void pull_slabs(uint32_t first, uint32_t count, cudaStream_t s);      // slots first .. first + count - 1 <- their slabs' bytes of the file

  ~SlabSampler() {
    for (auto& t : workers) if (t.joinable()) t.join();
    for (int k = 0; k < 2; ++k) { if (h_stage[k]) cudaFreeHost(h_stage[k]); if (h_desc[k]) cudaFreeHost(h_desc[k]); if (copied[k]) cudaEventDestroy(copied[k]); }
    if (map) { cudaDeviceSynchronize(); cudaHostUnregister(map); munmap(map, map_bytes); }
    if (pull_stream) cudaStreamDestroy(pull_stream);
    if (ev_drawn) cudaEventDestroy(ev_drawn);
    if (ev_pulled) cudaEventDestroy(ev_pulled);
    if (fd >= 0) close(fd);
  }

  // finish (and drop) a refresh that is still being read
  void drain() {
    for (auto& t : workers) if (t.joinable()) t.join();
    workers.clear();
    pending = false;
    if (pull_stream) cudaStreamSynchronize(pull_stream);
    pull_in_flight = false;
  }

  // preload every slot (:571-577), n_refresh at a time (the last group wraps around when n_slots % n_refresh != 0), and
  // start the first refresh (:579)
  void preload(cudaStream_t s) {
    host_rng.seed(1337, 2 + (uint64_t)rank);
    for (uint32_t i = 0; i < n_slots; i += n_refresh) {
      submit((int64_t)i);
      wait_and_upload(s);
    }
    VNR_CUDA(cudaStreamSynchronize(s));
    submit(-1);
  }

  // submit_one_job (:588-646): describe block (by, bz) and read its ghost-extended rows into `dst`
  SlabDesc describe(int by, int bz) const {
    const int y0 = by * block_rows, y1 = std::min(y0 + block_rows, dims[1]);
    const int z0 = bz, z1 = std::min(z0 + 1, dims[2]);
    SlabDesc d;
    d.first_voxel = ((uint64_t)z0 * dims[1] + (uint64_t)y0) * dims[0];
    d.length = (uint32_t)((uint64_t)dims[0] * (y1 - y0) * (z1 - z0));
    d.gy0 = std::max(y0 - 1, 0); d.gz0 = std::max(z0 - 1, 0);
    d.gny = std::min(y1 + 1, dims[1]) - d.gy0; d.gnz = std::min(z1 + 1, dims[2]) - d.gz0;
    return d;
  }
  void read_block(const SlabDesc& d, uint8_t* dst) {
    const size_t slice_bytes = (size_t)dims[0] * d.gny * elem;
    for (int z = 0; z < d.gnz; ++z) {
      const uint64_t off = file_offset + (((uint64_t)(d.gz0 + z) * dims[1] + (uint64_t)d.gy0) * dims[0]) * elem;
      size_t got = 0;
      while (got < slice_bytes) {
        const ssize_t r = pread(fd, dst + z * slice_bytes + got, slice_bytes - got, (off_t)(off + got));
        if (r <= 0) { io_error = 1; return; }
        got += (size_t)r;
      }
    }
  }

  // submit_all_jobs (:648-656): n_refresh consecutive slots starting at `first` get new random blocks; the reads run on
  // worker threads into the staging buffer `cur ^ 1` while the GPU works on the current step
  void submit(int64_t first_or_neg) {
    const int k = cur ^ 1;
    VNR_CUDA(cudaEventSynchronize(copied[k]));                 // the staging buffer's previous upload has left the host
    const uint32_t first = first_or_neg < 0 ? host_rng.next_uint() % n_slots : (uint32_t)first_or_neg;
    for (uint32_t j = 0; j < n_refresh; ++j) {
      const uint32_t b = host_rng.next_uint() % (uint32_t)(nby * nbz);      // random_grid_index(block_index_space)
      h_desc[k][j] = describe((int)(b % (uint32_t)nby), (int)(b / (uint32_t)nby));
    }
    pending = true; pending_first = first;
    if (map_dev) return;                                       // pull mode: nothing to read on the host
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const uint32_t nt = std::min<uint32_t>(hw, n_refresh);
    workers.clear();
    for (uint32_t t = 0; t < nt; ++t)
      workers.emplace_back([this, k, t, nt] { for (uint32_t j = t; j < n_refresh; j += nt) read_block(h_desc[k][j], h_stage[k] + (size_t)j * slot_bytes); });
    pending = true; pending_first = first;
  }

  // wait_all_jobs (:658-661) + upload: slots first .. first + n_refresh - 1 (mod n_slots) <- staging buffer
  void wait_and_upload(cudaStream_t s) {
    if (!pending) return;
    for (auto& t : workers) if (t.joinable()) t.join();
    workers.clear();
    if (io_error) throw InvalidError("reading the volume file failed (shorter than dims * voxel size?)");
    const int k = cur ^ 1;
    const uint32_t first = pending_first, n0 = std::min(n_refresh, n_slots - first), n1 = n_refresh - n0;
    if (!map_dev) VNR_CUDA(cudaMemcpyAsync(pool.p + (size_t)first * slot_bytes, h_stage[k], (size_t)n0 * slot_bytes, cudaMemcpyHostToDevice, s));
    VNR_CUDA(cudaMemcpyAsync(table.p + first, h_desc[k], (size_t)n0 * sizeof(SlabDesc), cudaMemcpyHostToDevice, s));
    if (n1) {
      if (!map_dev) VNR_CUDA(cudaMemcpyAsync(pool.p, h_stage[k] + (size_t)n0 * slot_bytes, (size_t)n1 * slot_bytes, cudaMemcpyHostToDevice, s));
      VNR_CUDA(cudaMemcpyAsync(table.p, h_desc[k] + n0, (size_t)n1 * sizeof(SlabDesc), cudaMemcpyHostToDevice, s));
    }
    if (map_dev) { pull_slabs(first, n0, s); if (n1) pull_slabs(0, n1, s); }
    VNR_CUDA(cudaEventRecord(copied[k], s));
    for (uint32_t j = 0; j < n_refresh; ++j) h_table[(first + j) % n_slots] = h_desc[k][j];
    bytes_uploaded += (uint64_t)n_refresh * slot_bytes;
    pending = false;
    cur = k;
  }
};

// Pull mode: persistent blocks of 1024 threads copy the slices of the refreshed slabs from the mapped file (zero-copy loads over
// PCIe, 16 bytes per thread where source, destination and length allow, bytes otherwise) into the pool.  The pull of the next
// refresh is enqueued behind the draw of a batch, on its own stream, and runs under that batch's training step.  Measured at
// 1024^3 / 1024 slabs per step on one B200 (steps/s against the number of pull blocks, VNR_OOC_PULL_SMS): 4: 235, 8: 354, 16: 396,
// 32: 420, 148: 426 -- the link wants many loads in flight; the fused training kernel (one CTA per SM, whole register file) runs
// on the SMs the pull leaves free and finishes its remaining CTAs afterwards.  The step is then 2.35 ms of which the pull is
// 2.1 (51 GB/s): PCIe-bound.
constexpr unsigned kPullBlocks = 64;
__global__ void __launch_bounds__(1024) slab_pull_kernel(const uint8_t* __restrict__ file, uint64_t file_offset, size_t elem, int3 dims,
                                                         const SlabDesc* __restrict__ table, uint32_t first, uint32_t count, uint8_t* __restrict__ pool, size_t slot_bytes) {
  for (uint32_t item = blockIdx.x; item < count * 3u; item += gridDim.x) {
    const uint32_t j = item / 3u; const int z = (int)(item % 3u);
    const SlabDesc d = table[first + j];
    if (z >= d.gnz) continue;
    const size_t slice_bytes = (size_t)dims.x * (size_t)d.gny * elem;
    const uint8_t* __restrict__ src = file + file_offset + (((uint64_t)(d.gz0 + z) * (uint64_t)dims.y + (uint64_t)d.gy0) * (uint64_t)dims.x) * elem;
    uint8_t* __restrict__ dst = pool + (size_t)(first + j) * slot_bytes + (size_t)z * slice_bytes;
    if ((((uintptr_t)src | (uintptr_t)dst | slice_bytes) & 15u) == 0) {
      const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src); uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst);
      const size_t nv = slice_bytes / 16;
      size_t i = threadIdx.x;
      for (; i + 3 * (size_t)blockDim.x < nv; i += 4 * (size_t)blockDim.x) {          // four loads in flight per thread
        const uint4 a = s4[i], b = s4[i + blockDim.x], c = s4[i + 2 * (size_t)blockDim.x], e = s4[i + 3 * (size_t)blockDim.x];
        d4[i] = a; d4[i + blockDim.x] = b; d4[i + 2 * (size_t)blockDim.x] = c; d4[i + 3 * (size_t)blockDim.x] = e;
      }
      for (; i < nv; i += blockDim.x) d4[i] = s4[i];
    } else {
      for (size_t i = threadIdx.x; i < slice_bytes; i += blockDim.x) dst[i] = src[i];
    }
  }
}

void SlabSampler::pull_slabs(uint32_t first, uint32_t count, cudaStream_t s) {
  if (!count) return;
  static unsigned blocks = 0;
  if (!blocks) { blocks = kPullBlocks; if (const char* e = getenv("VNR_OOC_PULL_SMS")) { const int k = atoi(e); if (k >= 1 && k <= 1024) blocks = (unsigned)k; } }
  slab_pull_kernel<<<blocks, 1024, 0, s>>>(map_dev, file_offset, elem, make_int3(dims[0], dims[1], dims[2]), table.p, first, count, pool.p, slot_bytes);
  VNR_CUDA(cudaGetLastError());
}

template <typename T> __device__ __forceinline__ float slab_load(const uint8_t* __restrict__ p, size_t i) { return (float)reinterpret_cast<const T*>(p)[i]; }
__device__ __forceinline__ float slab_value(const uint8_t* __restrict__ p, size_t i, int type) {
  switch (type) {      // read_typed_pointer
    case SLAB_U8: return slab_load<uint8_t>(p, i);
    case SLAB_I8: return slab_load<int8_t>(p, i);
    case SLAB_U16: return slab_load<uint16_t>(p, i);
    case SLAB_I16: return slab_load<int16_t>(p, i);
    case SLAB_U32: return slab_load<uint32_t>(p, i);
    case SLAB_I32: return slab_load<int32_t>(p, i);
    case SLAB_F64: return slab_load<double>(p, i);
    default: return slab_load<float>(p, i);
  }
}

// One thread per training sample (the body of the tbb::parallel_for, :1087-1113).  Uniforms: stream positions
// 5 s .. 5 s + 4 of the sampler's pcg32 = cell jitter x, y, z, slab selector, voxel selector.
__global__ void slab_sample_kernel(uint32_t n, Pcg32 base, const uint8_t* __restrict__ pool, size_t slot_bytes, const SlabDesc* __restrict__ table,
                                   uint32_t n_slots, int type, int3 dims, float vlo, float vscale,
                                   float* __restrict__ coords, float* __restrict__ values) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  Pcg32 r = base;
  r.advance(5ull * s);
  const float jx = r.next_float(), jy = r.next_float(), jz = r.next_float(), ub = r.next_float(), uv = r.next_float();
  const uint32_t bidx = min((uint32_t)(ub * (float)n_slots), n_slots - 1u);
  const SlabDesc d = table[bidx];
  const uint32_t vidx = min((uint32_t)(uv * (float)d.length), d.length - 1u);
  // to_grid_index(locate_voxel(bidx, vidx), dims)
  const uint64_t lin = d.first_voxel + vidx;
  const uint64_t sy = (uint64_t)dims.x, sz = (uint64_t)dims.x * dims.y;
  const int vx = (int)(lin % sy), vy = (int)((lin % sz) / sy), vz = (int)(lin / sz);
  const float px = jx + (float)vx, py = jy + (float)vy, pz = jz + (float)vz;
  // p * rfdims * (upper - lower) + lower with the training box lower = 0, upper = 1 (network.cu:236-238)
  coords[3 * (size_t)s] = px * (1.f / (float)dims.x) * 1.f + 0.f;
  coords[3 * (size_t)s + 1] = py * (1.f / (float)dims.y) * 1.f + 0.f;
  coords[3 * (size_t)s + 2] = pz * (1.f / (float)dims.z) * 1.f + 0.f;
  // trilinear_vkl(clamp(p, 0.5, dims - 0.5)) :302-329
  const float bx = fminf(fmaxf(px, 0.5f), (float)dims.x - 0.5f) - 0.5f;
  const float by = fminf(fmaxf(py, 0.5f), (float)dims.y - 0.5f) - 0.5f;
  const float bz = fminf(fmaxf(pz, 0.5f), (float)dims.z - 0.5f) - 0.5f;
  const float ix = truncf(bx), iy = truncf(by), iz = truncf(bz);           // std::modf: b >= 0 here
  const float wx = bx - ix, wy = by - iy, wz = bz - iz;
  const int x0 = min(max((int)ix, 0), dims.x - 1), y0 = min(max((int)iy, 0), dims.y - 1), z0 = min(max((int)iz, 0), dims.z - 1);
  const int x1 = min(x0 + 1, dims.x - 1), y1 = min(y0 + 1, dims.y - 1), z1 = min(z0 + 1, dims.z - 1);
  const uint8_t* __restrict__ slab = pool + (size_t)bidx * slot_bytes;
  auto at = [&](int x, int y, int z) {      // access_voxel :663-673 + normalise before interpolating :1105
    const size_t i = ((size_t)(z - d.gz0) * d.gny + (size_t)(y - d.gy0)) * dims.x + (size_t)x;
    const float v = (slab_value(slab, i, type) - vlo) * vscale;
    return fminf(fmaxf(v, 0.f), 1.f);
  };
  const float c000 = at(x0, y0, z0), c001 = at(x1, y0, z0), c010 = at(x0, y1, z0), c011 = at(x1, y1, z0);
  const float c100 = at(x0, y0, z1), c101 = at(x1, y0, z1), c110 = at(x0, y1, z1), c111 = at(x1, y1, z1);
  const float ox = 1.f - wx, oy = 1.f - wy, oz = 1.f - wz;
  values[s] = ox * oy * oz * c000 + wx * oy * oz * c001 + ox * wy * oz * c010 + wx * wy * oz * c011 +
              ox * oy * wz * c100 + wx * oy * wz * c101 + ox * wy * wz * c110 + wx * wy * wz * c111;
}

static uint32_t env_u32(const char* name, uint32_t fallback) {
  if (const char* e = getenv(name)) { const long v = atol(e); if (v > 0) return (uint32_t)v; }
  return fallback;
}

void outofcore_release(Volume* v) { delete v->ooc; v->ooc = nullptr; }
void outofcore_preload_kernels() { cudaFuncAttributes fa; VNR_CUDA(cudaFuncGetAttributes(&fa, slab_sample_kernel)); VNR_CUDA(cudaFuncGetAttributes(&fa, slab_pull_kernel)); }

// OutOfCoreSampler::OutOfCoreSampler (:1040-1063) + RandomBuffer::RandomBuffer (:526-582)
void outofcore_open(Volume* v, const char* path, int type, uint64_t offset, float vmin, float vmax, uint32_t n_concurrent, uint32_t n_blocks) {
  if (!(vmax > vmin)) throw InvalidError("a valid value range must be provided");          // :1068-1070
  std::unique_ptr<SlabSampler> sp(new SlabSampler);
  SlabSampler& q = *sp;
  q.elem = slab_elem_size(type); q.type = type; q.file_offset = offset; q.vmin = vmin; q.vmax = vmax;
  for (int k = 0; k < 3; ++k) q.dims[k] = v->dims[k];
  q.fd = open(path, O_RDONLY);
  if (q.fd < 0) throw InvalidError(std::string("cannot open volume file ") + path);
  const uint64_t need = offset + (uint64_t)q.dims[0] * q.dims[1] * q.dims[2] * q.elem;
  if ((uint64_t)lseek(q.fd, 0, SEEK_END) < need) throw InvalidError("volume file is shorter than dims * voxel size");
  constexpr uint64_t kStream = 32 * 1024, kAlign = 512;                                    // STREAM_SIZE, ALIGNMENT :490-491
  const uint64_t row_bytes = (uint64_t)q.dims[0] * q.elem;
  q.block_rows = (int)std::min<uint64_t>((kStream + row_bytes - 1) / row_bytes, (uint64_t)q.dims[1]);
  q.nby = (q.dims[1] + q.block_rows - 1) / q.block_rows; q.nbz = q.dims[2];
  const uint64_t gy = std::min(q.block_rows + 2, q.dims[1]), gz = std::min(3, q.dims[2]);
  q.slot_bytes = (size_t)(((uint64_t)q.dims[0] * gy * gz * q.elem + kAlign - 1) / kAlign * kAlign);
  q.n_refresh = n_concurrent ? n_concurrent : env_u32("VNR_NUM_CONCURRENT_BLOCKS", 1024);
  q.n_slots = n_blocks ? n_blocks : env_u32("VNR_NUM_BLOCKS", q.n_refresh * 64);
  if (q.n_refresh > q.n_slots) throw InvalidError("more concurrent blocks than blocks");
  if ((uint64_t)q.n_slots * q.slot_bytes > ((uint64_t)96 << 30)) throw InvalidError("slab pool larger than 96 GiB: lower VNR_NUM_BLOCKS");
  q.pool.alloc((size_t)q.n_slots * q.slot_bytes);
  q.table.alloc(q.n_slots);
  q.h_table.resize(q.n_slots);
  // pull mode (see the header): map + register the file; any failure leaves the pread path in place
  const char* pull_env = getenv("VNR_OOC_PULL");
  const uint64_t fsize = (uint64_t)lseek(q.fd, 0, SEEK_END);
  // (registering a private writable mapping may give the process its own pinned copy of the pages: files beyond 8 GiB -- or
  // VNR_OOC_PULL_MAX_GB -- stay on the pread path)
  uint64_t pull_max = (uint64_t)8 << 30;
  if (const char* e = getenv("VNR_OOC_PULL_MAX_GB")) { const long g = atol(e); if (g > 0) pull_max = (uint64_t)g << 30; }
  if ((!pull_env || atoi(pull_env) != 0) && fsize <= pull_max) {
    void* p = mmap(nullptr, (size_t)fsize, PROT_READ | PROT_WRITE, MAP_PRIVATE, q.fd, 0);
    if (p != MAP_FAILED) {
      void* dp = nullptr;
      if (cudaHostRegister(p, (size_t)fsize, cudaHostRegisterMapped) == cudaSuccess && cudaHostGetDevicePointer(&dp, p, 0) == cudaSuccess && dp) {
        q.map = p; q.map_bytes = (size_t)fsize; q.map_dev = static_cast<const uint8_t*>(dp);
        VNR_CUDA(cudaStreamCreateWithFlags(&q.pull_stream, cudaStreamNonBlocking));
        VNR_CUDA(cudaEventCreateWithFlags(&q.ev_drawn, cudaEventDisableTiming));
        VNR_CUDA(cudaEventCreateWithFlags(&q.ev_pulled, cudaEventDisableTiming));
      } else {
        cudaGetLastError();
        cudaHostUnregister(p); cudaGetLastError();
        munmap(p, (size_t)fsize);
      }
    }
  }
  for (int k = 0; k < 2; ++k) {
    if (!q.map_dev) VNR_CUDA(cudaMallocHost((void**)&q.h_stage[k], (size_t)q.n_refresh * q.slot_bytes));
    VNR_CUDA(cudaMallocHost((void**)&q.h_desc[k], (size_t)q.n_refresh * sizeof(SlabDesc)));
    VNR_CUDA(cudaEventCreateWithFlags(&q.copied[k], cudaEventDisableTiming));
  }
  // a rank of a data-parallel group selects its slabs from its own pcg32 stream (see outofcore_set_rank)
  q.rank = v->dp_world > 0 ? v->dp_rank : 0;
  q.preload(v->stream);
  outofcore_release(v);
  v->ooc = sp.release();
}

// Data-parallel training (BASELINE configs[3]): every rank keeps its OWN pool of random slabs and refreshes it from its own
// selection stream, so `world` ranks together cover `world` times as many distinct slabs per step -- the ranks' pools must
// not be copies of each other.  The reference has one process and an unseeded std::mt19937 (neural_sampler.cu:62-75); here
// rank r draws its slab choices from pcg32 stream 2 + r of seed 1337.  Called when a volume that already has an out-of-core
// sampler joins a group (vnr_volume_attach_comm / vnr_volume_dp_attach); a changed rank re-draws and re-reads the pool.
void outofcore_set_rank(Volume* v, int rank) {
  if (!v->ooc || v->ooc->rank == rank) return;
  SlabSampler& q = *v->ooc;
  VNR_CUDA(cudaStreamSynchronize(v->stream));
  q.drain();
  q.rank = rank;
  q.preload(v->stream);
}

// OutOfCoreSampler::sample (:1065-1120)
void outofcore_sample(Volume* v, float* d_xyz, float* d_target, size_t n, cudaStream_t s) {
  SlabSampler& q = *v->ooc;
  if (q.pull_in_flight) { VNR_CUDA(cudaStreamWaitEvent(s, q.ev_pulled, 0)); q.pull_in_flight = false; }     // the refresh pulled under the last step
  q.wait_and_upload(s);                                                                      // randbuf.wait_all_jobs()
  const int3 dims = make_int3(q.dims[0], q.dims[1], q.dims[2]);
  slab_sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t)n, v->sampler_rng, q.pool.p, q.slot_bytes, q.table.p, q.n_slots, q.type, dims,
                                                                  q.vmin, 1.f / (q.vmax - q.vmin), d_xyz, d_target);
  VNR_CUDA(cudaGetLastError());
  v->sampler_rng.advance(5 * (uint64_t)n);
  q.submit(-1);                                                                              // randbuf.submit_all_jobs()
  if (q.map_dev && q.pull_stream) {            // pull mode: the refresh starts now, behind this draw, on its own stream
    VNR_CUDA(cudaEventRecord(q.ev_drawn, s));
    VNR_CUDA(cudaStreamWaitEvent(q.pull_stream, q.ev_drawn, 0));
    q.wait_and_upload(q.pull_stream);
    VNR_CUDA(cudaEventRecord(q.ev_pulled, q.pull_stream));
    q.pull_in_flight = true;
  }
}

void outofcore_info(Volume* v, uint32_t* n_slots, uint32_t* n_refresh, uint64_t* slot_bytes, uint64_t* first_voxel, uint32_t* length, uint64_t* bytes_uploaded) {
  if (!v->ooc) throw StateError("no out-of-core sampler on this volume");
  SlabSampler& q = *v->ooc;
  if (n_slots) *n_slots = q.n_slots;
  if (n_refresh) *n_refresh = q.n_refresh;
  if (slot_bytes) *slot_bytes = q.slot_bytes;
  if (bytes_uploaded) *bytes_uploaded = q.bytes_uploaded;
  // the table the NEXT sample call will see: the pending refresh is applied here so the caller can restate the batch
  if (first_voxel || length) {
    q.wait_and_upload(v->stream);
    for (uint32_t i = 0; i < q.n_slots; ++i) { if (first_voxel) first_voxel[i] = q.h_table[i].first_voxel; if (length) length[i] = q.h_table[i].length; }
  }
}

}  // namespace vnr

// evaluate.cu -- quality evaluators and volume export of the neural volume:
//   vnrNeuralVolumeGetSSIM        (api.h:130; NeuralVolume::Impl::get_mssim   core/network.cu:474-549, compute_ssim :70-125)
//   vnrNeuralVolumeGetTestingLoss (api.h:131; NeuralVolume::Impl::test        core/network.cu:261-288)
//   vnrNeuralVolumeDecodeInference / DecodeReference (api.h:139-140; save_inference_volume / save_reference_volume :327-408)
// The reference evaluates the volume in 4096x16x16 blocks because a 24 GB GPU cannot hold a second copy of it;
// on B200 the decoded volume is simply materialised in HBM once (1024^3 floats = 4 GiB of 180 GB) and the
// evaluators run over it in one launch.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>
#include "train.h"
#include "vnr_device.cuh"

namespace vnr {

cudaError_t launch_decode(const DecoderDesc& d, const __half* params, const float* coords, float* out, size_t n, __half* enc_out, cudaStream_t stream);

// voxel-centre coordinates of a run of `n` voxels starting at linear index `first` of a slab whose rows are
// size.x long and planes size.x*size.y (generate_coords, network.cu:51-68; rdims = 1/dims)
__global__ void eval_coords_kernel(uint32_t n, uint64_t first, int3 size, float3 rdims, float* __restrict__ coords) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t idx = first + i, stride = (uint64_t)size.x * size.y;
  const int x = (int)(idx % size.x), y = (int)((idx % stride) / size.x), z = (int)(idx / stride);
  coords[3 * (size_t)i] = ((float)x + 0.5f) * rdims.x;
  coords[3 * (size_t)i + 1] = ((float)y + 0.5f) * rdims.y;
  coords[3 * (size_t)i + 2] = ((float)z + 0.5f) * rdims.z;
}

// decode voxel centres [first, first+count) (linear index, rows of dims.x, planes of dims.x*dims.y) into out[0..count)
static void decode_voxels(Volume* v, uint64_t first, uint64_t count, float* out, cudaStream_t s) {
  const int3 dims = make_int3(v->dims[0], v->dims[1], v->dims[2]);
  const float3 rdims = make_float3(1.f / (float)dims.x, 1.f / (float)dims.y, 1.f / (float)dims.z);
  const uint32_t chunk = (uint32_t)std::min<uint64_t>(count, 1u << 24);
  v->train_x.ensure(3 * (size_t)chunk);
  for (uint64_t o = 0; o < count; o += chunk) {
    const uint32_t n = (uint32_t)std::min<uint64_t>(chunk, count - o);
    eval_coords_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, first + o, dims, rdims, v->train_x.p);
    VNR_CUDA(cudaGetLastError());
    VNR_CUDA(launch_decode(v->cfg.desc, v->params.p, v->train_x.p, out + o, n, nullptr, s));
  }
}

// ------------------------------------------------------------------------------------------
// mean structural similarity over 7^3 uniform windows (compute_ssim, network.cu:70-125): the five window moments
// are accumulated in the reference's kz, ky, kx order in fp32 (every multiply-add explicit; the library is built
// with -fmad=false), sample covariance (NP/(NP-1)), K1 = 0.01, K2 = 0.03, data range 1.  One thread per window
// origin; per-voxel S is summed in double.
// ------------------------------------------------------------------------------------------
constexpr int kSsimWin = 7;

__device__ __forceinline__ float ssim_at(const float* __restrict__ fx_, const float* __restrict__ fy_, int3 dims, int x, int y, int z) {
  float ux = 0.f, uy = 0.f, uxx = 0.f, uyy = 0.f, uxy = 0.f;
  for (int kz = 0; kz < kSsimWin; ++kz)
    for (int ky = 0; ky < kSsimWin; ++ky) {
      const size_t row = (size_t)x + (size_t)(y + ky) * dims.x + (size_t)(z + kz) * dims.x * dims.y;
#pragma unroll
      for (int kx = 0; kx < kSsimWin; ++kx) {
        const float fx = __ldg(fx_ + row + kx), fy = __ldg(fy_ + row + kx);
        ux += fx; uy += fy;
        uxx = __fmaf_rn(fx, fx, uxx); uyy = __fmaf_rn(fy, fy, uyy); uxy = __fmaf_rn(fx, fy, uxy);
      }
    }
  const float w = 1.f / (float)(kSsimWin * kSsimWin * kSsimWin);
  ux *= w; uy *= w; uxx *= w; uyy *= w; uxy *= w;
  constexpr float NP = (float)(kSsimWin * kSsimWin * kSsimWin);
  const float cov_norm = NP / (NP - 1.f);
  const float vx = cov_norm * __fmaf_rn(-ux, ux, uxx);
  const float vy = cov_norm * __fmaf_rn(-uy, uy, uyy);
  const float vxy = cov_norm * __fmaf_rn(-ux, uy, uxy);
  const float C1 = (0.01f * 1.f) * (0.01f * 1.f), C2 = (0.03f * 1.f) * (0.03f * 1.f);
  const float A1 = __fmaf_rn(2.f * ux, uy, C1);
  const float A2 = __fmaf_rn(2.f, vxy, C2);
  const float B1 = __fmaf_rn(ux, ux, uy * uy) + C1;
  const float B2 = vx + vy + C2;
  return (A1 * A2) / (B1 * B2);
}

__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int3 dims, int3 odims, double* __restrict__ acc, float* __restrict__ s_out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int z = blockIdx.z;
  double s = 0.0;
  if (x < odims.x && y < odims.y) {
    const float S = ssim_at(gt, pred, dims, x, y, z);
    s = (double)S;
    if (s_out) s_out[(size_t)x + (size_t)y * odims.x + (size_t)z * odims.x * odims.y] = S;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(acc, t);
  }
}

double volume_ssim(Volume* v, float* h_map, cudaStream_t s) {
  if (!v->have_gt) throw StateError("[error]: missing a reference volume.");              // network.cu:476-478
  if (!v->have_params) throw StateError("the neural volume has no parameters");
  const int3 dims = make_int3(v->dims[0], v->dims[1], v->dims[2]);
  const int3 od = make_int3(dims.x - kSsimWin + 1, dims.y - kSsimWin + 1, dims.z - kSsimWin + 1);
  if (od.x <= 0 || od.y <= 0 || od.z <= 0) throw InvalidError("SSIM needs a volume of at least 7 voxels per side");
  const uint64_t total = (uint64_t)dims.x * dims.y * dims.z, n_out = (uint64_t)od.x * od.y * od.z;
  DevBuf<float> pred, smap; DevBuf<double> acc;
  pred.alloc(total); acc.alloc(1); acc.zero(s);
  if (h_map) smap.alloc(n_out);
  decode_voxels(v, 0, total, pred.p, s);
  const dim3 grid((od.x + 31) / 32, (od.y + 7) / 8, od.z);
  ssim_kernel<<<grid, 256, 0, s>>>(v->gt.p, pred.p, dims, od, acc.p, smap.p);
  VNR_CUDA(cudaGetLastError());
  double sum = 0;
  VNR_CUDA(cudaMemcpyAsync(&sum, acc.p, sizeof sum, cudaMemcpyDeviceToHost, s));
  if (h_map) VNR_CUDA(cudaMemcpyAsync(h_map, smap.p, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
  VNR_CUDA(cudaStreamSynchronize(s));
  return sum / (double)n_out;                                                               // network.cu:548
}

// ------------------------------------------------------------------------------------------
// testing loss: one fresh batch from the training sampler (it advances the shared sampler stream exactly as
// the reference's static rng does), decoded, mean |pred - target|
// ------------------------------------------------------------------------------------------
__global__ void l1_accum_kernel(uint32_t n, const float* __restrict__ pred, const float* __restrict__ targ, double* __restrict__ acc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = i < n ? (double)fabsf(pred[i] - targ[i]) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, e);
}

double volume_test_loss(Volume* v, size_t batch, cudaStream_t s) {
  if (!v->have_gt && !v->ooc) throw StateError("[error]: missing a reference volume.");  // network.cu:263-265
  if (!v->have_params) throw StateError("the neural volume has no parameters");
  if (batch == 0) batch = 1 << 16;                                                        // m_batch_size, network.cu:183
  DevBuf<float> x, y, p; DevBuf<double> acc;
  x.alloc(3 * batch); y.alloc(batch); p.alloc(batch); acc.alloc(1); acc.zero(s);
  sample_batch(v, x.p, y.p, batch, s);
  VNR_CUDA(launch_decode(v->cfg.desc, v->params.p, x.p, p.p, batch, nullptr, s));
  l1_accum_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, s>>>((uint32_t)batch, p.p, y.p, acc.p);
  VNR_CUDA(cudaGetLastError());
  double sum = 0;
  VNR_CUDA(cudaMemcpyAsync(&sum, acc.p, sizeof sum, cudaMemcpyDeviceToHost, s));
  VNR_CUDA(cudaStreamSynchronize(s));
  return sum / (double)batch;
}

// ------------------------------------------------------------------------------------------
// volume export.  File layout of the reference: dims.z records of next_multiple(dims.x*dims.y, 256) floats; the
// padding of a record holds the first voxels of the NEXT slice (generate_coords keeps counting past the slice).
// which = 0: decoded volume (save_inference_volume), 1: ground truth (save_reference_volume).  Double-buffered
// pinned staging: the download of slab k overlaps the file write of slab k-1.
// (The reference's save_reference_volume ignores its filename and always writes ./reference.bin; this one writes
// the file it is given.)
// ------------------------------------------------------------------------------------------
__global__ void eval_gt_points_kernel(uint32_t n, const float* __restrict__ gt, int3 dims, uint64_t first, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t total = (uint64_t)dims.x * dims.y * dims.z;
  uint64_t idx = first + i;
  if (idx >= total) {                              // past the last slice: the sampler clamps z to the last voxel centre
    const uint64_t stride = (uint64_t)dims.x * dims.y;
    idx = (uint64_t)(dims.z - 1) * stride + idx % stride;
  }
  out[i] = gt[idx];
}

void volume_export(Volume* v, const char* path, int which, float* range_out, cudaStream_t s) {
  if (!path || !*path) throw InvalidError("no file name");
  if (which == 1 && !v->have_gt) throw StateError("[error]: missing a reference volume.");
  if (which == 0 && !v->have_params) throw StateError("the neural volume has no parameters");
  const int3 dims = make_int3(v->dims[0], v->dims[1], v->dims[2]);
  const uint64_t slice = (uint64_t)dims.x * dims.y, rec = (slice + 255) / 256 * 256;
  const int slab_z = (int)std::max<uint64_t>(1, std::min<uint64_t>(dims.z, (64u << 20) / (rec * sizeof(float))));
  FILE* f = fopen(path, "wb");
  if (!f) throw InvalidError(std::string("Cannot open file: ") + path);
  DevBuf<float> dev[2]; float* host[2] = {nullptr, nullptr}; cudaEvent_t done[2];
  const size_t slab_floats = (size_t)rec * slab_z;
  for (int b = 0; b < 2; ++b) { dev[b].alloc(slab_floats); VNR_CUDA(cudaMallocHost((void**)&host[b], slab_floats * sizeof(float))); VNR_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming)); }
  float vmin = 3.4e38f, vmax = -3.4e38f; bool ok = true;
  auto flush = [&](int b, int nz) {
    VNR_CUDA(cudaEventSynchronize(done[b]));
    const size_t n = (size_t)rec * nz;
    for (size_t i = 0; i < n; ++i) { vmin = std::min(vmin, host[b][i]); vmax = std::max(vmax, host[b][i]); }
    ok = ok && fwrite(host[b], sizeof(float), n, f) == n;
  };
  try {
    int k = 0, prev_nz = 0;
    for (int z0 = 0; z0 < dims.z; z0 += slab_z, ++k) {
      const int b = k & 1, nz = std::min(slab_z, dims.z - z0);
      for (int z = 0; z < nz; ++z) {               // record z0+z: voxels [(z0+z)*slice, +rec)
        const uint64_t first = (uint64_t)(z0 + z) * slice;
        float* out = dev[b].p + (size_t)z * rec;
        if (which == 0) decode_voxels(v, first, rec, out, s);
        else { eval_gt_points_kernel<<<(unsigned)((rec + 255) / 256), 256, 0, s>>>((uint32_t)rec, v->gt.p, dims, first, out); VNR_CUDA(cudaGetLastError()); }
      }
      VNR_CUDA(cudaMemcpyAsync(host[b], dev[b].p, (size_t)rec * nz * sizeof(float), cudaMemcpyDeviceToHost, s));
      VNR_CUDA(cudaEventRecord(done[b], s));
      if (k > 0) flush(b ^ 1, prev_nz);
      prev_nz = nz;
    }
    if (k > 0) flush((k - 1) & 1, prev_nz);
  } catch (...) {
    for (int b = 0; b < 2; ++b) { cudaFreeHost(host[b]); cudaEventDestroy(done[b]); }
    fclose(f);
    throw;
  }
  for (int b = 0; b < 2; ++b) { cudaFreeHost(host[b]); cudaEventDestroy(done[b]); }
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw InvalidError(std::string("Error occurred at writing ") + path);
  if (range_out) { range_out[0] = vmin; range_out[1] = vmax; }
}

}  // namespace vnr

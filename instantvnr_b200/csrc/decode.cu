// decode.cu -- fused neural-volume decode: coords[N] -> values[N].
//
// Replaces, in ONE kernel, the reference's decode boundary
//   NeuralVolume::inference (core/network.cu:1043-1052) -> tcnn_inference
//   (core/networks/tcnn_impl.cu:438-448) = extract_position + kernel_grid
//   (tcnn encodings/grid.h:106-286) + kernel_mlp_fused (fully_fused_mlp.cu:495-553)
//   + trim_and_cast (common_device.h:533-542)
// with no intermediate round trip through global memory: one persistent, warp-specialised CTA per SM (640 threads): four
// producer groups gather the hash-grid features of 128-sample tiles straight into swizzled shared-memory A operands, a
// consumer group runs the MLP chain of finished tiles on tcgen05 tensor cores with TMEM accumulators (mlp_tile.cuh, two
// tiles interleaved) and writes one float per sample.  Tiles are strided over the CTAs.
#include "mlp_tile.cuh"
#include "vnr_host.h"

namespace vnr {

// Warp-specialised persistent decode: one CTA per SM, kGatherGroups producer groups of 128 threads
// (one sample row each) keep the hash-grid gather running all the time and fill a ring of A tiles in
// shared memory; one consumer group of 128 threads (warps 0-3, one TMEM lane quarter each) runs the
// MLP chain of finished tiles on the tensor cores.  full[]/empty[] mbarriers hand the tiles over; the
// gather of the next tiles overlaps the MLP of the current one.
//
// STRIDE: floats per coordinate record (3 = xyz as NeuralVolume::inference takes them,
// 4 = the marcher's (x, y, z, dt) sample records).  n_dev != nullptr: the sample count is
// read from device memory (wavefront rounds are sized on the device, no host sync); with
// round_dev the round index itself lives on the device (graph-driven wavefront): the count is
// n_dev[round] and odd rounds read coords_alt (the marcher's ping-pong sample buffers).
//
// The ring holds 7 tiles, not 8: the CTA's shared memory (7 x 16 KB + weights) then stays below 160 KB, the largest
// allocation at which the SM's L1TEX still serves scattered 16-byte loads at its full rate (measured one CTA per SM,
// tools/exp_probe_cta.py: 0.95 addresses per cycle per SM up to 160 KB, 0.89 for 164..192 KB, 0.47 above) -- and that rate is
// what bounds this kernel.  Tile j lives in stage j % n_stages (one ring shared by the producer groups, in tile order).
constexpr int kGatherGroups = 4;                 // producer groups (128 threads each)
constexpr int kMaxStages = 8;                    // A-tile ring depth (n_stages <= kMaxStages, VNR_DECODE_STAGES)
constexpr int kDecodeThreads = 128 * (1 + kGatherGroups);

template <int F, int STRIDE>
__global__ void __launch_bounds__(kDecodeThreads, 1)
decode_kernel(const DecoderDesc d, const __half* __restrict__ params, const float* __restrict__ coords, const float* __restrict__ coords_alt,
              float* __restrict__ out, uint32_t n, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ round_dev, __half* __restrict__ enc_out, uint32_t n_stages) {
  if (n_dev) {
    const uint32_t r = round_dev ? *round_dev : 0u;
    n = n_dev[r];
    if (r & 1u) coords = coords_alt;
  }
  if (n == 0) return;
  const uint32_t n_tiles = (n + kTile - 1) / kTile;
  if (blockIdx.x >= n_tiles) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;                                          // n_stages tiles of 16 KB
  uint8_t* w_smem = smem + (size_t)n_stages * MlpSmem::kATile;
  __shared__ uint64_t mbar_mma[2];
  __shared__ uint64_t mbar_full[kMaxStages], mbar_empty[kMaxStages];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  const int group = tid >> 7;                                      // 0: MLP consumers, 1..kGatherGroups: producers
  const int gtid = tid & 127;
  if (tid == 0) {
    tc05::mbar_init(&mbar_mma[0], 1); tc05::mbar_init(&mbar_mma[1], 1);
    for (int s = 0; s < kMaxStages; ++s) { tc05::mbar_init(&mbar_full[s], 128); tc05::mbar_init(&mbar_empty[s], 1); }
    tc05::fence_mbar_init();
  }
  if (tid < 32) tc05::tmem_alloc(&tmem_slot, 128);
  stage_weights(w_smem, params, d, tid, kDecodeThreads);
  tc05::fence_before_sync();
  tc05::fence_async_smem();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  // tiles of this CTA: blockIdx.x + j * gridDim.x, j = 0 .. my_tiles-1; tile j belongs to producer group j % G and goes to
  // ring stage j % n_stages, use number j / n_stages of that stage.
  const uint32_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (group == 0) {
    // ---------------- consumer: MLP on finished tiles, in tile order, two tiles interleaved ----------------
    uint32_t phase[2] = {0u, 0u};
    for (uint32_t j = 0; j < my_tiles; j += 2) {
      uint8_t* a[2] = {nullptr, nullptr};
      uint64_t* full[2] = {nullptr, nullptr};
      uint32_t parity[2] = {0u, 0u}, stage[2] = {0u, 0u};
      const int nt = j + 1 < my_tiles ? 2 : 1;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q >= nt) break;
        const uint32_t jj = j + (uint32_t)q;
        stage[q] = jj % n_stages;
        parity[q] = (jj / n_stages) & 1u;
        a[q] = a_ring + (size_t)stage[q] * MlpSmem::kATile;
        full[q] = &mbar_full[stage[q]];
      }
      float v[2];
      mlp_forward_x2(a, full, parity, w_smem, mbar_mma, phase, tmem_base, d, gtid, 1, v);
      // the last MMAs have been waited for: the tiles can be refilled
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q >= nt) break;
        if (gtid == 0) tc05::mbar_arrive(&mbar_empty[stage[q]]);
        const uint32_t s = (blockIdx.x + (j + (uint32_t)q) * gridDim.x) * kTile + (uint32_t)gtid;
        if (s < n) out[s] = v[q];
      }
    }
  } else {
    // ---------------- producers: hash-grid gather straight into the swizzled A tiles ----------------
    const uint32_t g = (uint32_t)group - 1u;
    const __half* __restrict__ grid = params + d.n_mlp;
    for (uint32_t j = g; j < my_tiles; j += kGatherGroups) {
      const uint32_t stage = j % n_stages, use = j / n_stages;
      const uint32_t s = (blockIdx.x + j * gridDim.x) * kTile + (uint32_t)gtid;
      const uint32_t sc = s < n ? s : n - 1;
      float x, y, z;
      if constexpr (STRIDE == 4) { const float4 c = __ldg(reinterpret_cast<const float4*>(coords + 4 * (size_t)sc)); x = c.x; y = c.y; z = c.z; }
      else { x = __ldg(coords + 3 * (size_t)sc); y = __ldg(coords + 3 * (size_t)sc + 1); z = __ldg(coords + 3 * (size_t)sc + 2); }
      if (use > 0) tc05::mbar_wait(&mbar_empty[stage], (use - 1u) & 1u);
      uint8_t* a_smem = a_ring + (size_t)stage * MlpSmem::kATile;
      encode_row<F>(a_smem, d, grid, x, y, z, (uint32_t)gtid);
      if (enc_out && s < n) {   // debug / test tap of the encoded features (row-major [n][enc_pad])
        for (int k = 0; k < d.enc_pad; ++k)
          enc_out[(size_t)s * d.enc_pad + k] = *reinterpret_cast<__half*>(a_smem + tc05::sw128_off(gtid, k >> 3) + (k & 7) * 2);
      }
      tc05::fence_async_smem();          // my generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc05::mbar_arrive(&mbar_full[stage]);
    }
  }

  tc05::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc05::tmem_dealloc(tmem_base, 128);
}

// Measurement tap: the hash-grid gather alone (no MLP, no shared memory, full occupancy), one
// thread per sample, features folded into one word.  Gives the achievable gather rate of the
// memory system for a coordinate distribution -- the denominator of the decode roofline.
template <int F>
__global__ void __launch_bounds__(256)
gather_probe_kernel(const DecoderDesc d, const __half* __restrict__ params, const float* __restrict__ coords, uint32_t* __restrict__ out, uint32_t n) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const __half* __restrict__ grid = params + d.n_mlp;
  const float x = coords[3 * (size_t)s], y = coords[3 * (size_t)s + 1], z = coords[3 * (size_t)s + 2];
  typedef typename FeatVec<F>::type T;
  uint32_t fold = 0;
  LevelGather<F> cur, nxt;
  cur.issue(d.lv[0], grid, x, y, z);
  for (int l = 0; l < d.n_levels; ++l) {
    if (l + 1 < d.n_levels) nxt.issue(d.lv[l + 1], grid, x, y, z);
    const T r = cur.finish();
    if constexpr (F == 8) fold ^= r.x ^ r.y ^ r.z ^ r.w;
    else if constexpr (F == 4) fold ^= r.x ^ r.y;
    else fold ^= (uint32_t)r;
    cur = nxt;
  }
  out[s] = fold;
}

cudaError_t launch_gather_probe(const DecoderDesc& d, const __half* params, const float* coords, uint32_t* out, size_t n, cudaStream_t stream) {
  if (!n) return cudaSuccess;
  const unsigned grid = (unsigned)((n + 255) / 256);
  switch (d.n_feat) {
    case 8: gather_probe_kernel<8><<<grid, 256, 0, stream>>>(d, params, coords, out, (uint32_t)n); break;
    case 4: gather_probe_kernel<4><<<grid, 256, 0, stream>>>(d, params, coords, out, (uint32_t)n); break;
    case 2: gather_probe_kernel<2><<<grid, 256, 0, stream>>>(d, params, coords, out, (uint32_t)n); break;
    default: gather_probe_kernel<1><<<grid, 256, 0, stream>>>(d, params, coords, out, (uint32_t)n); break;
  }
  return cudaGetLastError();
}

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

template <int F, int STRIDE>
static cudaError_t launch_decode_t(const DecoderDesc& d, const __half* params, const float* coords, const float* coords_alt, float* out, size_t n,
                                   const uint32_t* n_dev, const uint32_t* round_dev, size_t n_max, __half* enc_out, cudaStream_t stream) {
  // ring depth: the deepest (<= 7) that keeps the allocation at or below 160 KB (see kMaxStages), at least 4; allocations that
  // would land in the slow band measured at 100..104 KB are padded past it
  int n_stages = 7;
  while (n_stages > 4 && 1024 + (size_t)n_stages * MlpSmem::kATile + MlpSmem::weights_bytes(d.n_hidden) > 160 * 1024) --n_stages;
  static int forced = -1;
  if (forced < 0) { forced = 0; if (const char* e = getenv("VNR_DECODE_STAGES")) { const int k = atoi(e); if (k >= 4 && k <= kMaxStages) forced = k; } }
  if (forced) n_stages = forced;
  size_t smem = 1024 + (size_t)n_stages * MlpSmem::kATile + MlpSmem::weights_bytes(d.n_hidden);
  if (smem > 96 * 1024 && smem < 116 * 1024) smem = 116 * 1024;
  static size_t configured_dev[kMaxDevices] = {};   // function attributes are per device
  if (smem > 226 * 1024) return cudaErrorInvalidValue;
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  size_t& configured = configured_dev[dev % kMaxDevices];
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(decode_kernel<F, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  // persistent: one CTA per SM
  const size_t n_tiles = (n_max + kTile - 1) / kTile;
  const uint32_t grid = (uint32_t)std::min<size_t>(n_tiles, (size_t)num_sms());
  decode_kernel<F, STRIDE><<<grid, kDecodeThreads, smem, stream>>>(d, params, coords, coords_alt, out, (uint32_t)n, n_dev, round_dev, enc_out, (uint32_t)n_stages);
  return cudaGetLastError();
}

template <int STRIDE>
static cudaError_t launch_decode_f(const DecoderDesc& d, const __half* params, const float* coords, const float* coords_alt, float* out, size_t n,
                                   const uint32_t* n_dev, const uint32_t* round_dev, size_t n_max, __half* enc_out, cudaStream_t stream) {
  switch (d.n_feat) {
    case 8: return launch_decode_t<8, STRIDE>(d, params, coords, coords_alt, out, n, n_dev, round_dev, n_max, enc_out, stream);
    case 4: return launch_decode_t<4, STRIDE>(d, params, coords, coords_alt, out, n, n_dev, round_dev, n_max, enc_out, stream);
    case 2: return launch_decode_t<2, STRIDE>(d, params, coords, coords_alt, out, n, n_dev, round_dev, n_max, enc_out, stream);
    case 1: return launch_decode_t<1, STRIDE>(d, params, coords, coords_alt, out, n, n_dev, round_dev, n_max, enc_out, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_decode(const DecoderDesc& d, const __half* params, const float* coords, float* out, size_t n, __half* enc_out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  return launch_decode_f<3>(d, params, coords, nullptr, out, n, nullptr, nullptr, n, enc_out, stream);
}

// marcher variant: (x,y,z,dt) records.  Host-driven rounds pass the round's buffer and counter (round_dev = nullptr,
// samples_alt unused); graph-driven rounds pass both ping-pong buffers, the counter array and the device round index.
cudaError_t launch_decode_samples(const DecoderDesc& d, const __half* params, const float4* samples, const float4* samples_alt, float* out,
                                  const uint32_t* n_dev, const uint32_t* round_dev, size_t n_max, cudaStream_t stream) {
  if (n_max == 0) return cudaSuccess;
  return launch_decode_f<4>(d, params, reinterpret_cast<const float*>(samples), reinterpret_cast<const float*>(samples_alt), out, 0, n_dev, round_dev, n_max,
                            nullptr, stream);
}

}  // namespace vnr

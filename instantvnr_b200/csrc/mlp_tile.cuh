// mlp_tile.cuh -- the 64-wide fully-fused MLP on one 128-sample tile, on the sm_100a
// tensor path: activations (A, 128 x 64 fp16) and all weight matrices (B) live in
// 128B-swizzled shared memory, accumulators in TMEM (fp32), one thread issues
// tcgen05.mma, four warps (one TMEM lane quarter each) run the ReLU epilogue and write
// the next layer's A operand back to the same shared tile.
//
// Replaces the reference's kernel_mlp_fused (tcnn/src/fully_fused_mlp.cu:495-553;
// wmma m16n16k16 with fp16 accumulators).  Numerics: fp16 operands, fp32 accumulation,
// activations rounded to fp16 after each layer, ReLU, no biases, output layer padded to
// 16 rows with no activation, result = fp16(out[0]) widened to float.
#pragma once
#include "tc05.cuh"
#include "vnr_device.cuh"

namespace vnr {

// shared-memory carve-up of one MLP "engine" (all offsets from a 1024-B aligned base)
struct MlpSmem {
  static constexpr uint32_t kATile = kTile * 128;                 // 16 KB
  static constexpr uint32_t kWHidden = kWidth * 128;              // 8 KB per hidden matrix
  static constexpr uint32_t kWOut = kOutPad * 128;                // 2 KB
  __host__ __device__ static uint32_t weights_bytes(int n_hidden) { return (uint32_t)n_hidden * kWHidden + kWOut; }
};

// Copy all MLP matrices (row-major [out][in], fp16, params blob order: input, hidden...,
// output; fully_fused_mlp.cu:957-967) into swizzled K-major B tiles.  Columns >= in_w of
// the input matrix are zero-filled.  Called by all `nthreads` threads of the CTA.
__device__ __forceinline__ void stage_weights(uint8_t* w_smem, const __half* __restrict__ params, const DecoderDesc& d, int tid, int nthreads) {
  const uint4 zero = make_uint4(0, 0, 0, 0);
  // hidden-type matrices: matrix 0 is [64][enc_pad], others [64][64]
  for (int m = 0; m < d.n_hidden; ++m) {
    const int in_w = m == 0 ? d.enc_pad : kWidth;
    const __half* src = params + (m == 0 ? 0 : (size_t)kWidth * d.enc_pad + (size_t)(m - 1) * kWidth * kWidth);
    uint8_t* dst = w_smem + (size_t)m * MlpSmem::kWHidden;
    for (int i = tid; i < kWidth * 8; i += nthreads) {
      const int row = i >> 3, chunk = i & 7;
      uint4 v = zero;
      if (chunk * 8 < in_w) v = *reinterpret_cast<const uint4*>(src + (size_t)row * in_w + chunk * 8);
      *reinterpret_cast<uint4*>(dst + tc05::sw128_off(row, chunk)) = v;
    }
  }
  {
    const __half* src = params + (size_t)kWidth * d.enc_pad + (size_t)(d.n_hidden - 1) * kWidth * kWidth;
    uint8_t* dst = w_smem + (size_t)d.n_hidden * MlpSmem::kWHidden;
    for (int i = tid; i < kOutPad * 8; i += nthreads) {
      const int row = i >> 3, chunk = i & 7;
      *reinterpret_cast<uint4*>(dst + tc05::sw128_off(row, chunk)) = *reinterpret_cast<const uint4*>(src + (size_t)row * kWidth + chunk * 8);
    }
  }
}

__device__ __forceinline__ uint32_t relu_pack(uint32_t a, uint32_t b) {
  // fp32 accumulators -> fp16 (RNE) -> ReLU, packed as half2
  __half2 h = __floats2half2_rn(__uint_as_float(a), __uint_as_float(b));
  h = __hmax2(h, __float2half2_rn(0.f));
  return h2_as_u32(h);
}

// One hidden-layer epilogue of tile row `tid`: 64 fp32 accumulator columns -> fp16 -> ReLU -> the
// row of the (same) A tile, which becomes the next layer's operand.
__device__ __forceinline__ void mlp_epilogue_row(uint8_t* a_smem, uint32_t t_row, int tid) {
  using namespace tc05;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld32(t_row + (uint32_t)half * 32u, r);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 v = make_uint4(relu_pack(r[8 * c + 0], r[8 * c + 1]), relu_pack(r[8 * c + 2], r[8 * c + 3]),
                           relu_pack(r[8 * c + 4], r[8 * c + 5]), relu_pack(r[8 * c + 6], r[8 * c + 7]));
      *reinterpret_cast<uint4*>(a_smem + sw128_off((uint32_t)tid, (uint32_t)(half * 4 + c))) = v;
    }
  }
}

// Issue the MMAs of one layer (layer == n_hidden: the 16-row output layer) of one tile; one thread.
__device__ __forceinline__ void mlp_issue_layer(uint32_t a_addr, uint32_t w_addr, uint32_t tmem_acc, const DecoderDesc& d, int layer, uint64_t* mbar) {
  using namespace tc05;
  constexpr uint32_t idesc_hidden = make_idesc_f16(kTile, kWidth, 0, 0);
  constexpr uint32_t idesc_out = make_idesc_f16(kTile, kOutPad, 0, 0);
  fence_after_sync();
  const int ksteps = (layer == 0 ? d.enc_pad : kWidth) >> 4;
  const uint64_t ad = make_desc_sw128(a_addr);
  const uint64_t bd = make_desc_sw128(w_addr + (uint32_t)layer * MlpSmem::kWHidden);
  const uint32_t idesc = layer == d.n_hidden ? idesc_out : idesc_hidden;
  for (int k = 0; k < ksteps; ++k) mma_f16_ss(tmem_acc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0);
  mma_commit(mbar);
}

// The MLP chain on up to TWO A tiles at once, interleaved so that the tensor core works on one tile
// while the 128 threads run the other tile's epilogue: tile q uses TMEM columns [64q, 64q+64) and
// mbarrier mbar[q].  Called by exactly 128 threads (warps 0-3 of the CTA, `tid` in [0,128), thread
// `tid` owns row `tid` of both tiles) that synchronise on named barrier `bar_id`.  The tiles were
// written and proxy-fenced by other threads; `full[q]` / `full_parity[q]` is the mbarrier that says so.
// a[1] == nullptr: single tile.  Returns the network output of row `tid` of each tile (fp16-rounded).
__device__ __forceinline__ void mlp_forward_x2(uint8_t* const (&a)[2], uint64_t* const (&full)[2], const uint32_t (&full_parity)[2],
                                               const uint8_t* w_smem, uint64_t* mbar /*[2]*/, uint32_t (&phase)[2], uint32_t tmem_base,
                                               const DecoderDesc& d, int tid, int bar_id, float (&out)[2]) {
  using namespace tc05;
  const uint32_t w_addr = smem_u32(w_smem);
  const uint32_t warp_u = uniform_u32((uint32_t)tid >> 5);        // warp-uniform: warp 0 of the group issues the MMAs
  tmem_base = uniform_u32(tmem_base);
  const uint32_t lane_base = ((uint32_t)tid >> 5 & 3u) * 32u;    // TMEM lane quarter this warp may touch
  const int nt = a[1] ? 2 : 1;
  uint32_t a_addr[2], t_acc[2], t_row[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    a_addr[q] = a[q] ? uniform_u32(smem_u32(a[q])) : 0u;
    t_acc[q] = tmem_base + 64u * (uint32_t)q;
    t_row[q] = t_acc[q] + (lane_base << 16);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q >= nt) break;
    mbar_wait(full[q], full_parity[q]);
    if (warp_u == 0) { if (elect_one_sync()) mlp_issue_layer(a_addr[q], w_addr, t_acc[q], d, 0, &mbar[q]); __syncwarp(); }
  }
  for (int layer = 0; layer < d.n_hidden; ++layer) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (q >= nt) break;
      mbar_wait(&mbar[q], phase[q]);
      phase[q] ^= 1u;
      fence_after_sync();
      mlp_epilogue_row(a[q], t_row[q], tid);
      fence_before_sync();
      fence_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (warp_u == 0) { if (elect_one_sync()) mlp_issue_layer(a_addr[q], w_addr, t_acc[q], d, layer + 1, &mbar[q]); __syncwarp(); }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q >= nt) break;
    mbar_wait(&mbar[q], phase[q]);
    phase[q] ^= 1u;
    fence_after_sync();
    const uint32_t raw = tmem_ld1(t_row[q]);
    tmem_ld_wait();
    out[q] = __half2float(__float2half_rn(__uint_as_float(raw)));
  }
  fence_before_sync();
}

}  // namespace vnr

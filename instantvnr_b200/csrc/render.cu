// render.cu -- the sample-streaming ray marcher (vnrRenderMode 5) as a device-driven wavefront.
//
// Replaces core/renderer/method_raymarching.cu:931-958 (iterative_raymarching_loop) and its
// kernels :687-915.  Differences in structure (results are the same per ray):
//   * no host round trip per round: the live sample count of every round lives in device
//     memory (counters[r]) and the decode kernel reads it there; the host enqueues a bounded
//     number of rounds (an upper bound of samples per ray / n_iters) and empty rounds exit
//     immediately;
//   * ONE kernel per round besides the decode: it composites the values of round r and walks
//     the macrocell DDA for round r+1 (the reference walks the DDA twice, in the intersect
//     and again in the compose kernel);
//   * samples are compacted: a ray that emits k <= n_iters samples takes k slots (warp-
//     aggregated atomic), so every 128-row tensor-core tile of the decode is full, instead of
//     n_iters slots per live ray.
#include <cmath>
#include <cstring>

#include "march.cuh"
#include "comm.h"
#include "render.h"
#include "train.h"
#include "volume_tex.cuh"

namespace vnr {

struct RayBuffers {
  float4* rgba;        // colour.xyz, alpha
  float4* tn_ncb;      // DDA t_next.xyz, next_cell_begin
  int4* cell_base;     // DDA cell.xyz, w = first sample slot of the current round
  uint32_t* state;     // bit 31: alive; low bits: samples in the current round
  float* jitter;
  // single-shade heuristic (SingleShotPayload + final_highest_* + shading_color + jitter_ssh, method_raymarching.cu:156-162,
  // 243-255).  Rays keep their index for the whole frame here, so the in-flight and the final copies are the same arrays.
  float4* ssh_org;     // highest_org.xyz, highest_alpha
  float4* ssh_col;     // highest_color.xyz
  float4* ssh_rgba;    // shading_color: the camera pass' composited colour
  float* ssh_jitter;   // second float of the pixel's generator: the shadow ray's jitter
};

__device__ __forceinline__ void write_pixel(const FrameParams& fp, float4* __restrict__ accum, uint32_t pixel, float4 c) {
  // writePixelColor raytracing.h:196-207.  fp.frame is this renderer's frame buffer, rank 0's frame buffer over NVLink
  // (tile-parallel gather) or the caller-visible pinned host frame (zero-copy download: the pixel crosses PCIe the
  // moment its ray finishes, overlapped with the rest of the wavefront, instead of a 16 B/pixel copy after the frame).
  float4* __restrict__ frame = fp.frame;
  if (fp.frame_index != 1) {
    const float4 a = fp.accum_prev[pixel];
    c = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
  accum[pixel] = c;
  const float fi = (float)fp.frame_index;
  if (fp.host_nonzero) {
    uint32_t* w = fp.host_nonzero + (pixel >> 5);
    const uint32_t b = 1u << (pixel & 31u);
    if (c.x == 0.f && c.y == 0.f && c.z == 0.f && c.w == 0.f) { if (!(atomicAnd(w, ~b) & b)) return; }      // zero over zero: nothing to send
    else atomicOr(w, b);
  }
  frame[pixel] = make_float4(__fdiv_rn(c.x, fi), __fdiv_rn(c.y, fi), __fdiv_rn(c.z, fi), __fdiv_rn(c.w, fi));
}

// what a finished ray leaves behind (iterative_compose_kernel :813-833)
template <int SHADE>
__device__ __forceinline__ void finish_ray(const FrameParams& fp, const RayBuffers& rb, float4* __restrict__ accum,
                                           uint32_t i, uint32_t pixel, float4 rgba, float4 hi_org, float4 hi_col) {
  if (SHADE == 2) {                 // camera pass of the single-shade heuristic: park the result for the shadow pass
    rb.ssh_org[i] = hi_org; rb.ssh_col[i] = hi_col; rb.ssh_rgba[i] = rgba;
  } else if (SHADE == 3) {          // shadow pass: blend the single shade in
    const float tr = 1.f - rgba.w;
    const float4 sc = rb.ssh_rgba[i], hc = rb.ssh_col[i];
    write_pixel(fp, accum, pixel, make_float4(lerp1(VNR_SHADING_SCALE, sc.x, (hc.x * sc.w) * tr), lerp1(VNR_SHADING_SCALE, sc.y, (hc.y * sc.w) * tr),
                                                     lerp1(VNR_SHADING_SCALE, sc.z, (hc.z * sc.w) * tr), sc.w));
  } else {
    write_pixel(fp, accum, pixel, rgba);
  }
}

// The frame constants live in device memory (so that a captured graph can be replayed with a new
// camera) and are staged in shared memory by every CTA.
__device__ __forceinline__ void stage_frame_params(FrameParams* dst, const FrameParams* __restrict__ src) {
  static_assert(sizeof(FrameParams) % 4 == 0, "FrameParams must be word-sized");
  for (uint32_t k = threadIdx.x; k < sizeof(FrameParams) / 4; k += blockDim.x)
    reinterpret_cast<uint32_t*>(dst)[k] = __ldg(reinterpret_cast<const uint32_t*>(src) + k);
  __syncthreads();
}

// counters: [0] rays that hit the volume, [1] samples composited, [2 + r] decode entries emitted for round r.
// The round index comes from the host (round_dev == nullptr: bounded host-enqueued rounds) or from device
// memory (graph-driven loop: *round_dev is the round whose values were just decoded).  Round r reads
// samples[(r-1)&1] and writes samples[r&1].
// SHADE 0: no shading; 1: gradient shading (every sample is 4 consecutive decode entries: the position and its three
// forward-difference neighbours, :719-726); 2: single-shade heuristic, camera pass; 3: its shadow pass (rays start at the
// highest-contribution point of pass 2 and run along the light direction; alpha only).
template <bool FIRST, int SHADE, int MAXI>
__global__ void __launch_bounds__(128)
march_round_kernel(const FrameParams* __restrict__ fpp, RayBuffers rb, float4* __restrict__ samples0, float4* __restrict__ samples1,
                   const float* __restrict__ values, uint32_t* __restrict__ counters, int round_host, const uint32_t* __restrict__ round_dev,
                   float4* __restrict__ accum) {
  constexpr uint32_t EPS = SHADE == 1 ? 4u : 1u;             // decode entries per sample
  const int round = FIRST ? 0 : (round_dev ? (int)(*round_dev) + 1 : round_host);
  if (!FIRST && counters[2 + round - 1] == 0) return;      // nothing was alive in the previous round
  __shared__ FrameParams fp_s;
  stage_frame_params(&fp_s, fpp);
  const FrameParams& fp = fp_s;
  const float4* __restrict__ prev_samples = (round & 1) ? samples0 : samples1;
  float4* __restrict__ next_samples = (round & 1) ? samples1 : samples0;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  bool active = i < fp.n_rays;
  uint32_t pixel = 0;
  if (active) {
    pixel = ray_to_pixel(fp, i);
    active = pixel < (uint32_t)fp.width * (uint32_t)fp.height;
    if (FIRST && !active && SHADE != 3) rb.state[i] = 0;          // padding rays of a partial strip
  }
  uint32_t st = 0;
  if (active && !FIRST) { st = rb.state[i]; active = (st >> 31) != 0; }

  F3 org = f3(0, 0, 0), dir = f3(0, 0, 1), m_dir = dir;
  float tmin = 0.f, tmax = VNR_FLOAT_LARGE, jitter = 0.5f;
  float4 rgba = make_float4(0, 0, 0, 0);
  float4 hi_org = make_float4(0, 0, 0, 0), hi_col = make_float4(0, 0, 0, 0);
  DDAState dda; dda.tnx = dda.tny = dda.tnz = 0.f; dda.cx = dda.cy = dda.cz = 0; dda.ncb = 0.f;
  uint32_t n_comp = 0, cnt = 0, cbase = 0;
  bool hit = false, composing = false;

  if (active) {
    if (SHADE == 3) { const float4 o = rb.ssh_org[i]; org = f3(o.x, o.y, o.z); hi_org = o; dir = shadow_dir(fp); }
    else compute_ray(fp, pixel, org, dir);
    m_dir = f3(dir.x * fp.mc_rcp[0], dir.y * fp.mc_rcp[1], dir.z * fp.mc_rcp[2]);
    bool ok = intersect_box(tmin, tmax, org, dir, fp.bbox_lo, fp.bbox_hi);
    if (FIRST) {
      if (SHADE == 3) { jitter = rb.ssh_jitter[i]; ok = ok && hi_org.w > 0.f; }          // iterative_raygen_kernel_shadow :877-900
      else if (SHADE == 2) {
        float j1 = 0.5f;
        if (fp.jitter_mode == 0) jitter_lcg_tea16_pair((uint32_t)fp.frame_index, pixel, jitter, j1);
        rb.ssh_jitter[i] = j1;
      } else jitter = fp.jitter_mode == 0 ? jitter_lcg_tea16((uint32_t)fp.frame_index, pixel) : 0.5f;
      if (ok) {
        const F3 m_org = f3(org.x * fp.mc_rcp[0], org.y * fp.mc_rcp[1], org.z * fp.mc_rcp[2]);
        dda_init(dda, m_org, m_dir, tmin, fp.mc_dims);
        rb.jitter[i] = jitter;
        hit = true;
      } else {
        if (SHADE == 3) write_pixel(fp, accum, pixel, rb.ssh_rgba[i]);
        else finish_ray<SHADE>(fp, rb, accum, i, pixel, rgba, hi_org, hi_col);     // SHADE 2: zeros for the shadow pass (the reference memsets)
        rb.state[i] = 0;
        active = false;
      }
    } else {
      jitter = rb.jitter[i];
      rgba = rb.rgba[i];
      if (SHADE == 2) { hi_org = rb.ssh_org[i]; hi_col = rb.ssh_col[i]; }
      const float4 t = rb.tn_ncb[i];
      const int4 c = rb.cell_base[i];
      dda.tnx = t.x; dda.tny = t.y; dda.tnz = t.z; dda.ncb = t.w; dda.cx = c.x; dda.cy = c.y; dda.cz = c.z;
      cnt = st & 0xFFFFu; cbase = (uint32_t)c.w;
      composing = true;
    }
  }

  // ---- compose the samples of the previous round (iterative_compose_kernel :757-806).  Transposed layout: the slots of a
  // warp are depth-major (the j-th samples of all its rays are adjacent), so the slot of (ray, j) follows from the warp's
  // sample counts -- the same 32 rays sit in the same lanes in every round.
  if (!FIRST) {
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    uint32_t off = 0;
    bool open = composing;
    for (uint32_t j = 0; j < maxcnt; ++j) {
      uint32_t e;
      if (fp.transpose) {
        const uint32_t mask = __ballot_sync(0xffffffffu, cnt > j);
        e = cbase + EPS * (off + (uint32_t)__popc(mask & lt));
        off += (uint32_t)__popc(mask);
      } else e = cbase + EPS * j;
      if (open && cnt > j) {
        const float value = values[e];
        const float4 smp = prev_samples[e];
        float r, g, b, a;
        classify(fp, fp.tfn_color, fp.tfn_alpha, value, smp.w, r, g, b, a);
        if (SHADE == 1) {
          const F3 grad = f3(__fdiv_rn(values[e + 1] - value, fp.grad_step[0]), __fdiv_rn(values[e + 2] - value, fp.grad_step[1]),
                             __fdiv_rn(values[e + 3] - value, fp.grad_step[2]));
          shade_gradient(fp, dir, grad, r, g, b);
        } else if (SHADE == 2) {
          const float contrib = (1.f - rgba.w) * a;
          if (hi_org.w < contrib) { hi_org = make_float4(smp.x, smp.y, smp.z, contrib); hi_col = make_float4(r, g, b, 0.f); }
        }
        const float tr = 1.f - rgba.w;
        rgba.w = __fmaf_rn(tr, a, rgba.w);
        if (SHADE != 3) {
          rgba.x = __fmaf_rn(tr * r, a, rgba.x);
          rgba.y = __fmaf_rn(tr * g, a, rgba.y);
          rgba.z = __fmaf_rn(tr * b, a, rgba.z);
        }
        ++n_comp;
        if (!(rgba.w < VNR_NEARLY_ONE)) open = false;
      }
    }
    if (composing) {
      const bool resumable = dda_resumable(dda, m_dir, tmin, tmax, fp.mc_dims);
      if (!(rgba.w < VNR_NEARLY_ONE && resumable)) {
        finish_ray<SHADE>(fp, rb, accum, i, pixel, rgba, hi_org, hi_col);
        rb.state[i] = 0;
        active = false;
      }
    }
  }

  // ---- walk the DDA for the next round: up to n_iters samples (iterative_intersect_kernel :687-730)
  float4 local[MAXI];           // MAXI = 16 (the reference's N_ITERS) or 32 (unshaded marching only, vnr_renderer_set_n_iters)
  uint32_t k = 0;
  if (active) {
    const uint32_t n_iters = (uint32_t)fp.n_iters;
    march_exec(fp, dda, m_dir, tmin, tmax, [&](float tx, float ty) {
      const float tl = __fmaf_rn(jitter, ty, (1.f - jitter) * tx);
      const F3 c = madd(tl, dir, org);
      local[k] = make_float4(c.x, c.y, c.z, ty - tx);
      return (++k) < n_iters;
    });
  }
  if (active && k == 0) {
    // no sample left on this ray: the reference would carry it through one more (empty) round and
    // then find it not resumable; finish it now.
    finish_ray<SHADE>(fp, rb, accum, i, pixel, rgba, hi_org, hi_col);
    rb.state[i] = 0;
    active = false;
  }
  // ---- warp-aggregated slot reservation (compaction)
  uint32_t incl = k;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t wbase = 0;
  if (lane == 31 && total) wbase = atomicAdd(&counters[2 + round], EPS * total);
  wbase = __shfl_sync(0xffffffffu, wbase, 31);
  auto put = [&](uint32_t e, const float4 c) {
    next_samples[e] = c;
    if (SHADE == 1) {
      next_samples[e + 1] = make_float4(c.x + fp.grad_step[0], c.y, c.z, 0.f);
      next_samples[e + 2] = make_float4(c.x, c.y + fp.grad_step[1], c.z, 0.f);
      next_samples[e + 3] = make_float4(c.x, c.y, c.z + fp.grad_step[2], 0.f);
    }
  };
  uint32_t base = wbase;
  if (fp.transpose) {
    // depth-major within the warp: adjacent decode rows are the same step of neighbouring rays (8 x 4 pixel tile), which
    // share hash-grid cells on the coarse and middle levels; the warp's stores and next round's loads coalesce
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t off = 0;
    for (uint32_t j = 0; j < (uint32_t)MAXI; ++j) {
      const uint32_t mask = __ballot_sync(0xffffffffu, k > j);
      if (!mask) break;
      if (k > j) put(wbase + EPS * (off + (uint32_t)__popc(mask & lt)), local[j]);
      off += (uint32_t)__popc(mask);
    }
  } else {
    base = wbase + EPS * (incl - k);
    if (active) for (uint32_t j = 0; j < k; ++j) put(base + EPS * j, local[j]);
  }
  if (active) {
    rb.rgba[i] = rgba;
    if (SHADE == 2) { rb.ssh_org[i] = hi_org; rb.ssh_col[i] = hi_col; }
    rb.tn_ncb[i] = make_float4(dda.tnx, dda.tny, dda.tnz, dda.ncb);
    rb.cell_base[i] = make_int4(dda.cx, dda.cy, dda.cz, (int)base);
    rb.state[i] = 0x80000000u | k;
  }
  // ---- statistics
  const uint32_t hits = __ballot_sync(0xffffffffu, hit);
  uint32_t comp = n_comp;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) comp += __shfl_xor_sync(0xffffffffu, comp, o);
  if (lane == 0) {
    if (FIRST && hits) atomicAdd(&counters[0], __popc(hits));
    if (comp) atomicAdd(&counters[1], comp);
  }
}

// rays still alive after the last enqueued round (cannot happen when the bound holds; counted)
__global__ void finalize_kernel(const FrameParams* __restrict__ fpp, RayBuffers rb, uint32_t* __restrict__ leftover, float4* __restrict__ accum,
                                const uint32_t* __restrict__ counters, const uint32_t* __restrict__ round_dev) {
  if (round_dev && counters[2 + *round_dev] == 0u) return;      // graph loop ran until a round emitted nothing: no ray is alive
  __shared__ FrameParams fp_s;
  stage_frame_params(&fp_s, fpp);
  const FrameParams& fp = fp_s;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= fp.n_rays) return;
  if (rb.state[i] >> 31) {
    const uint32_t pixel = ray_to_pixel(fp, i);
    const float4 rgba = rb.rgba[i];
    if (fp.shade_mode == 2) finish_ray<2>(fp, rb, accum, i, pixel, rgba, rb.ssh_org[i], rb.ssh_col[i]);
    else if (fp.shade_mode == 3) finish_ray<3>(fp, rb, accum, i, pixel, rgba, rb.ssh_org[i], rb.ssh_col[i]);
    else write_pixel(fp, accum, pixel, rgba);
    rb.state[i] = 0;
    atomicAdd(leftover, 1u);
  }
}

// ---------------------------------------------------------------------------------------
// The single-kernel marcher (raymarching_kernel + raymarching_traceray, method_raymarching.cu:400-545): one thread walks
// one ray to the end against a resident volume.  The reference uses it for the "decoding" modes 4 / 7 / 10 (the
// progressively decoded network, api.cpp:429-438) and, on a SimpleVolume, for those and the in-shader modes 6 / 9 / 12
// (renderer.cpp:143-180).  Differences to the wavefront: cells are divided into equal steps (sample_size_scaler), the
// gradient flips to a backward difference at the upper volume faces (sampleGradient raytracing.h:113-127), and the
// single shade's shadow ray is marched inline with twice the step (raymarching_transmittance :365-398).
// SHADE 0 none, 1 gradient shading, 2 single-shade heuristic.
template <int SHADE>
__global__ void __launch_bounds__(128)
march_volume_kernel(const FrameParams* __restrict__ fpp, const float* __restrict__ vol, int3 dims, uint32_t* __restrict__ counters,
                    float4* __restrict__ accum) {
  __shared__ FrameParams fp_s;
  stage_frame_params(&fp_s, fpp);
  const FrameParams& fp = fp_s;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  bool active = i < fp.n_rays;
  uint32_t pixel = 0;
  if (active) { pixel = ray_to_pixel(fp, i); active = pixel < (uint32_t)fp.width * (uint32_t)fp.height; }
  uint32_t n_samples = 0;
  bool hit = false;
  if (active) {
    F3 org, dir;
    compute_ray(fp, pixel, org, dir);
    float tmin = 0.f, tmax = VNR_FLOAT_LARGE;
    float4 rgba = make_float4(0, 0, 0, 0);
    if (intersect_box(tmin, tmax, org, dir, fp.bbox_lo, fp.bbox_hi)) {
      hit = true;
      // RandomTEA rng(frame_index, pixel): get_floats() draws two floats per call; the camera ray takes the first of the
      // first call (:420), the shadow ray the first of the second call (:378)
      float jitter = 0.5f, j_shadow = 0.5f;
      if (fp.jitter_mode == 0) jitter_lcg_tea16_triple((uint32_t)fp.frame_index, pixel, jitter, j_shadow);
      F3 hi_org = f3(0, 0, 0), hi_col = f3(0, 0, 0);
      float hi_alpha = 0.f;
      DDAState dda;
      F3 m_dir = f3(dir.x * fp.mc_rcp[0], dir.y * fp.mc_rcp[1], dir.z * fp.mc_rcp[2]);
      dda_init(dda, f3(org.x * fp.mc_rcp[0], org.y * fp.mc_rcp[1], org.z * fp.mc_rcp[2]), m_dir, tmin, fp.mc_dims);
      march_exec<true>(fp, dda, m_dir, tmin, tmax, [&](float tx, float ty) {
        const float tl = __fmaf_rn(jitter, ty, (1.f - jitter) * tx);
        const F3 p = madd(tl, dir, org);
        const float value = sample_volume(vol, dims, p.x, p.y, p.z);
        float r, g, b, a;
        classify(fp, fp.tfn_color, fp.tfn_alpha, value, ty - tx, r, g, b, a);
        ++n_samples;
        if (SHADE == 1) {
          float sx = fp.grad_step[0], sy = fp.grad_step[1], sz = fp.grad_step[2];
          if (p.x + sx > 1.f - FLT_EPSILON) sx = -sx;
          if (p.y + sy > 1.f - FLT_EPSILON) sy = -sy;
          if (p.z + sz > 1.f - FLT_EPSILON) sz = -sz;
          const F3 grad = f3(__fdiv_rn(sample_volume(vol, dims, p.x + sx, p.y, p.z) - value, sx), __fdiv_rn(sample_volume(vol, dims, p.x, p.y + sy, p.z) - value, sy),
                             __fdiv_rn(sample_volume(vol, dims, p.x, p.y, p.z + sz) - value, sz));
          n_samples += 3;
          shade_gradient(fp, dir, grad, r, g, b);
        } else if (SHADE == 2) {
          const float contrib = (1.f - rgba.w) * a;
          if (hi_alpha < contrib) { hi_org = p; hi_col = f3(r, g, b); hi_alpha = contrib; }
        }
        const float tr = 1.f - rgba.w;
        rgba.x = __fmaf_rn(tr * r, a, rgba.x);
        rgba.y = __fmaf_rn(tr * g, a, rgba.y);
        rgba.z = __fmaf_rn(tr * b, a, rgba.z);
        rgba.w = __fmaf_rn(tr, a, rgba.w);
        return rgba.w < VNR_NEARLY_ONE;
      });
      if (SHADE == 2 && hi_alpha > 0.f) {
        // raymarching_transmittance from the highest-contribution point towards the light
        const F3 ldir = shadow_dir(fp);
        float t0 = 0.f, t1 = VNR_FLOAT_LARGE, alpha = 0.f;
        if (intersect_box(t0, t1, hi_org, ldir, fp.bbox_lo, fp.bbox_hi)) {
          const F3 lm_dir = f3(ldir.x * fp.mc_rcp[0], ldir.y * fp.mc_rcp[1], ldir.z * fp.mc_rcp[2]);
          dda_init(dda, f3(hi_org.x * fp.mc_rcp[0], hi_org.y * fp.mc_rcp[1], hi_org.z * fp.mc_rcp[2]), lm_dir, t0, fp.mc_dims);
          march_exec<true>(fp, dda, lm_dir, t0, t1, [&](float tx, float ty) {
            const float tl = __fmaf_rn(j_shadow, ty, (1.f - j_shadow) * tx);
            const F3 p = madd(tl, ldir, hi_org);
            float r, g, b, a;
            classify(fp, fp.tfn_color, fp.tfn_alpha, sample_volume(vol, dims, p.x, p.y, p.z), ty - tx, r, g, b, a);
            ++n_samples;
            alpha = __fmaf_rn(1.f - alpha, a, alpha);
            return alpha < VNR_NEARLY_ONE;
          }, 2.f /* raymarching_shadow_sampling_scale, instantvnr_types.h:137 */);
        }
        const float tr = 1.f - alpha;
        rgba.x = lerp1(VNR_SHADING_SCALE, rgba.x, (hi_col.x * rgba.w) * tr);
        rgba.y = lerp1(VNR_SHADING_SCALE, rgba.y, (hi_col.y * rgba.w) * tr);
        rgba.z = lerp1(VNR_SHADING_SCALE, rgba.z, (hi_col.z * rgba.w) * tr);
      }
    }
    write_pixel(fp, accum, pixel, rgba);
  }
  const uint32_t hits = __ballot_sync(0xffffffffu, hit);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_samples += __shfl_xor_sync(0xffffffffu, n_samples, o);
  if (lane == 0) {
    if (hits) atomicAdd(&counters[0], __popc(hits));
    if (n_samples) { atomicAdd(&counters[1], n_samples); atomicAdd(&counters[2], n_samples); }
  }
}

}  // namespace vnr
#include "pathtrace.cuh"
namespace vnr {

// Loop control of the graph-driven wavefront: one thread advances the device round index and tells
// the WHILE node whether the round that was just emitted holds any sample.
__global__ void advance_round_kernel(uint32_t* __restrict__ counters, uint32_t* __restrict__ round_dev, cudaGraphConditionalHandle handle, int init, int bound) {
  const uint32_t r = init ? 0u : *round_dev + 1u;
  *round_dev = r;
  cudaGraphSetConditional(handle, (counters[2 + r] > 0u && (int)r < bound) ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------


FrameSlot::FrameSlot() {
  VNR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  VNR_CUDA(cudaEventCreateWithFlags(&frame_done[0], cudaEventDisableTiming));
  VNR_CUDA(cudaEventCreateWithFlags(&frame_done[1], cudaEventDisableTiming));
  VNR_CUDA(cudaEventCreateWithFlags(&vol_ready, cudaEventDisableTiming));
  VNR_CUDA(cudaMallocHost((void**)&h_counters, sizeof(uint32_t) * 2 * (kMaxRounds + 4)));
  memset(h_counters, 0, sizeof(uint32_t) * 2 * (kMaxRounds + 4));
}

FrameSlot::~FrameSlot() {
  if (stream) cudaStreamSynchronize(stream);
  for (int k = 0; k < 2; ++k) { if (h_frame[k] && !h_frame_external) cudaFreeHost(h_frame[k]); if (frame_done[k]) cudaEventDestroy(frame_done[k]); }
  for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
  destroy_graph();
  if (vol_ready) cudaEventDestroy(vol_ready);
  if (h_counters) cudaFreeHost(h_counters);
  if (stream) cudaStreamDestroy(stream);
}

void FrameSlot::resize(size_t npix) {
  VNR_CUDA(cudaStreamSynchronize(stream));
  accum.alloc(npix); frame.alloc(npix);
  accum.zero(stream); frame.zero(stream);
  for (int k = 0; k < 2; ++k) {
    if (h_frame[k] && !h_frame_external) cudaFreeHost(h_frame[k]);
    h_frame[k] = nullptr;
    VNR_CUDA(cudaMallocHost((void**)&h_frame[k], npix * sizeof(float4)));
    memset(h_frame[k], 0, npix * sizeof(float4));      // pixels outside a renderer's partition are never written: they read as zero
    host_nonzero[k].alloc((npix + 31) / 32); host_nonzero[k].zero(stream); host_nonzero_valid[k] = true;     // the host frame is all zero
  }
  h_frame_external = false;
  rendered = false; downloaded = false; mapped = true;
}

void FrameSlot::destroy_graph() {
  for (int k = 0; k < 2; ++k) {
    if (loop_exec[k]) { cudaGraphExecDestroy(loop_exec[k]); loop_exec[k] = nullptr; }
    if (loop_graph[k]) { cudaGraphDestroy(loop_graph[k]); loop_graph[k] = nullptr; }
  }
  if (pt_exec) { cudaGraphExecDestroy(pt_exec); pt_exec = nullptr; }
  if (pt_graph) { cudaGraphDestroy(pt_graph); pt_graph = nullptr; }
  if (capture_stream) { cudaStreamDestroy(capture_stream); capture_stream = nullptr; }
}

Renderer::Renderer(Volume* v) : vol(v) {
  slots.emplace_back(new FrameSlot());
  if (const char* e = getenv("VNR_RM_GRAPH")) use_graph = atoi(e) != 0;    // 0: host-enqueued rounds (profilers do not see graph-body kernels)
  if (const char* e = getenv("VNR_FRAME_ZEROCOPY")) zero_copy = atoi(e) != 0;
  if (const char* e = getenv("VNR_RM_TILED")) tiled = atoi(e) != 0;          // 0: scanline ray order (A/B)
  if (const char* e = getenv("VNR_RM_TRANSPOSE")) transpose = atoi(e) != 0;  // 0: per-ray contiguous sample slots (A/B)
  if (const char* e = getenv("VNR_RM_N_ITERS")) {       // method_raymarching.cu:30-38
    int n = atoi(e);
    if (n >= 1 && n <= 32) n_iters = n;
  }
  if (const char* e = getenv("VNR_FRAMES_IN_FLIGHT")) { const int n = atoi(e); if (n >= 1 && n <= kMaxFramesInFlight) set_frames_in_flight(n); }
  vol->renderers.push_back(this);
}

Renderer::~Renderer() {
  if (rcomm) { try { comm_detach_renderer(this); } catch (...) {} }
  auto& rs = vol->renderers;
  rs.erase(std::remove(rs.begin(), rs.end(), this), rs.end());
  slots.clear();
}

void Renderer::sync_all() { for (auto& s : slots) VNR_CUDA(cudaStreamSynchronize(s->stream)); }

// Depth of the frame ring.  1 (default) is the reference's behaviour: one frame at a time.  n > 1: vnr_render returns after
// enqueueing the frame on the next slot's stream, so up to n consecutive frames overlap on the device (the latency-bound tail
// rounds of one frame run under the head of the next); vnr_map_frame returns the oldest frame that has not been mapped yet.
void Renderer::set_frames_in_flight(int n) {
  if (n < 1 || n > kMaxFramesInFlight) throw InvalidError("frames in flight must be in [1, " + std::to_string(kMaxFramesInFlight) + "]");
  if (rcomm) throw StateError("detach the renderer from its communicator before changing the frames in flight");
  sync_all();
  while ((int)slots.size() > n) slots.pop_back();
  while ((int)slots.size() < n) {
    slots.emplace_back(new FrameSlot());
    if (width > 0) slots.back()->resize((size_t)width * height);
    slots.back()->frame_target = nullptr;
  }
  n_rendered = n_mapped = 0; last_slot = 0;
  for (auto& s : slots) { s->rendered = false; s->downloaded = false; s->mapped = true; }
  reset = true;
}

// volume-side ordering against frames in flight: work that rewrites what a frame reads (parameters, macrocells, transfer
// function) waits on `s` for the last frame of every slot of every renderer of the volume
void wait_for_frames(Volume* v, cudaStream_t s) {
  for (Renderer* r : v->renderers)
    for (auto& sl : r->slots)
      if (sl->rendered && !(sl->vol_waited && sl->vol_waited_on == s)) {
        VNR_CUDA(cudaStreamWaitEvent(s, sl->frame_done[sl->map_idx], 0));
        sl->vol_waited = true; sl->vol_waited_on = s;
      }
}

uint32_t Renderer::local_rays() const {
  if (part_world <= 1) return (uint32_t)width * (uint32_t)height;
  const uint32_t nstrips = ((uint32_t)height + strip_rows - 1) / strip_rows;
  const uint32_t owned = nstrips > (uint32_t)part_rank ? (nstrips - (uint32_t)part_rank + (uint32_t)part_world - 1) / (uint32_t)part_world : 0;
  return owned * strip_rows * (uint32_t)width;
}

void Renderer::resize(int w, int h) {
  if (w <= 0 || h <= 0) throw InvalidError("framebuffer size must be positive");
  if (rcomm) throw StateError("detach the renderer from its communicator before resizing");
  width = w; height = h;
  for (auto& s : slots) s->resize((size_t)w * h);
  n_rendered = n_mapped = 0; last_slot = 0;
  reset = true;
}

static void cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static void normalize3(float* v) {
  const float d = fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0]));
  const float r = 1.0f / sqrtf(d);
  v[0] *= r; v[1] *= r; v[2] *= r;
}

void Renderer::fill_frame_params(FrameParams& fp) {
  memset(&fp, 0, sizeof fp);
  fp.width = width; fp.height = height; fp.frame_index = frame_index; fp.n_iters = n_iters;
  fp.jitter_mode = jitter_mode; fp.tex_round = 0; fp.part_rank = part_rank; fp.part_world = part_world;
  fp.strip_rows = strip_rows; fp.n_rays = local_rays();
  fp.transpose = transpose ? 1 : 0;
  fp.tiled = (tiled && width % 8 == 0 && (fp.n_rays / (uint32_t)width) % 4u == 0) ? 1 : 0;
  // camera basis (renderer.cpp:87-96)
  const float t = 2.f * tanf(fovy * 0.5f * (float)M_PI / 180.f);
  const float aspect = width / float(height);
  float dir[3] = {cam_at[0] - cam_from[0], cam_at[1] - cam_from[1], cam_at[2] - cam_from[2]};
  normalize3(dir);
  float hor[3]; cross(dir, cam_up, hor); normalize3(hor);
  const float ta = t * aspect;
  for (int k = 0; k < 3; ++k) hor[k] = ta * hor[k];
  float ver[3]; cross(hor, dir, ver);
  for (int k = 0; k < 3; ++k) ver[k] = ver[k] / aspect;
  for (int k = 0; k < 3; ++k) { fp.cam_pos[k] = cam_from[k]; fp.cam_dir[k] = dir[k]; fp.cam_hor[k] = hor[k]; fp.cam_ver[k] = ver[k]; }
  // object -> world = translate(-dims/2) * scale(dims) (network.cu:569); world -> object = inverse
  const float d[3] = {(float)vol->dims[0] * scale[0], (float)vol->dims[1] * scale[1], (float)vol->dims[2] * scale[2]};
  const float det = d[0] * d[1] * d[2];
  const float il[3] = {(d[1] * d[2]) / det, (d[0] * d[2]) / det, (d[0] * d[1]) / det};
  fp.wto_l[0] = il[0]; fp.wto_l[4] = il[1]; fp.wto_l[8] = il[2];
  for (int k = 0; k < 3; ++k) { const float tp = d[k] / -2.f; fp.wto_p[k] = -(il[k] * tp); }
  for (int k = 0; k < 3; ++k) { fp.bbox_lo[k] = clip_lo[k]; fp.bbox_hi[k] = clip_hi[k]; }
  fp.step = 1.f / sampling_rate; fp.step_rcp = sampling_rate;          // object.cpp:303-304
  for (int k = 0; k < 3; ++k) {
    fp.mc_dims[k] = vol->mc_dims[k];
    const float spacing = 16.f / (float)vol->dims[k];                  // macrocell.cu:200
    fp.mc_rcp[k] = 1.f / spacing;                                      // object.cpp:318
  }
  fp.mc_maxop = vol->mc_maxop.p;
  fp.tfn_color = vol->tfn_color.p; fp.tfn_alpha = vol->tfn_alpha.p;
  fp.n_color = vol->n_color; fp.n_alpha = vol->n_alpha;
  fp.tfn_lo = vol->tfn_lo; fp.tfn_hi = vol->tfn_hi; fp.tfn_rcp = 1.f / (vol->tfn_hi - vol->tfn_lo);
  // correct the light direction (renderer.cpp:98-101): the flip persists in the renderer's parameters
  if (fmaf(dir[2], light_dir[2], fmaf(dir[1], light_dir[1], dir[0] * light_dir[0])) > 0.f)
    for (int k = 0; k < 3; ++k) light_dir[k] = -light_dir[k];
  for (int k = 0; k < 3; ++k) {
    fp.light_dir[k] = light_dir[k];
    fp.otw_diag[k] = d[k];
    fp.grad_step[k] = 1.f / (float)vol->dims[k];                       // object.cpp:305
    fp.light_rgb[k] = 1.0f;                                            // light_directional_rgb, instantvnr_types.h:147
  }
  fp.density_scale = density_scale;                                    // object.cpp:356-359
  fp.light_ambient = 1.5f;                                             // instantvnr_types.h:146
}

int Renderer::round_bound(int iters) const {
  // samples per ray <= world-space diagonal * sampling_rate (one per step) + one clipped sample per
  // macrocell crossed + 2; a live ray consumes n_iters samples per round.
  const double dx = vol->dims[0] * scale[0], dy = vol->dims[1] * scale[1], dz = vol->dims[2] * scale[2];
  const double diag = std::sqrt(dx * dx + dy * dy + dz * dz);
  const double max_samples = std::ceil(diag * sampling_rate) + vol->mc_dims[0] + vol->mc_dims[1] + vol->mc_dims[2] + 2;
  return (int)std::ceil(max_samples / iters) + 1;
}

typedef void (*march_kernel_t)(const FrameParams*, RayBuffers, float4*, float4*, const float*, uint32_t*, int, const uint32_t*, float4*);
static march_kernel_t march_kernel(bool first, int shade, int n_iters = 16) {
  switch (shade) {
    case 1: return first ? march_round_kernel<true, 1, 16> : march_round_kernel<false, 1, 16>;
    case 2: return first ? march_round_kernel<true, 2, 16> : march_round_kernel<false, 2, 16>;
    case 3: return first ? march_round_kernel<true, 3, 16> : march_round_kernel<false, 3, 16>;
    default:
      if (n_iters > 16) return first ? march_round_kernel<true, 0, 32> : march_round_kernel<false, 0, 32>;
      return first ? march_round_kernel<true, 0, 16> : march_round_kernel<false, 0, 16>;
  }
}

// (Re)build the loop graph of one pass of one slot when anything baked into its kernel nodes changed.
void Renderer::ensure_graph(FrameSlot& S, int pass, int shade, const RayBuffers& rb, unsigned grid, size_t cap, int rounds, const float* volume_src) {
  uint32_t* cnt = S.counters.p + (size_t)pass * (kMaxRounds + 4);
  const FrameParams* fpd = reinterpret_cast<const FrameParams*>(S.fp_dev.p) + pass;
  GraphKey key;
  memset(&key, 0, sizeof key);
  key.desc = vol->cfg.desc; key.params = vol->params.p;
  key.ptrs[0] = rb.rgba; key.ptrs[1] = rb.tn_ncb; key.ptrs[2] = rb.cell_base; key.ptrs[3] = rb.state; key.ptrs[4] = rb.jitter;
  key.ptrs[5] = S.samples[0].p; key.ptrs[6] = S.samples[1].p; key.ptrs[7] = S.values.p; key.ptrs[8] = cnt; key.ptrs[9] = S.accum.p;
  key.ptrs[11] = fpd; key.ptrs[12] = rb.ssh_org; key.ptrs[13] = rb.ssh_col; key.ptrs[14] = rb.ssh_rgba; key.ptrs[15] = rb.ssh_jitter;
  key.grid = grid; key.cap = cap; key.rounds = rounds; key.volume_src = volume_src; key.shade = shade | ((shade == 0 ? n_iters : 16) << 8);
  if (S.loop_exec[pass] && !memcmp(&key, &S.graph_key[pass], sizeof key)) return;
  VNR_CUDA(cudaStreamSynchronize(S.stream));
  if (S.loop_exec[pass]) { cudaGraphExecDestroy(S.loop_exec[pass]); S.loop_exec[pass] = nullptr; }
  if (S.loop_graph[pass]) { cudaGraphDestroy(S.loop_graph[pass]); S.loop_graph[pass] = nullptr; }
  VNR_CUDA(cudaGraphCreate(&S.loop_graph[pass], 0));
  cudaGraphConditionalHandle handle;
  VNR_CUDA(cudaGraphConditionalHandleCreate(&handle, S.loop_graph[pass], 0, 0));
  uint32_t* round_dev = cnt + kMaxRounds + 2;
  // node 1: initialise the round index and the loop condition from round 0
  cudaGraphNode_t init_node;
  {
    uint32_t* c = cnt; int init = 1, bound = rounds;
    void* args[] = {&c, &round_dev, &handle, &init, &bound};
    cudaKernelNodeParams kp = {};
    kp.func = (void*)advance_round_kernel; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    VNR_CUDA(cudaGraphAddKernelNode(&init_node, S.loop_graph[pass], nullptr, 0, &kp));
  }
  // node 2: WHILE
  cudaGraphNodeParams wp = {};
  wp.type = cudaGraphNodeTypeConditional;
  wp.conditional.handle = handle; wp.conditional.type = cudaGraphCondTypeWhile; wp.conditional.size = 1;
  cudaGraphNode_t while_node;
  VNR_CUDA(cudaGraphAddNode(&while_node, S.loop_graph[pass], &init_node, 1, &wp));
  cudaGraph_t body = wp.conditional.phGraph_out[0];
  if (!S.capture_stream) VNR_CUDA(cudaStreamCreateWithFlags(&S.capture_stream, cudaStreamNonBlocking));
  VNR_CUDA(cudaStreamBeginCaptureToGraph(S.capture_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  cudaError_t e = volume_src ? launch_volume_samples(volume_src, vol->dims, S.samples[0].p, S.samples[1].p, S.values.p, cnt + 2, round_dev, cap, S.capture_stream)
                             : launch_decode_samples(vol->cfg.desc, vol->params.p, S.samples[0].p, S.samples[1].p, S.values.p, cnt + 2, round_dev, cap, S.capture_stream);
  // (fusing this 1-thread kernel into the compositing kernel -- last CTA to finish sets the condition -- was measured
  // slower: 0.717 vs 0.693 ms/frame at 1024^2; a kernel that calls cudaGraphSetConditional pays for it in every CTA)
  march_kernel(false, shade, shade == 0 ? n_iters : 16)<<<grid, 128, 0, S.capture_stream>>>(fpd, rb, S.samples[0].p, S.samples[1].p, S.values.p, cnt, 0, round_dev, S.accum.p);
  advance_round_kernel<<<1, 1, 0, S.capture_stream>>>(cnt, round_dev, handle, 0, rounds);
  cudaGraph_t captured = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(S.capture_stream, &captured);
  VNR_CUDA(e); VNR_CUDA(e2);
  VNR_CUDA(cudaGraphInstantiate(&S.loop_exec[pass], S.loop_graph[pass], 0));
  S.graph_key[pass] = key;
}

// The path-tracing wavefront (do_path_tracing_iterative, method_pathtracing.cu:796-812): raygen, then rounds of
// [decode the sample queue -> shade + next delta-tracking step + compaction] until no ray is alive.  graph_loop: the
// rounds are the body of a CUDA-graph WHILE node; otherwise the host reads the live count back every round (the
// reference's loop; what profilers see).
void Renderer::render_pathtracing(FrameSlot& S, const float* volume_src, unsigned grid, size_t cap, bool graph_loop) {
  cudaStream_t stream = S.stream;
  uint32_t* cnt = S.counters.p;
  const FrameParams* fpd = reinterpret_cast<const FrameParams*>(S.fp_dev.p);
  PtBuffers pb{S.pt_org.p, S.pt_dir.p, S.pt_rad.p, S.pt_thr.p, S.pt_tn.p, S.pt_cell.p, S.pt_list[0].p, S.pt_list[1].p};
  const unsigned shade_grid = std::min<unsigned>(grid, (unsigned)num_sms() * 16u);
  auto decode = [&](cudaStream_t s) {
    return volume_src ? launch_volume_samples(volume_src, vol->dims, S.samples[0].p, S.samples[1].p, S.values.p, cnt + kPtLive, cnt + kPtParity, cap, s)
                      : launch_decode_samples(vol->cfg.desc, vol->params.p, S.samples[0].p, S.samples[1].p, S.values.p, cnt + kPtLive, cnt + kPtParity, cap, s);
  };
  pt_raygen_kernel<<<grid, 128, 0, stream>>>(fpd, pb, S.samples[0].p, cnt, S.accum.p);
  VNR_CUDA(cudaGetLastError());
  if (graph_loop) {
    GraphKey key;
    memset(&key, 0, sizeof key);
    key.desc = vol->cfg.desc; key.params = vol->params.p;
    key.ptrs[0] = pb.org; key.ptrs[1] = pb.dir; key.ptrs[2] = pb.radiance; key.ptrs[3] = pb.thr_rng; key.ptrs[4] = pb.tn; key.ptrs[5] = pb.cell;
    key.ptrs[6] = pb.list0; key.ptrs[7] = pb.list1; key.ptrs[8] = S.samples[0].p; key.ptrs[9] = S.samples[1].p; key.ptrs[10] = S.values.p;
    key.ptrs[11] = cnt; key.ptrs[12] = S.accum.p; key.ptrs[13] = fpd;
    key.grid = shade_grid; key.cap = cap; key.volume_src = volume_src;
    if (!S.pt_exec || memcmp(&key, &S.pt_key, sizeof key)) {
      VNR_CUDA(cudaStreamSynchronize(stream));
      if (S.pt_exec) { cudaGraphExecDestroy(S.pt_exec); S.pt_exec = nullptr; }
      if (S.pt_graph) { cudaGraphDestroy(S.pt_graph); S.pt_graph = nullptr; }
      VNR_CUDA(cudaGraphCreate(&S.pt_graph, 0));
      cudaGraphConditionalHandle handle;
      VNR_CUDA(cudaGraphConditionalHandleCreate(&handle, S.pt_graph, 0, 0));
      cudaGraphNode_t init_node;
      {
        int init = 1, use = 1;
        void* args[] = {&cnt, &handle, &init, &use};
        cudaKernelNodeParams kp = {};
        kp.func = (void*)pt_advance_kernel; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
        VNR_CUDA(cudaGraphAddKernelNode(&init_node, S.pt_graph, nullptr, 0, &kp));
      }
      cudaGraphNodeParams wp = {};
      wp.type = cudaGraphNodeTypeConditional;
      wp.conditional.handle = handle; wp.conditional.type = cudaGraphCondTypeWhile; wp.conditional.size = 1;
      cudaGraphNode_t while_node;
      VNR_CUDA(cudaGraphAddNode(&while_node, S.pt_graph, &init_node, 1, &wp));
      cudaGraph_t body = wp.conditional.phGraph_out[0];
      if (!S.capture_stream) VNR_CUDA(cudaStreamCreateWithFlags(&S.capture_stream, cudaStreamNonBlocking));
      VNR_CUDA(cudaStreamBeginCaptureToGraph(S.capture_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
      cudaError_t e = decode(S.capture_stream);
      pt_shade_kernel<<<shade_grid, 128, 0, S.capture_stream>>>(fpd, pb, S.samples[0].p, S.samples[1].p, S.values.p, cnt, 0, cnt + kPtParity, S.accum.p);
      pt_advance_kernel<<<1, 1, 0, S.capture_stream>>>(cnt, handle, 0, 1);
      cudaGraph_t captured = nullptr;
      cudaError_t e2 = cudaStreamEndCapture(S.capture_stream, &captured);
      VNR_CUDA(e); VNR_CUDA(e2);
      VNR_CUDA(cudaGraphInstantiate(&S.pt_exec, S.pt_graph, 0));
      S.pt_key = key;
    }
    VNR_CUDA(cudaGraphLaunch(S.pt_exec, stream));
    S.launches = 0;
  } else {
    cudaGraphConditionalHandle none = 0;
    pt_advance_kernel<<<1, 1, 0, stream>>>(cnt, none, 1, 0);
    S.launches = 2;
    for (uint32_t r = 0;; ++r) {
      uint32_t live = 0;                                               // iterative_ray_compaction (:789-794)
      VNR_CUDA(cudaMemcpyAsync(&live, cnt + kPtLive + (r & 1u), sizeof live, cudaMemcpyDeviceToHost, stream));
      VNR_CUDA(cudaStreamSynchronize(stream));
      if (!live) break;
      VNR_CUDA(decode(stream));
      pt_shade_kernel<<<shade_grid, 128, 0, stream>>>(fpd, pb, S.samples[0].p, S.samples[1].p, S.values.p, cnt, 0, cnt + kPtParity, S.accum.p);
      pt_advance_kernel<<<1, 1, 0, stream>>>(cnt, none, 0, 0);
      VNR_CUDA(cudaGetLastError());
      S.launches += 3;
    }
  }
}

// vnrRenderMode -> shading of the marcher (renderer.cpp:152-225): 4-6 none, 7-9 gradient shading, 10-12 single-shade heuristic
static int shade_of_mode(int mode) { return mode >= 10 ? 2 : (mode >= 7 ? 1 : 0); }

void Renderer::render() {
  if (width <= 0 || height <= 0) return;                               // renderer.cpp:62
  if (mode < 4 || mode > 15) throw UnsupportedError("rendering mode " + std::to_string(mode) + " is outside this library's path (OptiX modes 0-3 are not built)");
  const bool pathtracing = mode >= 13;
  // value source of the wavefront: the network (sample streaming / in-shader modes), the progressively decoded volume
  // (decoding modes 4 / 7 / 10: the reference marches neural.texture(), api.cpp:429-438) or the ground truth (SimpleVolume renderer)
  const bool decoding = mode == 4 || mode == 7 || mode == 10 || mode == 13;
  const int shade = pathtracing ? 0 : shade_of_mode(mode);
  const float* volume_src = nullptr;
  if (gt_source) {
    if (!vol->have_gt) throw StateError("no ground-truth volume set");
    volume_src = vol->gt.p;
  } else if (decoding) {
    if (!vol->decoded.p) { vol->decoded.alloc((size_t)vol->dims[0] * vol->dims[1] * vol->dims[2]); vol->decoded.zero(vol->stream); vol->decode_blob = 0; }
    volume_src = vol->decoded.p;
  } else if (!vol->have_params) throw StateError("the neural volume has no parameters");
  // the slot this frame runs in; a frame that was never mapped is overwritten (the stream orders the reuse of the buffers)
  const int k = (int)(n_rendered % slots.size());
  FrameSlot& S = slot(k);
  FrameSlot& P = last();                                               // the previous frame: accumulation source when frame_index > 1
  cudaStream_t stream = S.stream;
  apply_l2_policy(vol, stream);            // no-op unless VNR_L2_PERSIST=1 (train.cu)
  if (n_rendered - n_mapped >= slots.size()) n_mapped = n_rendered - slots.size() + 1;      // ring full: the oldest unmapped frame is dropped
  if (reset) frame_index = 0;
  frame_index++;
  reset = false;
  FrameParams fp[2]; fill_frame_params(fp[0]);
  // zero-copy download: finished pixels are stored straight into the pinned host frame map_frame() will return
  // (cudaMallocHost memory is device-addressable under UVA); the device frame buffer is then not written
  // Communicator-attached (tile-parallel): the pinned host frames of a slot are shared by all ranks, every rank stores the
  // pixels of its strips there over its own PCIe link and a peer barrier on the slots' streams closes the frame; with the
  // download disabled the pixels go to rank 0's device frame over NVLink instead.
  RendererComm* rc = (rcomm && rcomm->resolved && rcomm->comm->world > 1) ? rcomm : nullptr;
  if (rc && (rc->width != width || rc->height != height || rc->n_slots != (int)slots.size()))
    throw StateError("frame size / frames in flight changed after vnr_renderer_attach_comm: detach and attach again on every rank");
  const int hb = S.cur;                                                // host frame this render writes; the next one takes the other
  const bool zc = zero_copy && download && (rc || !S.frame_target);
  fp[0].frame = zc ? S.h_frame[hb] : S.frame_out();
  // zero pixels that are zero in this host buffer already are not stored again (FrameParams::host_nonzero)
  static const bool skip_zero = !getenv("VNR_ZC_STORE_ALL");
  const bool track = zc && skip_zero && S.host_nonzero[hb].p;      // tile-parallel: every rank tracks the pixels of its own strips in the shared frame
  fp[0].host_nonzero = track ? S.host_nonzero[hb].p : nullptr;
  if (track && !S.host_nonzero_valid[hb]) {            // the buffer was last written by a copy: every pixel may be non-zero
    VNR_CUDA(cudaMemsetAsync(S.host_nonzero[hb].p, 0xFF, S.host_nonzero[hb].bytes(), stream));
    S.host_nonzero_valid[hb] = true;
  }
  if (download && !track) S.host_nonzero_valid[hb] = false;             // this frame reaches the host buffer some other way
  fp[0].accum_prev = frame_index > 1 ? P.accum.p : nullptr;
  const int iters = shade == 0 ? n_iters : std::min(n_iters, 16);      // shaded passes keep the reference's 16 samples per round
  fp[0].n_iters = iters;
  fp[0].shade_mode = shade; fp[1] = fp[0]; fp[1].shade_mode = 3;
  const uint32_t n_rays = fp[0].n_rays;
  const int rounds = round_bound(iters);
  const int n_pass = shade == 2 ? 2 : 1;
  // make the volume's pending work (training, tfn upload) visible to the frame stream
  VNR_CUDA(cudaEventRecord(S.vol_ready, vol->stream));
  VNR_CUDA(cudaStreamWaitEvent(stream, S.vol_ready, 0));
  // rank 0 copies the gathered device frame to the host after the frame (download without zero-copy): the peers must not
  // store pixels of this frame into that buffer before the copy of the slot's previous frame has been issued and finished
  if (rc && download && !zc) peer_barrier_sync(rc->barriers[(size_t)k], stream);
  // progressive accumulation reads the previous frame's sums: wait for that frame when it ran in another slot
  if (frame_index > 1 && &P != &S && P.rendered) VNR_CUDA(cudaStreamWaitEvent(stream, P.frame_done[P.map_idx], 0));

  // decoding modes, and a SimpleVolume in every mode but the sample-streaming ones, run the single-kernel marcher
  const bool single_kernel = decoding || (gt_source && mode != 5 && mode != 8 && mode != 11 && mode != 14);
  const size_t cap = (size_t)n_rays * iters * (shade == 1 ? 4 : 1);
  if (!single_kernel) {
    S.samples[0].ensure(cap); S.samples[1].ensure(cap); S.values.ensure(cap);
    if (pathtracing) {
      S.pt_org.ensure(n_rays); S.pt_dir.ensure(n_rays); S.pt_rad.ensure(n_rays); S.pt_thr.ensure(n_rays); S.pt_tn.ensure(n_rays); S.pt_cell.ensure(n_rays);
      S.pt_list[0].ensure(n_rays); S.pt_list[1].ensure(n_rays);
    } else { S.ray_rgba.ensure(n_rays); S.ray_tn.ensure(n_rays); S.ray_cell.ensure(n_rays); S.ray_state.ensure(n_rays); S.ray_jitter.ensure(n_rays); }
    if (shade == 2) { S.ssh_org.ensure(n_rays); S.ssh_col.ensure(n_rays); S.ssh_rgba.ensure(n_rays); S.ssh_jitter.ensure(n_rays); }
  }
  const size_t cstride = kMaxRounds + 4;
  S.counters.ensure(2 * cstride);
  if (rounds + 3 > kMaxRounds) throw UnsupportedError("sampling rate too high for the round bound");
  VNR_CUDA(cudaMemsetAsync(S.counters.p, 0, S.counters.bytes(), stream));
  RayBuffers rb{S.ray_rgba.p, S.ray_tn.p, S.ray_cell.p, S.ray_state.p, S.ray_jitter.p, S.ssh_org.p, S.ssh_col.p, S.ssh_rgba.p, S.ssh_jitter.p};
  const unsigned grid = (n_rays + 127) / 128;
  S.launches = 0;
  S.prof_used = 0;
  if (profiling) {
    while ((int)S.prof_events.size() < 2 * rounds * n_pass) { cudaEvent_t e; VNR_CUDA(cudaEventCreate(&e)); S.prof_events.push_back(e); }
  }
  S.fp_dev.ensure(2 * sizeof(FrameParams));
  // a few hundred bytes from pageable memory: staged by the driver at call time, ordered on the stream
  VNR_CUDA(cudaMemcpyAsync(S.fp_dev.p, fp, sizeof fp, cudaMemcpyHostToDevice, stream));
  const bool graph_loop = use_graph && !profiling;
  if (single_kernel && n_rays) {
    const int3 d3 = make_int3(vol->dims[0], vol->dims[1], vol->dims[2]);
    const FrameParams* fpd = reinterpret_cast<const FrameParams*>(S.fp_dev.p);
    if (pathtracing) pt_volume_kernel<<<grid, 128, 0, stream>>>(fpd, volume_src, d3, S.counters.p, S.accum.p);
    else if (shade == 1) march_volume_kernel<1><<<grid, 128, 0, stream>>>(fpd, volume_src, d3, S.counters.p, S.accum.p);
    else if (shade == 2) march_volume_kernel<2><<<grid, 128, 0, stream>>>(fpd, volume_src, d3, S.counters.p, S.accum.p);
    else march_volume_kernel<0><<<grid, 128, 0, stream>>>(fpd, volume_src, d3, S.counters.p, S.accum.p);
    VNR_CUDA(cudaGetLastError());
    S.launches = 1;
  }
  if (pathtracing && !single_kernel && n_rays) render_pathtracing(S, volume_src, grid, cap, graph_loop);
  for (int pass = 0; pass < n_pass && n_rays && !single_kernel && !pathtracing; ++pass) {
    const int sh = pass == 1 ? 3 : shade;
    uint32_t* cnt = S.counters.p + (size_t)pass * cstride;
    const FrameParams* fpd = reinterpret_cast<const FrameParams*>(S.fp_dev.p) + pass;
    march_kernel(true, sh, iters)<<<grid, 128, 0, stream>>>(fpd, rb, S.samples[0].p, S.samples[1].p, nullptr, cnt, 0, nullptr, S.accum.p);
    if (graph_loop) {
      // device-driven loop: WHILE (round has samples) { decode; compose + march; advance }
      ensure_graph(S, pass, sh, rb, grid, cap, rounds, volume_src);
      VNR_CUDA(cudaGraphLaunch(S.loop_exec[pass], stream));
    } else {
      for (int r = 0; r < rounds; ++r) {
        if (profiling) VNR_CUDA(cudaEventRecord(S.prof_events[S.prof_used++], stream));
        if (volume_src) VNR_CUDA(launch_volume_samples(volume_src, vol->dims, S.samples[r & 1].p, nullptr, S.values.p, cnt + 2 + r, nullptr, cap, stream));
        else VNR_CUDA(launch_decode_samples(vol->cfg.desc, vol->params.p, S.samples[r & 1].p, nullptr, S.values.p, cnt + 2 + r, nullptr, cap, stream));
        if (profiling) VNR_CUDA(cudaEventRecord(S.prof_events[S.prof_used++], stream));
        march_kernel(false, sh, iters)<<<grid, 128, 0, stream>>>(fpd, rb, S.samples[0].p, S.samples[1].p, S.values.p, cnt, r + 1, nullptr, S.accum.p);
      }
    }
    finalize_kernel<<<grid, 128, 0, stream>>>(fpd, rb, cnt + kMaxRounds + 3, S.accum.p, cnt, graph_loop ? cnt + kMaxRounds + 2 : nullptr);
    VNR_CUDA(cudaGetLastError());
    if (!graph_loop) S.launches += 2 + 2 * (uint64_t)rounds;     // graph path: counted from the device counters in stats()
  }
  S.last_graph = graph_loop && !single_kernel;
  S.last_pt = pathtracing;
  S.last_rounds = rounds;
  S.last_passes = single_kernel ? 1 : n_pass;
  // framebuffer.download_async (renderer.cpp:133)
  if (rc) peer_barrier_sync(rc->barriers[(size_t)k], stream);         // every rank's pixels of this frame have landed
  S.downloaded = false;
  if (download && (!rc || rc->comm->rank == 0)) {                      // the frame is mapped on rank 0
    if (!zc) VNR_CUDA(cudaMemcpyAsync(S.h_frame[hb], S.frame.p, S.frame.bytes(), cudaMemcpyDeviceToHost, stream));
    S.downloaded = true;
  }
  VNR_CUDA(cudaMemcpyAsync(S.h_counters, S.counters.p, sizeof(uint32_t) * 2 * cstride, cudaMemcpyDeviceToHost, stream));
  VNR_CUDA(cudaEventRecord(S.frame_done[hb], stream));
  S.map_idx = hb; S.cur = hb ^ 1;                                      // double-buffer swap (renderer.h:93), at render time on every rank
  S.rendered = true; S.mapped = false; S.frame_index = frame_index; S.vol_waited = false;
  last_slot = k;
  ++n_rendered;
  if (!S.downloaded) n_mapped = n_rendered;          // nothing to map: frames without a download never enter the map queue
}

// explicit framebuffer.download_async for callers that disabled the automatic one (multi-GPU rank 0
// downloads after the peers' pixels have arrived): the most recent frame
void Renderer::download_now() {
  FrameSlot& S = last();
  if (!S.rendered) throw StateError("vnr_renderer_download called before vnr_render");
  VNR_CUDA(cudaMemcpyAsync(S.h_frame[S.map_idx], S.frame.p, S.frame.bytes(), cudaMemcpyDeviceToHost, S.stream));
  S.host_nonzero_valid[S.map_idx] = false;
  VNR_CUDA(cudaEventRecord(S.frame_done[S.map_idx], S.stream));
  if (!S.downloaded) n_mapped = n_rendered - 1;      // the most recent frame becomes mappable
  S.downloaded = true; S.mapped = false; S.vol_waited = false;
}

// vnrRendererMapFrame (renderer.h:84-94): the oldest rendered frame that has not been mapped yet; with one slot that is
// the frame of the last vnr_render.  The pointer stays valid until the second-next map of the same slot.
const float* Renderer::map_frame() {
  if (n_rendered == n_mapped)
    throw StateError(!n_rendered ? "vnr_map_frame called before vnr_render"
                                 : (!last().downloaded ? "frame download is disabled on this renderer" : "vnr_map_frame: every rendered frame has already been mapped"));
  FrameSlot& S = slot((int)(n_mapped % slots.size()));
  if (!S.downloaded) throw StateError("frame download is disabled on this renderer");
  VNR_CUDA(cudaEventSynchronize(S.frame_done[S.map_idx]));
  if (rcomm && rcomm->resolved) peer_barrier_require_healthy(rcomm->barriers[(size_t)(n_mapped % slots.size())], "tile-parallel frame");
  const float* p = reinterpret_cast<const float*>(S.h_frame[S.map_idx]);
  S.mapped = true;
  ++n_mapped;
  return p;
}

void Renderer::profile(float* decode_ms, int* decode_launches) {
  FrameSlot& S = last();
  VNR_CUDA(cudaStreamSynchronize(S.stream));
  float total = 0.f; int n = 0;
  for (int k = 0; k + 1 < S.prof_used; k += 2) {
    const int pass = (k / 2) / S.last_rounds, r = (k / 2) % S.last_rounds;
    if (S.h_counters[(size_t)pass * (kMaxRounds + 4) + 2 + r] == 0) continue;           // empty round: the kernel exits immediately
    float ms = 0.f;
    VNR_CUDA(cudaEventElapsedTime(&ms, S.prof_events[k], S.prof_events[k + 1]));
    total += ms; ++n;
  }
  if (decode_ms) *decode_ms = total;
  if (decode_launches) *decode_launches = n;
}

void Renderer::stats(uint64_t* s4) {
  FrameSlot& S = last();
  VNR_CUDA(cudaStreamSynchronize(S.stream));
  const uint32_t* h_counters = S.h_counters;
  uint64_t dec = 0, rounds = 0, comp = 0, leftover = 0;
  if (S.last_pt) {                       // path tracer: every sample taken is one collision event; rounds counted on the device
    const bool wavefront = h_counters[kPtRounds] != 0;
    s4[0] = h_counters[0]; s4[1] = h_counters[1]; s4[2] = h_counters[1]; s4[3] = wavefront ? h_counters[kPtRounds] : 1;
    if (S.last_graph) S.launches = 2 + 3 * (uint64_t)h_counters[kPtRounds];
    return;
  }
  for (int pass = 0; pass < S.last_passes; ++pass) {
    const uint32_t* c = h_counters + (size_t)pass * (kMaxRounds + 4);
    for (int r = 0; r <= S.last_rounds && r < kMaxRounds; ++r) { dec += c[2 + r]; if (c[2 + r]) ++rounds; }
    comp += c[1]; leftover += c[kMaxRounds + 3];
  }
  // graph path: first round + loop init + 3 kernels per non-empty round + finalize, per pass
  if (S.last_graph) S.launches = 3 * (uint64_t)S.last_passes + 3 * rounds;
  s4[0] = h_counters[0]; s4[1] = dec; s4[2] = comp; s4[3] = rounds;
  if (leftover) throw StateError("round bound exceeded: " + std::to_string(leftover) + " rays were cut short");
}

}  // namespace vnr

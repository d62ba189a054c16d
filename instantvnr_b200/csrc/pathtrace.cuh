// pathtrace.cuh -- the volumetric path tracer (vnrRenderMode 13-15) on the same device-driven wavefront as the marcher.
// Included by render.cu after write_pixel / stage_frame_params.
//
// Follows core/renderer/method_pathtracing.cu (VARYING_MAJORANT = USE_DELTA_TRACKING_ITER = 1):
//   DeltaTrackingIter::hashit                       :545-573   delta tracking, macrocell max opacity x density scale as majorant
//   iterative_take_sample / iterative_shade          :596-672   the sample-streaming state machine (mode 14, and 15 on a network)
//   iterative_raygen_kernel / iterative_shade_kernel :674-747
//   do_path_tracing_iterative                        :796-812   host loop with a 4-byte read-back + sync per collision event
//   delta_tracking / path_tracing_traceray / path_tracing_kernel :260-296,424-470,478-516   single-kernel tracer (mode 13 on the
//                                                    decoded volume; 13 / 15 on a SimpleVolume)
//   russian_roulette :366-376, uniform_sample_sphere raytracing.h:250-269, PHASE(albedo) = 0.6 albedo :35
// Structure here: rays keep their index for the whole frame (96 B of state per ray); what is compacted every round is the
// list of live rays and, slot for slot with it, the sample queue the decode kernel reads -- so the fused decode always
// runs on full 128-row tiles.  One round = decode (or trilinear lookup) of the queue + pt_shade_kernel (accept / reject the
// collision, scatter, next delta-tracking step, append).  The round loop is a CUDA-graph WHILE node driven by the device
// counter (two ping-pong counters, so the number of rounds is unbounded); the reference syncs with the host every round.
// The reference re-uses slot indices it is still reading when it compacts (save() into params.*[atomicAdd] while other
// threads load() the same arrays, :147-170,735-746); results here are those of the race-free reading.
//
// RandomTEA is gdt::LCG<16> (instantvnr_types.h:155), the TEA-16-seeded LCG of the marcher's jitter; get_float() one draw,
// get_floats() two consecutive draws.  Checked against the reference's own path tracer compiled in place (oracle/ref_marcher).
#pragma once

namespace vnr {

struct PtBuffers {
  float4* org;         // ray origin (object space); w: bit 0 shadow ray, bits 1.. scatter index (as uint bits)
  float4* dir;         // ray direction (object space); w: DDA next_cell_begin
  float4* radiance;    // L.xyz; w: throughput.x
  float4* thr_rng;     // throughput.y, throughput.z, rng state (uint bits), unused
  float4* tn;          // DDA t_next.xyz
  int4* cell;          // DDA cell.xyz
  uint32_t* list0;     // live-ray lists (ping-pong), slot for slot with the sample queues
  uint32_t* list1;
};

// counters of a path-tracing pass: [0] rays that hit the box, [1] samples taken, [2],[3] live counts (ping-pong),
// [4] rounds run, [5] parity of the current round (what the decode kernel indexes [2..3] and its two queues with)
enum { kPtLive = 2, kPtRounds = 4, kPtParity = 5 };

struct PtRay {
  F3 org, dir, L, thr;
  float tnear, tfar;
  uint32_t rng, scatter; bool shadow;
  DDAState dda;
};

__device__ __forceinline__ F3 xfm_vec(const FrameParams& fp, F3 d) {
  const float* l = fp.wto_l;
  return f3(__fmaf_rn(d.x, l[0], __fmaf_rn(d.y, l[3], d.z * l[6])), __fmaf_rn(d.x, l[1], __fmaf_rn(d.y, l[4], d.z * l[7])),
            __fmaf_rn(d.x, l[2], __fmaf_rn(d.y, l[5], d.z * l[8])));
}

__device__ __forceinline__ F3 uniform_sample_sphere(float sx, float sy) {
  const float phi = (float)(2.0 * 3.14159265358979323846 * (double)sx);
  const float cos_theta = 1.f - 2.f * sy;
  const float sin_theta = 2.f * __fsqrt_rn(sy * (1.f - sy));
  float sp, cp;
  sincosf(phi, &sp, &cp);
  return f3(cp * sin_theta, sp * sin_theta, cos_theta);
}

__device__ __forceinline__ F3 pt_sphere_dir(const FrameParams& fp, uint32_t& rng) {
  const float sx = lcg_next(rng), sy = lcg_next(rng);
  return xfm_vec(fp, uniform_sample_sphere(sx, sy));
}

__device__ __forceinline__ void pt_new_iter(const FrameParams& fp, PtRay& r) {
  dda_init(r.dda, f3(r.org.x * fp.mc_rcp[0], r.org.y * fp.mc_rcp[1], r.org.z * fp.mc_rcp[2]),
           f3(r.dir.x * fp.mc_rcp[0], r.dir.y * fp.mc_rcp[1], r.dir.z * fp.mc_rcp[2]), r.tnear, fp.mc_dims);
}

// sampleTransferFunction without opacity correction
__device__ __forceinline__ void tfn_raw(const FrameParams& fp, float value, float& r, float& g, float& b, float& a) {
  const float v = (clampf(value, fp.tfn_lo, fp.tfn_hi) - fp.tfn_lo) * fp.tfn_rcp;
  r = g = b = a = 0.f;
  int i0, i1; float w;
  if (fp.n_color > 0) {
    tfn_coeff(v, fp.n_color, fp.tex_round, i0, i1, w);
    const float4 c0 = fp.tfn_color[i0], c1 = fp.tfn_color[i1];
    r = lerp_tex(w, c0.x, c1.x); g = lerp_tex(w, c0.y, c1.y); b = lerp_tex(w, c0.z, c1.z);
  }
  if (fp.n_alpha > 0) {
    tfn_coeff(v, fp.n_alpha, fp.tex_round, i0, i1, w);
    a = lerp_tex(w, fp.tfn_alpha[i0], fp.tfn_alpha[i1]);
  }
}

// DeltaTrackingIter::hashit: `while (DDAIter::next(lambda)) {}` with the lambda written in place
__device__ __forceinline__ bool pt_hashit(const FrameParams& fp, PtRay& r, float& rayt, float& majorant) {
  DDAState& s = r.dda;
  const F3 m_dir = f3(r.dir.x * fp.mc_rcp[0], r.dir.y * fp.mc_rcp[1], r.dir.z * fp.mc_rcp[2]);
  const int stopx = m_dir.x > 0.f ? fp.mc_dims[0] : -1, stopy = m_dir.y > 0.f ? fp.mc_dims[1] : -1, stopz = m_dir.z > 0.f ? fp.mc_dims[2] : -1;
  const float tsx = fabsf(__frcp_rn(m_dir.x)), tsy = fabsf(__frcp_rn(m_dir.y)), tsz = fabsf(__frcp_rn(m_dir.z));
  const int dx = m_dir.x > 0.f ? 1 : -1, dy = m_dir.y > 0.f ? 1 : -1, dz = m_dir.z > 0.f ? 1 : -1;
  const float tnear = r.tnear, tfar = r.tfar;
  bool found = false;
  float tau = -logf(1.f - lcg_next(r.rng));
  float t = s.ncb + tnear;
  for (;;) {
    if (s.cx == stopx || s.cy == stopy || s.cz == stopz) return found;
    const float t_closest = fminf(s.tnx, fminf(s.tny, s.tnz));
    const float cell_t0 = fmaxf(tnear + s.ncb, tnear);
    const float cell_t1 = fminf(tnear + t_closest, tfar);
    if (cell_t0 >= cell_t1) return found;
    bool go = true;
    const uint32_t idx = (uint32_t)s.cx + (uint32_t)s.cy * (uint32_t)fp.mc_dims[0] + (uint32_t)s.cz * (uint32_t)fp.mc_dims[0] * (uint32_t)fp.mc_dims[1];
    majorant = __ldg(fp.mc_maxop + idx) * fp.density_scale;
    if (!(fabsf(majorant) <= FLT_EPSILON)) {                       // empty macrocell: move on (t is not advanced, as in the reference)
      tau = __fmaf_rn(-(cell_t1 - t), majorant, tau);
      t = cell_t1;
      if (!(tau > 0.f)) {
        t = t + __fdiv_rn(tau, majorant);
        found = true;
        s.ncb = t - tnear;
        rayt = t;
        go = false;
      }
    }
    if (go || fmaxf(tnear + s.ncb, tnear) >= cell_t1) {
      bool left = false;
      if (s.tnx == t_closest) { s.tnx += tsx; s.cx += dx; if (s.cx == stopx) left = true; }
      if (!left && s.tny == t_closest) { s.tny += tsy; s.cy += dy; if (s.cy == stopy) left = true; }
      if (!left && s.tnz == t_closest) { s.tnz += tsz; s.cz += dz; if (s.cz == stopz) left = true; }
      if (left) return found;
      s.ncb = t_closest;
    }
    if (!go) return found;
  }
}

__device__ __forceinline__ bool pt_russian_roulette(PtRay& r) {
  if (r.scatter > 4u) {
    const float q = fminf(0.95f, fmaxf(r.thr.x, fmaxf(r.thr.y, r.thr.z)));
    if (lcg_next(r.rng) > q) return true;
    r.thr = f3(__fdiv_rn(r.thr.x, q), __fdiv_rn(r.thr.y, q), __fdiv_rn(r.thr.z, q));
  }
  return false;
}

__device__ __forceinline__ void pt_add_light(PtRay& r, float lr, float lg, float lb) {
  r.L = f3(__fmaf_rn(r.thr.x, lr, r.L.x), __fmaf_rn(r.thr.y, lg, r.L.y), __fmaf_rn(r.thr.z, lb, r.L.z));
}

// iterative_take_sample: next tentative collision, or the light the ray picks up when it leaves the volume
__device__ __forceinline__ bool pt_take_sample(const FrameParams& fp, PtRay& r, F3& coord, float& majorant) {
  float t;
  if (pt_hashit(fp, r, t, majorant)) { coord = madd(t, r.dir, r.org); return true; }
  if (r.scatter > 0u) {
    if (r.shadow) {
      pt_add_light(r, fp.light_rgb[0], fp.light_rgb[1], fp.light_rgb[2]);
      r.shadow = false;
      r.dir = pt_sphere_dir(fp, r.rng);
      if (!intersect_box(r.tnear, r.tfar, r.org, r.dir, fp.bbox_lo, fp.bbox_hi)) return false;      // range carried in, as the reference does
      pt_new_iter(fp, r);
      if (pt_hashit(fp, r, t, majorant)) { coord = madd(t, r.dir, r.org); return true; }
    } else pt_add_light(r, fp.light_ambient, fp.light_ambient, fp.light_ambient);
  }
  return false;
}

// iterative_shade: accept / reject the collision at `coord` with the decoded `value`
__device__ __forceinline__ bool pt_shade(const FrameParams& fp, PtRay& r, F3 coord, float value, float majorant) {
  float cr, cg, cb, a;
  tfn_raw(fp, value, cr, cg, cb, a);
  if (lcg_next(r.rng) * majorant >= a * fp.density_scale) return true;                               // null collision
  if (r.shadow) {
    r.shadow = false;
    r.dir = pt_sphere_dir(fp, r.rng);
  } else {
    if (pt_russian_roulette(r)) return false;
    ++r.scatter;
    r.org = coord; r.tnear = 0.f; r.tfar = VNR_FLOAT_LARGE;
    r.thr = f3(r.thr.x * (cr * 0.6f), r.thr.y * (cg * 0.6f), r.thr.z * (cb * 0.6f));
    r.shadow = true;
    r.dir = shadow_dir(fp);
  }
  if (!intersect_box(r.tnear, r.tfar, r.org, r.dir, fp.bbox_lo, fp.bbox_hi)) return false;
  pt_new_iter(fp, r);
  return true;
}

__device__ __forceinline__ void pt_save(const PtBuffers& pb, uint32_t i, const PtRay& r) {
  pb.org[i] = make_float4(r.org.x, r.org.y, r.org.z, __uint_as_float((r.scatter << 1) | (r.shadow ? 1u : 0u)));
  pb.dir[i] = make_float4(r.dir.x, r.dir.y, r.dir.z, r.dda.ncb);
  pb.radiance[i] = make_float4(r.L.x, r.L.y, r.L.z, r.thr.x);
  pb.thr_rng[i] = make_float4(r.thr.y, r.thr.z, __uint_as_float(r.rng), 0.f);
  pb.tn[i] = make_float4(r.dda.tnx, r.dda.tny, r.dda.tnz, 0.f);
  pb.cell[i] = make_int4(r.dda.cx, r.dda.cy, r.dda.cz, 0);
}

__device__ __forceinline__ void pt_load(const FrameParams& fp, const PtBuffers& pb, uint32_t i, PtRay& r) {
  const float4 o = pb.org[i], d = pb.dir[i], l = pb.radiance[i], t = pb.thr_rng[i], tn = pb.tn[i];
  const int4 c = pb.cell[i];
  const uint32_t bits = __float_as_uint(o.w);
  r.org = f3(o.x, o.y, o.z); r.dir = f3(d.x, d.y, d.z); r.L = f3(l.x, l.y, l.z); r.thr = f3(l.w, t.x, t.y);
  r.rng = __float_as_uint(t.z); r.scatter = bits >> 1; r.shadow = (bits & 1u) != 0u;
  r.dda.tnx = tn.x; r.dda.tny = tn.y; r.dda.tnz = tn.z; r.dda.ncb = d.w; r.dda.cx = c.x; r.dda.cy = c.y; r.dda.cz = c.z;
  r.tnear = 0.f; r.tfar = VNR_FLOAT_LARGE;                                                           // load() :127-130
  intersect_box(r.tnear, r.tfar, r.org, r.dir, fp.bbox_lo, fp.bbox_hi);
}

// append live rays of a warp to the next round's list / sample queue (warp-aggregated reservation)
__device__ __forceinline__ void pt_append(bool alive, uint32_t ray, F3 coord, float majorant, uint32_t* __restrict__ live_count,
                                          uint32_t* __restrict__ list, float4* __restrict__ queue) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t mask = __ballot_sync(0xffffffffu, alive);
  if (!mask) return;
  uint32_t base = 0;
  const int leader = __ffs(mask) - 1;
  if ((int)lane == leader) base = atomicAdd(live_count, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (alive) {
    const uint32_t slot = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
    list[slot] = ray;
    queue[slot] = make_float4(coord.x, coord.y, coord.z, majorant);
  }
}

__global__ void __launch_bounds__(128)
pt_raygen_kernel(const FrameParams* __restrict__ fpp, PtBuffers pb, float4* __restrict__ queue0, uint32_t* __restrict__ counters, float4* __restrict__ accum) {
  __shared__ FrameParams fp_s;
  stage_frame_params(&fp_s, fpp);
  const FrameParams& fp = fp_s;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = i < fp.n_rays;
  uint32_t pixel = 0;
  if (active) { pixel = ray_to_pixel(fp, i); active = pixel < (uint32_t)fp.width * (uint32_t)fp.height; }
  bool alive = false, hit = false;
  PtRay r; F3 coord = f3(0, 0, 0); float majorant = 0.f;
  if (active) {
    compute_ray(fp, pixel, r.org, r.dir);
    r.L = f3(0, 0, 0); r.thr = f3(1, 1, 1); r.scatter = 0; r.shadow = false;
    r.rng = tea16((uint32_t)fp.frame_index, pixel);
    r.tnear = 0.f; r.tfar = VNR_FLOAT_LARGE;
    r.dda.tnx = r.dda.tny = r.dda.tnz = 0.f; r.dda.cx = r.dda.cy = r.dda.cz = 0; r.dda.ncb = 0.f;
    if (intersect_box(r.tnear, r.tfar, r.org, r.dir, fp.bbox_lo, fp.bbox_hi)) {
      hit = true;
      pt_new_iter(fp, r);
      alive = pt_take_sample(fp, r, coord, majorant);
    }
    if (alive) pt_save(pb, i, r);
    else write_pixel(fp, accum, pixel, make_float4(r.L.x, r.L.y, r.L.z, 1.f));
  }
  pt_append(alive, i, coord, majorant, counters + kPtLive, pb.list0, queue0);
  const uint32_t hits = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31u) == 0 && hits) atomicAdd(&counters[0], __popc(hits));
}

// Round r >= 1: the values of queue (r-1)&1 are in; shade, take the next sample, append to queue r&1.
// parity_dev == nullptr: host-enqueued round with parity `round_host & 1`.
__global__ void __launch_bounds__(128)
pt_shade_kernel(const FrameParams* __restrict__ fpp, PtBuffers pb, float4* __restrict__ queue0, float4* __restrict__ queue1, const float* __restrict__ values,
                uint32_t* __restrict__ counters, int round_host, const uint32_t* __restrict__ parity_dev, float4* __restrict__ accum) {
  __shared__ FrameParams fp_s;
  stage_frame_params(&fp_s, fpp);
  const FrameParams& fp = fp_s;
  const uint32_t prev = parity_dev ? *parity_dev : (uint32_t)((round_host - 1) & 1);                 // parity of the round that was just decoded
  const uint32_t n_prev = counters[kPtLive + prev];
  const float4* __restrict__ prev_queue = prev ? queue1 : queue0;
  float4* __restrict__ next_queue = prev ? queue0 : queue1;
  const uint32_t* __restrict__ prev_list = prev ? pb.list1 : pb.list0;
  uint32_t* __restrict__ next_list = prev ? pb.list0 : pb.list1;
  uint32_t* live_next = counters + kPtLive + (prev ^ 1u);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); k0 < n_prev; k0 += stride) {    // warp-uniform trip count
    const uint32_t k = k0 + (threadIdx.x & 31u);
    bool alive = false;
    uint32_t ray = 0; F3 coord = f3(0, 0, 0); float majorant = 0.f;
    if (k < n_prev) {
      ray = prev_list[k];
      const float4 q = prev_queue[k];
      PtRay r;
      pt_load(fp, pb, ray, r);
      alive = pt_shade(fp, r, f3(q.x, q.y, q.z), values[k], q.w) && pt_take_sample(fp, r, coord, majorant);
      if (alive) pt_save(pb, ray, r);
      else write_pixel(fp, accum, ray_to_pixel(fp, ray), make_float4(r.L.x, r.L.y, r.L.z, 1.f));
    }
    pt_append(alive, ray, coord, majorant, live_next, next_list, next_queue);
  }
}

// Loop control: after raygen (init) or after a shade round.  Publishes the parity of the round to decode next, keeps the
// totals, clears the counter the next shade round appends to, and tells the WHILE node whether anything is alive.
__global__ void pt_advance_kernel(uint32_t* __restrict__ counters, cudaGraphConditionalHandle handle, int init, int use_handle) {
  const uint32_t cur = init ? 0u : (counters[kPtParity] ^ 1u);
  counters[kPtParity] = cur;
  const uint32_t n = counters[kPtLive + cur];
  counters[kPtLive + (cur ^ 1u)] = 0u;
  if (n) { counters[1] += n; counters[kPtRounds] += 1u; }
  if (use_handle) cudaGraphSetConditional(handle, n > 0u ? 1u : 0u);
}

// The single-kernel tracer against a resident volume (path_tracing_kernel / path_tracing_traceray / delta_tracking)
__global__ void __launch_bounds__(128)
pt_volume_kernel(const FrameParams* __restrict__ fpp, const float* __restrict__ vol, int3 dims, uint32_t* __restrict__ counters, float4* __restrict__ accum) {
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
    PtRay r;
    compute_ray(fp, pixel, r.org, r.dir);
    r.L = f3(0, 0, 0); r.thr = f3(1, 1, 1); r.scatter = 0; r.shadow = false;
    r.rng = tea16((uint32_t)fp.frame_index, pixel);
    r.tnear = 0.f; r.tfar = VNR_FLOAT_LARGE;
    bool first = true;
    while (intersect_box(r.tnear, r.tfar, r.org, r.dir, fp.bbox_lo, fp.bbox_hi)) {
      if (first) { hit = true; first = false; }
      float t = r.tnear, majorant = 0.f, ar = 0.f, ag = 0.f, ab = 0.f;
      bool found = false;
      pt_new_iter(fp, r);
      while (pt_hashit(fp, r, t, majorant)) {
        const F3 c = madd(t, r.dir, r.org);
        float cr, cg, cb, a;
        tfn_raw(fp, sample_volume(vol, dims, c.x, c.y, c.z), cr, cg, cb, a);
        ++n_samples;
        if (lcg_next(r.rng) * majorant < a * fp.density_scale) { ar = cr; ag = cg; ab = cb; found = true; break; }
      }
      if (r.shadow) {
        if (!found) pt_add_light(r, fp.light_rgb[0], fp.light_rgb[1], fp.light_rgb[2]);
        r.tnear = 0.f; r.tfar = VNR_FLOAT_LARGE;
        r.dir = pt_sphere_dir(fp, r.rng);
        r.shadow = false;
      } else {
        if (!found) { if (r.scatter > 0u) pt_add_light(r, fp.light_ambient, fp.light_ambient, fp.light_ambient); break; }
        if (pt_russian_roulette(r)) break;
        ++r.scatter;
        r.org = madd(t, r.dir, r.org);
        r.thr = f3(r.thr.x * (ar * 0.6f), r.thr.y * (ag * 0.6f), r.thr.z * (ab * 0.6f));
        r.tnear = 0.f; r.tfar = VNR_FLOAT_LARGE;
        r.dir = shadow_dir(fp);
        r.shadow = true;
      }
    }
    write_pixel(fp, accum, pixel, make_float4(r.L.x, r.L.y, r.L.z, 1.f));
  }
  const uint32_t hits = __ballot_sync(0xffffffffu, hit);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_samples += __shfl_xor_sync(0xffffffffu, n_samples, o);
  if (lane == 0) {
    if (hits) atomicAdd(&counters[0], __popc(hits));
    if (n_samples) atomicAdd(&counters[1], n_samples);
  }
}

}  // namespace vnr

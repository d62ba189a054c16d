// march.cuh -- per-ray device functions of the marcher: camera ray, box test, macrocell
// DDA with adaptive step, transfer-function classification, front-to-back compositing.
//
// Follows the reference's sample-streaming marcher:
//   compute_ray            core/renderer/method_raymarching.cu:658-685
//   _intersectBox          core/renderer/raytracing.h:9-36
//   DDAIter                core/renderer/dda.h:20-138
//   RayMarchingIter::exec  core/renderer/method_raymarching.cu:555-600 (ADAPTIVE_SAMPLING=1)
//   adaptiveSamplingRate   raytracing.h:188-194,  opacityUpperBound :172-186
//   sampleTransferFunction raytracing.h:147-155 (array1dNodal :71-81), opacityCorrection :166-170
//   blending + early termination   method_raymarching.cu:797-806 (nearly_one = 0.9999)
// Unlike the reference the DDA is walked ONCE per round (the step length of every sample is
// stored with the sample instead of replaying the iterator in the compose kernel).
// Compiled with -fmad=false; fused multiply-adds are explicit (__fmaf_rn) and match the oracle.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace vnr {

#define VNR_FLOAT_LARGE 1e20f
#define VNR_NEARLY_ONE 0.9999f

struct FrameParams {
  int width, height, frame_index, n_iters;
  int jitter_mode, tex_round, part_rank, part_world;
  int tiled, transpose, pad_[2];                // rays are dealt to warps in 8 x 4 pixel tiles (width % 8 == 0, local rows % 4 == 0)
  uint32_t n_rays, strip_rows;       // local rays of this partition; rows per interleaved strip
  float cam_pos[3], cam_dir[3], cam_hor[3], cam_ver[3];
  float wto_l[9], wto_p[3];
  float bbox_lo[3], bbox_hi[3];
  float step, step_rcp;
  int mc_dims[3];
  float mc_rcp[3];
  const float* mc_maxop;
  const float4* tfn_color;
  const float* tfn_alpha;
  int n_color, n_alpha;
  float tfn_lo, tfn_hi, tfn_rcp;
  // shaded modes (method_raymarching.cu:773-833): 0 none, 1 gradient shading, 2 single-shade heuristic, 3 its shadow pass
  int shade_mode;
  float light_dir[3];                // world space, sign-corrected against the camera (renderer.cpp:98-101)
  float otw_diag[3];                 // object->world linear part (network.cu:569)
  float grad_step[3];                // object.cpp:305
  // path tracer (method_pathtracing.cu): majorant scale and lights (instantvnr_types.h:102,146-147)
  float density_scale, light_ambient, light_rgb[3];
  float4* frame;                     // where finished pixels go: local / peer frame buffer or the mapped pinned host frame
  uint32_t* host_nonzero;            // zero-copy host frame only (else nullptr): one bit per pixel, set while the pixel of THAT host
                                     // buffer holds a non-zero value.  A pixel that is zero and was zero (background around the
                                     // volume, frame after frame) is not stored again: stores from the SMs into pinned host memory
                                     // are the slow part of the download (~20-25 GB/s), a third of a frame's pixels are background
  const float4* accum_prev;          // accumulation buffer of the previous frame (read when frame_index != 1; frames in flight
                                     // accumulate into their own slot's buffer)
};

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 f3(const float* p) { return f3(p[0], p[1], p[2]); }
__device__ __forceinline__ F3 madd(float s, F3 a, F3 b) { return f3(__fmaf_rn(s, a.x, b.x), __fmaf_rn(s, a.y, b.y), __fmaf_rn(s, a.z, b.z)); }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// local ray index -> pixel index.  tiled: a warp owns an 8 x 4 pixel tile (its 32 rays stay spatially close all along their
// way through the volume, so the samples a warp emits share hash-grid cells on the coarse and middle levels); otherwise
// scanline order.  Partition: strips of `strip_rows` image rows are dealt round-robin to the ranks; rank r owns strips
// r, r+world, ...
__device__ __forceinline__ uint32_t ray_to_pixel(const FrameParams& fp, uint32_t i) {
  const uint32_t w = (uint32_t)fp.width;
  uint32_t lrow, x;
  if (fp.tiled) {
    const uint32_t tiles_x = w >> 3, tile = i >> 5, lane = i & 31u;
    const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
    lrow = ty * 4u + (lane >> 3); x = tx * 8u + (lane & 7u);
  } else { lrow = i / w; x = i - lrow * w; }
  if (fp.part_world <= 1) return lrow * w + x;
  const uint32_t strip = lrow / fp.strip_rows, in = lrow - strip * fp.strip_rows;
  const uint32_t y = (strip * (uint32_t)fp.part_world + (uint32_t)fp.part_rank) * fp.strip_rows + in;
  return y * w + x;
}

// gdt::LCG<16> (TEA-initialised LCG; OVR gdt/random/random.h, un-vendored).  The fork's get_floats() is taken to
// draw two consecutive floats per call: the streaming raygen uses .x for the camera ray and .y for the shadow ray
// (method_raymarching.cu:851-852,870), the single-kernel marcher .x of the first call and .x of the second (:378,420).
__device__ __forceinline__ uint32_t tea16(uint32_t val0, uint32_t val1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
__device__ __forceinline__ float lcg_next(uint32_t& state) {
  state = 1664525u * state + 1013904223u;
  return (float)(state & 0x00FFFFFFu) / (float)0x01000000u;
}
__device__ __forceinline__ float jitter_lcg_tea16(uint32_t val0, uint32_t val1) { uint32_t s = tea16(val0, val1); return lcg_next(s); }
__device__ __forceinline__ void jitter_lcg_tea16_pair(uint32_t val0, uint32_t val1, float& j0, float& j1) {
  uint32_t s = tea16(val0, val1); j0 = lcg_next(s); j1 = lcg_next(s);
}
__device__ __forceinline__ void jitter_lcg_tea16_triple(uint32_t val0, uint32_t val1, float& j0, float& j2) {
  uint32_t s = tea16(val0, val1); j0 = lcg_next(s); (void)lcg_next(s); j2 = lcg_next(s);
}

__device__ __forceinline__ void compute_ray(const FrameParams& fp, uint32_t pixel, F3& org, F3& dir) {
  const uint32_t ix = pixel % (uint32_t)fp.width, iy = pixel / (uint32_t)fp.width;
  const float sx = __fdiv_rn((float)ix + .5f, (float)fp.width), sy = __fdiv_rn((float)iy + .5f, (float)fp.height);
  F3 d = madd(sy - 0.5f, f3(fp.cam_ver), madd(sx - 0.5f, f3(fp.cam_hor), f3(fp.cam_dir)));
  const float r = __fdiv_rn(1.0f, __fsqrt_rn(dot3(d, d)));
  d = f3(r * d.x, r * d.y, r * d.z);
  const float* l = fp.wto_l;
  const F3 p = f3(fp.cam_pos);
  org = f3(__fmaf_rn(p.x, l[0], __fmaf_rn(p.y, l[3], __fmaf_rn(p.z, l[6], fp.wto_p[0]))),
           __fmaf_rn(p.x, l[1], __fmaf_rn(p.y, l[4], __fmaf_rn(p.z, l[7], fp.wto_p[1]))),
           __fmaf_rn(p.x, l[2], __fmaf_rn(p.y, l[5], __fmaf_rn(p.z, l[8], fp.wto_p[2]))));
  dir = f3(__fmaf_rn(d.x, l[0], __fmaf_rn(d.y, l[3], d.z * l[6])),
           __fmaf_rn(d.x, l[1], __fmaf_rn(d.y, l[4], d.z * l[7])),
           __fmaf_rn(d.x, l[2], __fmaf_rn(d.y, l[5], d.z * l[8])));
}

__device__ __forceinline__ bool intersect_box(float& t0, float& t1, F3 o, F3 d, const float* lo, const float* hi) {
  const float od[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  float tmin = 0.f, tmax = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const bool small = fabsf(dd[k]) <= FLT_MIN;
    const float rcp = __frcp_rn(dd[k]);
    const float tlo = small ? VNR_FLOAT_LARGE : (lo[k] - od[k]) * rcp;
    const float thi = small ? -VNR_FLOAT_LARGE : (hi[k] - od[k]) * rcp;
    const float mn = fminf(tlo, thi), mx = fmaxf(tlo, thi);
    tmin = k == 0 ? mn : fmaxf(tmin, mn);
    tmax = k == 0 ? mx : fminf(tmax, mx);
  }
  t0 = fmaxf(t0, tmin); t1 = fminf(t1, tmax);
  return t1 > t0;
}

struct DDAState {
  float tnx, tny, tnz;     // t_next
  int cx, cy, cz;          // cell
  float ncb;               // next_cell_begin
};

// m_dir = dir * macrocell_spacings_rcp
__device__ __forceinline__ void dda_init(DDAState& s, F3 m_org, F3 m_dir, float t_min, const int* gs) {
  const F3 o = madd(t_min, m_dir, m_org);
  const float fcx = fmaxf(0.f, fminf((float)gs[0] - 1.f, floorf(o.x)));
  const float fcy = fmaxf(0.f, fminf((float)gs[1] - 1.f, floorf(o.y)));
  const float fcz = fmaxf(0.f, fminf((float)gs[2] - 1.f, floorf(o.z)));
  const float fex = m_dir.x > 0.f ? fcx + 1.f : fcx, fey = m_dir.y > 0.f ? fcy + 1.f : fcy, fez = m_dir.z > 0.f ? fcz + 1.f : fcz;
  s.tnx = m_dir.x == 0.f ? VNR_FLOAT_LARGE : fabsf(fex - o.x) * fabsf(__frcp_rn(m_dir.x));
  s.tny = m_dir.y == 0.f ? VNR_FLOAT_LARGE : fabsf(fey - o.y) * fabsf(__frcp_rn(m_dir.y));
  s.tnz = m_dir.z == 0.f ? VNR_FLOAT_LARGE : fabsf(fez - o.z) * fabsf(__frcp_rn(m_dir.z));
  s.cx = (int)fcx; s.cy = (int)fcy; s.cz = (int)fcz;
  s.ncb = 0.f;
}

__device__ __forceinline__ bool dda_resumable(const DDAState& s, F3 m_dir, float t_min, float t_max, const int* gs) {
  const int sx = m_dir.x > 0.f ? gs[0] : -1, sy = m_dir.y > 0.f ? gs[1] : -1, sz = m_dir.z > 0.f ? gs[2] : -1;
  if (s.cx == sx || s.cy == sy || s.cz == sz) return false;
  const float t_closest = fminf(s.tnx, fminf(s.tny, s.tnz));
  const float cell_t0 = fmaxf(t_min + s.ncb, t_min);
  const float cell_t1 = fminf(t_min + t_closest, t_max);
  return !(cell_t0 >= cell_t1);
}

__device__ __forceinline__ float adaptive_sampling_rate(float base, float max_opacity) {
  const float scale = 15.f * base;
  const float r = fabsf(clampf(max_opacity, 0.1f, 1.f) - 1.f);
  return fmaxf(__fmaf_rn(scale, r * r, base), base);
}

// Walk the macrocell DDA from state `s`, calling body(t0, t1) for every sample interval until
// body returns false or the ray leaves the grid.  Equivalent to
// `while (DDAIter::next(..., lambda)) {}` with the lambda of RayMarchingIter::exec.
// UNIFORM: the single-kernel marcher's raymarching_iterator (method_raymarching.cu:269-297) -- same traversal (dda3,
// dda.h:140-288), but every cell is divided into equal steps (sample_size_scaler :262-267) and `step` may be scaled.
template <bool UNIFORM = false, typename Body>
__device__ __forceinline__ void march_exec(const FrameParams& fp, DDAState& s, F3 m_dir, float tMin, float tMax, Body&& body, float step_scale = 1.f) {
  const int stopx = m_dir.x > 0.f ? fp.mc_dims[0] : -1, stopy = m_dir.y > 0.f ? fp.mc_dims[1] : -1, stopz = m_dir.z > 0.f ? fp.mc_dims[2] : -1;
  const float tsx = fabsf(__frcp_rn(m_dir.x)), tsy = fabsf(__frcp_rn(m_dir.y)), tsz = fabsf(__frcp_rn(m_dir.z));
  const int dx = m_dir.x > 0.f ? 1 : -1, dy = m_dir.y > 0.f ? 1 : -1, dz = m_dir.z > 0.f ? 1 : -1;
  for (;;) {
    if (s.cx == stopx || s.cy == stopy || s.cz == stopz) return;
    const float t_closest = fminf(s.tnx, fminf(s.tny, s.tnz));
    const float cell_t0 = fmaxf(tMin + s.ncb, tMin);
    const float cell_t1 = fminf(tMin + t_closest, tMax);
    if (cell_t0 >= cell_t1) return;
    // ---- lambda(cell, cell_t0, cell_t1)
    bool go = true;
    {
      const uint32_t idx = (uint32_t)s.cx + (uint32_t)s.cy * (uint32_t)fp.mc_dims[0] + (uint32_t)s.cz * (uint32_t)fp.mc_dims[0] * (uint32_t)fp.mc_dims[1];
      const float r = __ldg(fp.mc_maxop + idx);
      if (!(fabsf(r) <= FLT_EPSILON)) {
        float ss = adaptive_sampling_rate(UNIFORM ? step_scale * fp.step : fp.step, r);
        if (UNIFORM) {
          const int n = (int)(__fdiv_rn(cell_t1 - cell_t0, ss) + 1.f);
          ss = __fdiv_rn(cell_t1 - cell_t0, (float)n);
        }
        float tx = cell_t0, ty = fminf(cell_t1, cell_t0 + ss);
        while (ty > tx) {
          s.ncb = ty - tMin;
          if (!body(tx, ty)) { go = false; break; }
          tx = ty; ty = fminf(tx + ss, cell_t1);
        }
      }
    }
    if (go || fmaxf(tMin + s.ncb, tMin) >= cell_t1) {
      bool left = false;
      if (s.tnx == t_closest) { s.tnx += tsx; s.cx += dx; if (s.cx == stopx) left = true; }
      if (!left && s.tny == t_closest) { s.tny += tsy; s.cy += dy; if (s.cy == stopy) left = true; }
      if (!left && s.tnz == t_closest) { s.tnz += tsz; s.cz += dz; if (s.cz == stopz) left = true; }
      if (left) return;                    // DDAIter::next returns false before updating next_cell_begin
      s.ncb = t_closest;
    }
    if (!go) return;
  }
}

// 1-D table lookup with CUDA linear-filter semantics (1.8 fixed-point weight)
__device__ __forceinline__ void tfn_coeff(float v, int n, int round_mode, int& i0, int& i1, float& a) {
  v = clampf(v, 0.f, 1.f);
  const float xb = v * (float)(n - 1);
  const float fl = floorf(xb);
  const float fr = xb - fl;
  const float q = round_mode == 0 ? floorf(__fmaf_rn(fr, 256.f, 0.5f)) : floorf(fr * 256.f);
  a = q * (1.f / 256.f);
  const int i = (int)fl;
  i0 = min(max(i, 0), n - 1);
  i1 = min(max(i + 1, 0), n - 1);
}

__device__ __forceinline__ float lerp_tex(float w, float p, float q) { return __fmaf_rn(w, q, (1.f - w) * p); }

// sampleTransferFunction + opacityCorrection.  colors/alphas may point to shared or global memory.
__device__ __forceinline__ void classify(const FrameParams& fp, const float4* __restrict__ colors, const float* __restrict__ alphas,
                                         float value, float dt, float& r, float& g, float& b, float& a) {
  const float v = (clampf(value, fp.tfn_lo, fp.tfn_hi) - fp.tfn_lo) * fp.tfn_rcp;
  r = g = b = a = 0.f;
  int i0, i1; float w;
  if (fp.n_color > 0) {
    tfn_coeff(v, fp.n_color, fp.tex_round, i0, i1, w);
    const float4 c0 = colors[i0], c1 = colors[i1];
    r = lerp_tex(w, c0.x, c1.x); g = lerp_tex(w, c0.y, c1.y); b = lerp_tex(w, c0.z, c1.z);
  }
  if (fp.n_alpha > 0) {
    tfn_coeff(v, fp.n_alpha, fp.tex_round, i0, i1, w);
    a = lerp_tex(w, alphas[i0], alphas[i1]);
  }
  a = 1.f - powf(1.f - a, fp.step_rcp * dt);
}

// ---- shading (all vectors world space; raytracing.h:209-246, constants instantvnr_types.h:137-148)
#define VNR_SHADING_SCALE 0.95f
__device__ __forceinline__ float lerp1(float f, float a, float b) { return __fmaf_rn(f, b, (1.f - f) * a); }
__device__ __forceinline__ F3 normalize3(F3 a) { const float r = __fdiv_rn(1.0f, __fsqrt_rn(dot3(a, a))); return f3(r * a.x, r * a.y, r * a.z); }

__device__ __forceinline__ F3 shade_simple_light(F3 ray_dir, F3 normal, F3 albedo) {
  if (dot3(normal, normal) > 1.0e-6f) {
    const F3 n = normalize3(normal);
    const float s = __fmaf_rn(.8f, fabsf(dot3(f3(-ray_dir.x, -ray_dir.y, -ray_dir.z), n)), 0.2f);
    return f3(s * albedo.x, s * albedo.y, s * albedo.z);
  }
  return f3(0.f, 0.f, 0.f);
}

// mat = {ambient, diffuse, specular, shininess}; light_diffuse = light_directional_rgb = 1
__device__ __forceinline__ F3 shade_scivis_light(F3 ray_dir, F3 normal, F3 albedo, float m_amb, float m_dif, float m_spec, float m_shin, F3 light_dir) {
  F3 color = f3(0.f, 0.f, 0.f);
  if (dot3(normal, normal) > 1.0e-6f) {
    const F3 L = normalize3(light_dir), N = normalize3(normal), V = f3(-ray_dir.x, -ray_dir.y, -ray_dir.z);
    color = f3(color.x + m_amb * albedo.x, color.y + m_amb * albedo.y, color.z + m_amb * albedo.z);
    const float cosNL = fmaxf(dot3(N, L), 0.f);
    if (cosNL > 0.f) {
      const float dc = m_dif * cosNL;
      color = f3(color.x + dc * albedo.x, color.y + dc * albedo.y, color.z + dc * albedo.z);
      const F3 H = normalize3(f3(L.x + V.x, L.y + V.y, L.z + V.z));
      const float cosNH = fmaxf(dot3(N, H), 0.f);
      const float sp = m_spec * powf(cosNH, m_shin);
      color = f3(color.x + sp, color.y + sp, color.z + sp);
    }
  }
  const F3 s2 = shade_simple_light(ray_dir, normal, albedo);
  return f3(lerp1(0.5f, s2.x, color.x), lerp1(0.5f, s2.y, color.y), lerp1(0.5f, s2.z, color.z));
}

// GRADIENT_SHADING (method_raymarching.cu:773-788): g = forward differences / grad_step (= -No), dir_obj = object-space ray direction
__device__ __forceinline__ void shade_gradient(const FrameParams& fp, F3 dir_obj, F3 g, float& r, float& gg, float& b) {
  const F3 Nw = f3(-g.x * fp.wto_l[0], -g.y * fp.wto_l[4], -g.z * fp.wto_l[8]);          // xfmNormal(otw, No)
  const F3 dirw = f3(dir_obj.x * fp.otw_diag[0], dir_obj.y * fp.otw_diag[1], dir_obj.z * fp.otw_diag[2]);
  const F3 sc = shade_scivis_light(dirw, Nw, f3(r, gg, b), .6f, .9f, .4f, 40.f, f3(fp.light_dir));   // mat_gradient_shading
  r = lerp1(VNR_SHADING_SCALE, r, sc.x); gg = lerp1(VNR_SHADING_SCALE, gg, sc.y); b = lerp1(VNR_SHADING_SCALE, b, sc.z);
}

// shadow ray direction in object space: xfmVector(wto, normalize(light_directional_dir)) (compute_ray<SHADOW> :640-654)
__device__ __forceinline__ F3 shadow_dir(const FrameParams& fp) {
  const F3 d = normalize3(f3(fp.light_dir));
  const float* l = fp.wto_l;
  return f3(__fmaf_rn(d.x, l[0], __fmaf_rn(d.y, l[3], d.z * l[6])),
            __fmaf_rn(d.x, l[1], __fmaf_rn(d.y, l[4], d.z * l[7])),
            __fmaf_rn(d.x, l[2], __fmaf_rn(d.y, l[5], d.z * l[8])));
}

}  // namespace vnr

// synthetic.hpp -- seeded synthetic inputs for the headless drivers (no data files are needed):
// a normalised float volume of Gaussian blobs on a low-frequency background (empty space for the
// macrocell skipping, smooth features that train to > 30 dB), a 256-entry transfer function with
// zero opacity below 0.25, and the default orbit camera.
#pragma once
#include <algorithm>
#include <cmath>
#include <random>
#include <vector>

#include "../include/vnr_api.hpp"

namespace synthetic {

inline std::vector<float> make_volume(vnr::vec3i dims, unsigned seed = 42) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> uc(0.2f, 0.8f), us(0.05f, 0.15f), ua(0.5f, 1.0f);
  struct Blob { float cx, cy, cz, inv2s2, a; };
  std::vector<Blob> blobs(8);
  for (auto& b : blobs) { b.cx = uc(rng); b.cy = uc(rng); b.cz = uc(rng); const float s = us(rng); b.inv2s2 = 1.f / (2.f * s * s); b.a = ua(rng); }
  std::vector<float> v((size_t)dims.x * dims.y * dims.z);
  float lo = 1e30f, hi = -1e30f;
  for (int z = 0; z < dims.z; ++z)
    for (int y = 0; y < dims.y; ++y)
      for (int x = 0; x < dims.x; ++x) {
        const float px = (x + 0.5f) / dims.x, py = (y + 0.5f) / dims.y, pz = (z + 0.5f) / dims.z;
        float f = 0.05f * std::sin(6.2831853f * px) * std::sin(6.2831853f * py) * std::sin(6.2831853f * pz);
        for (auto& b : blobs) {
          const float d2 = (px - b.cx) * (px - b.cx) + (py - b.cy) * (py - b.cy) + (pz - b.cz) * (pz - b.cz);
          f += b.a * std::exp(-d2 * b.inv2s2);
        }
        v[((size_t)z * dims.y + y) * dims.x + x] = f;
        lo = std::min(lo, f); hi = std::max(hi, f);
      }
  const float r = 1.f / (hi - lo);
  for (auto& f : v) f = (f - lo) * r;
  return v;
}

inline vnrTransferFunction make_tfn(int n = 256) {
  auto t = vnrCreateTransferFunction();
  std::vector<vnr::vec3f> colors(n);
  std::vector<vnr::vec2f> alphas(n);
  const float stops[5][3] = {{0.23f, 0.30f, 0.75f}, {0.55f, 0.69f, 0.99f}, {0.86f, 0.86f, 0.86f}, {0.96f, 0.60f, 0.48f}, {0.70f, 0.02f, 0.15f}};
  for (int i = 0; i < n; ++i) {
    const float u = (float)i / (float)(n - 1);
    const float s = u * 4.f; const int k = std::min(3, (int)s); const float w = s - (float)k;
    colors[i] = vnr::vec3f((1 - w) * stops[k][0] + w * stops[k + 1][0], (1 - w) * stops[k][1] + w * stops[k + 1][1], (1 - w) * stops[k][2] + w * stops[k + 1][2]);
    float a = 0.f;
    if (u > 0.25f) { const float q = (u - 0.25f) / 0.75f; a = 0.8f * q * q * (3.f - 2.f * q); }
    alphas[i] = vnr::vec2f(u, a);
  }
  vnrTransferFunctionSetColor(t, colors);
  vnrTransferFunctionSetAlpha(t, alphas);
  vnrTransferFunctionSetValueRange(t, vnr::range1f(0.f, 1.f));      // batch_renderer.cpp:194
  return t;
}

// orbit view `k` of `n` around the volume centre at 1.5 x the largest dimension
inline vnrCamera orbit_camera(vnr::vec3i dims, int k, int n = 16) {
  const float r = 1.5f * (float)std::max(dims.x, std::max(dims.y, dims.z));
  const float phi = 6.2831853f * (float)k / (float)n, elev = 0.35f * std::sin(2.f * phi + 0.5f);
  auto c = vnrCreateCamera();
  vnrCameraSet(c, vnr::vec3f(r * std::cos(elev) * std::sin(phi), r * std::sin(elev), -r * std::cos(elev) * std::cos(phi)), vnr::vec3f(0, 0, 0), vnr::vec3f(0, 1, 0));
  return c;
}

}  // namespace synthetic

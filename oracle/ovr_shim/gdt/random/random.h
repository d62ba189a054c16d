// ovr_shim/gdt/random/random.h -- stand-in for gdt::LCG<N> of the un-vendored OVR framework: the TEA-initialised linear
// congruential generator of the OptiX SDK samples (tea<N>, lcg) that gdt's random.h is known to wrap; get_floats() -- the
// fork-specific call the reference makes (method_raymarching.cu:851) -- is taken to return two consecutive draws.
#pragma once
#include <cstdint>
#include "../math/vec.h"
namespace gdt {
template <unsigned int N = 16> struct LCG {
  uint32_t state;
  inline __both__ LCG() : state(0) {}
  inline __both__ LCG(unsigned int val0, unsigned int val1) { init(val0, val1); }
  inline __both__ void init(unsigned int val0, unsigned int val1) {
    unsigned int v0 = val0, v1 = val1, s0 = 0;
    for (unsigned int n = 0; n < N; n++) {
      s0 += 0x9e3779b9;
      v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
      v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
    }
    state = v0;
  }
  inline __both__ float operator()() {
    const uint32_t LCG_A = 1664525u, LCG_C = 1013904223u;
    state = (LCG_A * state + LCG_C);
    return (state & 0x00FFFFFF) / (float)0x01000000;
  }
  inline __both__ float get_float() { return (*this)(); }
  inline __both__ vec2f get_floats() { const float x = (*this)(); const float y = (*this)(); return vec2f(x, y); }
};
}  // namespace gdt

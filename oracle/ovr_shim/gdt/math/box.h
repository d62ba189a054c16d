// ovr_shim/gdt/math/box.h -- see vec.h
#pragma once
#include "vec.h"
namespace gdt {
template <typename V> struct box_t {
  V lower, upper;
  __both__ box_t() {}
  __both__ box_t(const V& l, const V& u) : lower(l), upper(u) {}
  __both__ V size() const { return upper - lower; }
};
typedef box_t<vec3f> box3f; typedef box_t<vec3i> box3i;
}  // namespace gdt

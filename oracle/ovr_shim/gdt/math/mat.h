// ovr_shim/gdt/math/mat.h -- see vec.h.  linear3f is three column vectors (vx, vy, vz), affine3f = linear part l + translation p;
// inverse() through the adjoint / determinant, xfmPoint = l * p + t, xfmVector = l * v, xfmNormal = inverse-transpose * n -- the
// conventions of the embree-derived gdt library the reference was written against (field names as used in raytracing.h:44-66).
#pragma once
#include "vec.h"
namespace gdt {
struct linear3f {
  vec3f vx, vy, vz;
  __both__ linear3f() {}
  __both__ linear3f(const vec3f& x, const vec3f& y, const vec3f& z) : vx(x), vy(y), vz(z) {}
  __both__ float det() const { return dot(vx, cross(vy, vz)); }
  __both__ linear3f adjoint() const { return linear3f(cross(vy, vz), cross(vz, vx), cross(vx, vy)).transposed(); }
  __both__ linear3f transposed() const { return linear3f(vec3f(vx.x, vy.x, vz.x), vec3f(vx.y, vy.y, vz.y), vec3f(vx.z, vy.z, vz.z)); }
  __both__ linear3f inverse() const { const linear3f a = adjoint(); const float d = det(); return linear3f(a.vx / d, a.vy / d, a.vz / d); }
  static __both__ linear3f scale(const vec3f& s) { return linear3f(vec3f(s.x, 0, 0), vec3f(0, s.y, 0), vec3f(0, 0, s.z)); }
  static __both__ linear3f identity() { return scale(vec3f(1.f)); }
};
inline __both__ vec3f operator*(const linear3f& l, const vec3f& v) { return v.x * l.vx + v.y * l.vy + v.z * l.vz; }
inline __both__ linear3f operator*(const linear3f& a, const linear3f& b) { return linear3f(a * b.vx, a * b.vy, a * b.vz); }
struct affine3f {
  linear3f l; vec3f p;
  __both__ affine3f() : l(linear3f::identity()), p(0.f) {}
  __both__ affine3f(const linear3f& l_, const vec3f& p_) : l(l_), p(p_) {}
  __both__ affine3f inverse() const { const linear3f il = l.inverse(); return affine3f(il, -(il * p)); }
  static __both__ affine3f scale(const vec3f& s) { return affine3f(linear3f::scale(s), vec3f(0.f)); }
  static __both__ affine3f translate(const vec3f& t) { return affine3f(linear3f::identity(), t); }
};
inline __both__ affine3f operator*(const affine3f& a, const affine3f& b) { return affine3f(a.l * b.l, a.l * b.p + a.p); }
inline __both__ vec3f xfmPoint(const affine3f& m, const vec3f& p) { return m.l * p + m.p; }
inline __both__ vec3f xfmVector(const affine3f& m, const vec3f& v) { return m.l * v; }
inline __both__ vec3f xfmNormal(const affine3f& m, const vec3f& n) { return m.l.inverse().transposed() * n; }
}  // namespace gdt

// ovr_shim/gdt/math/vec.h -- stand-in for the vector types of the un-vendored OVR framework (VIDILabs/open-volume-renderer,
// gdt/math/*.h), written from the way the reference's sources USE them (core/mathdef.h:24-46, core/renderer/*.h): plain
// component-wise fp32 / int32 vectors.  TEST INFRASTRUCTURE: lets the reference's own marcher sources compile in place
// (oracle/ref_marcher/); nothing in the product includes this.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include <algorithm>
#include <limits>
#include <iostream>

#ifndef __both__
#define __both__ __host__ __device__
#endif

namespace gdt {

inline __both__ float rcp(float x) { return 1.f / x; }
// scalar min / max are templates so that CUDA's own ::min / ::max (non-template) win overload resolution where the reference
// calls them unqualified under `using namespace gdt` (dda.h:83-84)
template <typename T> inline __both__ T min(const T& a, const T& b) { return a < b ? a : b; }
template <typename T> inline __both__ T max(const T& a, const T& b) { return a > b ? a : b; }
template <typename T> inline __both__ T clamp(T v, T lo = T(0), T hi = T(1)) { return v < lo ? lo : (v > hi ? hi : v); }
inline __both__ float floor(float v) { return ::floorf(v); }
inline __both__ float ceil(float v) { return ::ceilf(v); }

template <typename T> struct vec2 {
  T x, y;
  __both__ vec2() {}
  __both__ vec2(T s) : x(s), y(s) {}
  __both__ vec2(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> __both__ explicit vec2(const vec2<U>& o) : x((T)o.x), y((T)o.y) {}
  __both__ vec2(const float2& f) : x((T)f.x), y((T)f.y) {}
  __both__ T long_product() const { return x * y; }
};
template <typename T> struct vec4;
template <typename T> struct vec3 {
  T x, y, z;
  __both__ explicit vec3(const vec4<T>& v);
  __both__ vec3() {}
  __both__ vec3(T s) : x(s), y(s), z(s) {}
  __both__ vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  template <typename U> __both__ explicit vec3(const vec3<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
  __both__ vec3(const float3& f) : x((T)f.x), y((T)f.y), z((T)f.z) {}
  __both__ size_t long_product() const { return (size_t)x * (size_t)y * (size_t)z; }
  __both__ T& operator[](int i) { return (&x)[i]; }
  __both__ const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct vec4 {
  T x, y, z, w;
  __both__ vec4() {}
  __both__ vec4(T s) : x(s), y(s), z(s), w(s) {}
  __both__ vec4(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
  __both__ vec4(const vec3<T>& v, T w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
  __both__ vec4(const float4& f) : x((T)f.x), y((T)f.y), z((T)f.z), w((T)f.w) {}
  // the reference ASSIGNS through xyz() (`shadingColor.xyz() = lerp(...)`, method_raymarching.cu:824: the blend of the single shade
  // into the pixel): the accessor must alias the first three components, or the shadow pass would have no effect
  __both__ vec3<T>& xyz() { return *reinterpret_cast<vec3<T>*>(this); }
  __both__ const vec3<T>& xyz() const { return *reinterpret_cast<const vec3<T>*>(this); }
};
template <typename T> __both__ inline vec3<T>::vec3(const vec4<T>& v) : x(v.x), y(v.y), z(v.z) {}
typedef vec2<float> vec2f; typedef vec2<int> vec2i;
typedef vec3<float> vec3f; typedef vec3<int> vec3i;
typedef vec4<float> vec4f; typedef vec4<int> vec4i;

#define GDT_BINOP(op)                                                                                                              \
  template <typename T> inline __both__ vec2<T> operator op(const vec2<T>& a, const vec2<T>& b) { return vec2<T>(a.x op b.x, a.y op b.y); } \
  template <typename T> inline __both__ vec2<T> operator op(const vec2<T>& a, T b) { return vec2<T>(a.x op b, a.y op b); }               \
  template <typename T> inline __both__ vec2<T> operator op(T a, const vec2<T>& b) { return vec2<T>(a op b.x, a op b.y); }               \
  template <typename T> inline __both__ vec3<T> operator op(const vec3<T>& a, const vec3<T>& b) { return vec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
  template <typename T> inline __both__ vec3<T> operator op(const vec3<T>& a, T b) { return vec3<T>(a.x op b, a.y op b, a.z op b); }     \
  template <typename T> inline __both__ vec3<T> operator op(T a, const vec3<T>& b) { return vec3<T>(a op b.x, a op b.y, a op b.z); }     \
  template <typename T> inline __both__ vec4<T> operator op(const vec4<T>& a, const vec4<T>& b) { return vec4<T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
  template <typename T> inline __both__ vec4<T> operator op(const vec4<T>& a, T b) { return vec4<T>(a.x op b, a.y op b, a.z op b, a.w op b); } \
  template <typename T> inline __both__ vec4<T> operator op(T a, const vec4<T>& b) { return vec4<T>(a op b.x, a op b.y, a op b.z, a op b.w); }
GDT_BINOP(+) GDT_BINOP(-) GDT_BINOP(*) GDT_BINOP(/)
#undef GDT_BINOP
// mixed scalar types the reference writes (int * vec3f, double literal * vec3f)
inline __both__ vec3f operator*(int a, const vec3f& b) { return (float)a * b; }
inline __both__ vec3f operator*(const vec3f& a, int b) { return a * (float)b; }
inline __both__ vec3f operator*(double a, const vec3f& b) { return (float)a * b; }
inline __both__ vec3f operator*(const vec3f& a, double b) { return a * (float)b; }
inline __both__ vec3f operator/(const vec3f& a, int b) { return a / (float)b; }
inline __both__ vec4f operator*(double a, const vec4f& b) { return (float)a * b; }
inline __both__ vec4f operator/(const vec4f& a, int b) { return a / (float)b; }
inline __both__ vec2f operator/(const vec2f& a, const vec2i& b) { return vec2f(a.x / (float)b.x, a.y / (float)b.y); }
inline __both__ vec3f operator-(const vec3f& a, double b) { return a - (float)b; }

#define GDT_ASSIGN(op)                                                                                                             \
  template <typename T> inline __both__ vec3<T>& operator op##=(vec3<T>& a, const vec3<T>& b) { a = a op b; return a; }                 \
  template <typename T> inline __both__ vec3<T>& operator op##=(vec3<T>& a, T b) { a = a op b; return a; }                              \
  template <typename T> inline __both__ vec4<T>& operator op##=(vec4<T>& a, const vec4<T>& b) { a = a op b; return a; }                 \
  template <typename T> inline __both__ vec4<T>& operator op##=(vec4<T>& a, T b) { a = a op b; return a; }                              \
  template <typename T> inline __both__ vec2<T>& operator op##=(vec2<T>& a, const vec2<T>& b) { a = a op b; return a; }
GDT_ASSIGN(+) GDT_ASSIGN(-) GDT_ASSIGN(*) GDT_ASSIGN(/)
#undef GDT_ASSIGN
inline __both__ vec3f& operator*=(vec3f& a, int b) { a = a * (float)b; return a; }

template <typename T> inline __both__ vec3<T> operator-(const vec3<T>& a) { return vec3<T>(-a.x, -a.y, -a.z); }
template <typename T> inline __both__ bool operator==(const vec3<T>& a, const vec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> inline __both__ bool operator!=(const vec3<T>& a, const vec3<T>& b) { return !(a == b); }
template <typename T> inline __both__ bool operator==(const vec2<T>& a, const vec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> inline __both__ bool operator!=(const vec2<T>& a, const vec2<T>& b) { return !(a == b); }

inline __both__ float dot(const vec3f& a, const vec3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline __both__ vec3f cross(const vec3f& a, const vec3f& b) { return vec3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline __both__ float length(const vec3f& a) { return sqrtf(dot(a, a)); }
inline __both__ vec3f normalize(const vec3f& a) { return a * (1.f / sqrtf(dot(a, a))); }
inline __both__ vec3f min(const vec3f& a, const vec3f& b) { return vec3f(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
inline __both__ vec3f max(const vec3f& a, const vec3f& b) { return vec3f(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
inline __both__ vec3i min(const vec3i& a, const vec3i& b) { return vec3i(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }
inline __both__ vec3i max(const vec3i& a, const vec3i& b) { return vec3i(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }
inline __both__ vec3f clamp(const vec3f& v, const vec3f& lo, const vec3f& hi) { return min(max(v, lo), hi); }
inline __both__ float reduce_min(const vec3f& a) { return fminf(a.x, fminf(a.y, a.z)); }
inline __both__ float reduce_max(const vec3f& a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
inline __both__ vec3f rcp(const vec3f& a) { return vec3f(1.f / a.x, 1.f / a.y, 1.f / a.z); }
inline __both__ vec3f abs(const vec3f& a) { return vec3f(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
}  // namespace gdt
namespace util { inline gdt::vec3i div_round_up(const gdt::vec3i& a, const gdt::vec3i& b) { return gdt::vec3i((a.x + b.x - 1) / b.x, (a.y + b.y - 1) / b.y, (a.z + b.z - 1) / b.z); } }
namespace gdt {
template <typename T> inline std::ostream& operator<<(std::ostream& o, const vec3<T>& v) { return o << "(" << v.x << "," << v.y << "," << v.z << ")"; }
template <typename T> inline std::ostream& operator<<(std::ostream& o, const vec2<T>& v) { return o << "(" << v.x << "," << v.y << ")"; }

template <typename T> struct interval {
  union { T lower; T lo; };          // the reference reads both spellings (macrocell.cu:149-150, raytracing.h:150)
  union { T upper; T hi; };
  __both__ interval() : lower(std::numeric_limits<T>::max()), upper(std::numeric_limits<T>::lowest()) {}
  __both__ interval(T l, T u) : lower(l), upper(u) {}
  __both__ void extend(T v) { lower = v < lower ? v : lower; upper = v > upper ? v : upper; }
  __both__ bool is_empty() const { return upper < lower; }
  __both__ T span() const { return upper - lower; }
};
typedef interval<float> range1f; typedef interval<int> range1i;
template <typename T> inline std::ostream& operator<<(std::ostream& o, const interval<T>& v) { return o << "[" << v.lower << "," << v.upper << "]"; }

}  // namespace gdt

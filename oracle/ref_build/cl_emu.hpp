// cl_emu.hpp — the minimum of OpenCL C 1.2 needed to compile the reference's UNMODIFIED kernel sources
// (/root/reference/opencl_kernels/*.cl) as C++ for the host CPU.
//
// TEST INFRASTRUCTURE (oracle/_ref): never linked into the product.  The kernels themselves are not copied into this
// repository — oracle/ref_build/build_ref.py reads them where they lie, expands the reference's own
// `#clw_include_once` directive (opencl_wrapper/include/clw_function.hpp:23-72) into oracle/_ref/ (git-ignored) and
// compiles them against this header.
//
// What lives in the vendor OpenCL driver and therefore has to be defined here (the same definitions the restatement in
// oracle/oracle.cpp uses):
//   * vector types with component-wise operators, scalar widening and brace construction
//   * geometric built-ins: dot = (x*x' + y*y') + z*z', length = sqrt(dot), normalize = v / length (0 for the zero vector)
//   * min/max as (b < a ? b : a) / (a < b ? b : a); clamp; abs; round (half away from zero)
//   * images: integer reads are NEAREST (texel floor(coord)); CLK_ADDRESS_CLAMP reads outside the image return 0;
//     CLK_ADDRESS_CLAMP_TO_EDGE clamps; normalised coordinates scale by the image size first
//   * 32-bit atomics on __global memory; get_global_id via thread-local state set by the NDRange loop in ref_driver
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <sys/types.h>
#include <type_traits>

#define __kernel
#define __global
#define __constant const
#define __read_only
#define __write_only
#define __read_write

typedef unsigned char uchar;
typedef unsigned short ushort;

// ---- vector types ------------------------------------------------------------------------------------------------
#define EMU_ARITH(A) typename std::enable_if<std::is_arithmetic<A>::value || std::is_enum<A>::value, int>::type = 0

// scalar conversion used by the vector constructors: float -> integer is round-toward-zero, saturating, NaN -> 0 (what
// the hardware's conversion instruction does and what oracle.cpp defines); everything else is the plain C conversion
template <class T, class A>
inline T emu_cv(A a) {
  if constexpr (std::is_integral<T>::value && std::is_floating_point<A>::value) {
    if (a != a) return (T)0;
    const double d = (double)a, lo = (double)std::numeric_limits<T>::min(), hi = (double)std::numeric_limits<T>::max();
    return d <= lo ? std::numeric_limits<T>::min() : (d >= hi ? std::numeric_limits<T>::max() : (T)d);
  } else {
    return (T)a;
  }
}

template <class T>
struct vec2 {
  T x, y;
  vec2() : x(0), y(0) {}
  template <class A, EMU_ARITH(A)> vec2(A a) : x(emu_cv<T>(a)), y(emu_cv<T>(a)) {}
  template <class A, class B> vec2(A a, B b) : x(emu_cv<T>(a)), y(emu_cv<T>(b)) {}
  template <class U> explicit vec2(const vec2<U>& o) : x((T)o.x), y((T)o.y) {}
};
template <class T>
struct vec3 {
  T x, y, z;
  vec3() : x(0), y(0), z(0) {}
  template <class A, EMU_ARITH(A)> vec3(A a) : x(emu_cv<T>(a)), y(emu_cv<T>(a)), z(emu_cv<T>(a)) {}
  template <class A, class B, class C> vec3(A a, B b, C c) : x(emu_cv<T>(a)), y(emu_cv<T>(b)), z(emu_cv<T>(c)) {}
};
template <class T>
struct vec4 {
  T x, y, z, w;
  vec4() : x(0), y(0), z(0), w(0) {}
  template <class A, EMU_ARITH(A)> vec4(A a) : x(emu_cv<T>(a)), y(emu_cv<T>(a)), z(emu_cv<T>(a)), w(emu_cv<T>(a)) {}
  template <class A, class B, class C, class D> vec4(A a, B b, C c, D d) : x(emu_cv<T>(a)), y(emu_cv<T>(b)), z(emu_cv<T>(c)), w(emu_cv<T>(d)) {}
};
typedef vec2<float> float2;
typedef vec3<float> float3;
typedef vec4<float> float4;
typedef vec2<int> int2;
typedef vec3<int> int3;
typedef vec4<int> int4;
typedef vec2<unsigned> uint2;
typedef vec4<unsigned> uint4;

#define EMU_OPS2(OP)                                                                                               \
  template <class T> inline vec2<T> operator OP(vec2<T> a, vec2<T> b) { return vec2<T>(a.x OP b.x, a.y OP b.y); }  \
  template <class T> inline vec3<T> operator OP(vec3<T> a, vec3<T> b) { return vec3<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
  template <class T> inline vec4<T> operator OP(vec4<T> a, vec4<T> b) { return vec4<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
  template <class T, class S, EMU_ARITH(S)> inline vec2<T> operator OP(vec2<T> a, S s) { return a OP vec2<T>(s); } \
  template <class T, class S, EMU_ARITH(S)> inline vec3<T> operator OP(vec3<T> a, S s) { return a OP vec3<T>(s); } \
  template <class T, class S, EMU_ARITH(S)> inline vec4<T> operator OP(vec4<T> a, S s) { return a OP vec4<T>(s); } \
  template <class T, class S, EMU_ARITH(S)> inline vec2<T> operator OP(S s, vec2<T> a) { return vec2<T>(s) OP a; } \
  template <class T, class S, EMU_ARITH(S)> inline vec3<T> operator OP(S s, vec3<T> a) { return vec3<T>(s) OP a; } \
  template <class T, class S, EMU_ARITH(S)> inline vec4<T> operator OP(S s, vec4<T> a) { return vec4<T>(s) OP a; } \
  template <class T, class R> inline vec2<T>& operator OP##=(vec2<T>& a, R b) { a = a OP b; return a; }            \
  template <class T, class R> inline vec3<T>& operator OP##=(vec3<T>& a, R b) { a = a OP b; return a; }            \
  template <class T, class R> inline vec4<T>& operator OP##=(vec4<T>& a, R b) { a = a OP b; return a; }
EMU_OPS2(+)
EMU_OPS2(-)
EMU_OPS2(*)
EMU_OPS2(/)
template <class T> inline vec3<T> operator-(vec3<T> a) { return vec3<T>(-a.x, -a.y, -a.z); }
template <class T> inline vec4<T> operator-(vec4<T> a) { return vec4<T>(-a.x, -a.y, -a.z, -a.w); }

// `(float4)(value, other)` (utility.cl:3): in C++ the parenthesised pair is a comma expression, so give it the meaning
// the OpenCL vector literal has.
inline float4 operator,(const float3& v, float w) { return float4(v.x, v.y, v.z, w); }

// ---- math built-ins -------------------------------------------------------------------------------------------------
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline float max(float a, double b) { return max(a, (float)b); }
inline float min(float a, double b) { return min(a, (float)b); }
inline int4 clamp(int4 v, int4 lo, int4 hi) {
  return int4(min(max(v.x, lo.x), hi.x), min(max(v.y, lo.y), hi.y), min(max(v.z, lo.z), hi.z), min(max(v.w, lo.w), hi.w));
}
inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(float3 a) { return sqrtf(dot(a, a)); }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float3 normalize(float3 a) {
  float l = length(a);
  if (l == 0.0f) return float3(0.0f, 0.0f, 0.0f);
  return float3(a.x / l, a.y / l, a.z / l);
}
using std::abs;
using std::asin;
using std::atan;
using std::atan2;
using std::cos;
using std::exp;
using std::fabs;
using std::floor;
using std::ldexp;
using std::pow;
using std::round;
using std::sin;

// ---- work-item functions -----------------------------------------------------------------------------------------------
struct emu_ndrange { size_t gid[3]; };
extern thread_local emu_ndrange emu_wi;
inline size_t get_global_id(unsigned d) { return emu_wi.gid[d]; }
inline size_t get_local_id(unsigned d) { return 0; }

// ---- images ---------------------------------------------------------------------------------------------------------------
enum { EMU_S8, EMU_S16, EMU_U8x4 };
struct emu_image {
  void* data;
  int w, h, d;
  int fmt;
  // __read_write images (2d_image_filter.cl:6): when set, writes go here and reads keep seeing `data` — "every work-item
  // reads the frame as it was before the launch", the one scheduling-independent meaning of the in-place kernel
  void* wdata = nullptr;
};
typedef emu_image* image2d_t;
typedef emu_image* image3d_t;
typedef int sampler_t;
enum {
  CLK_FILTER_NEAREST = 0, CLK_FILTER_LINEAR = 1,
  CLK_ADDRESS_NONE = 0, CLK_ADDRESS_CLAMP = 2, CLK_ADDRESS_CLAMP_TO_EDGE = 4,
  CLK_NORMALIZED_COORDS_FALSE = 0, CLK_NORMALIZED_COORDS_TRUE = 8
};
inline int get_image_width(const emu_image* i) { return i->w; }
inline int get_image_height(const emu_image* i) { return i->h; }
inline int get_image_depth(const emu_image* i) { return i->d; }
struct emu_dim {  // get_image_dim: int4 for 3-D images, int2 for 2-D ones
  const emu_image* i;
  operator int4() const { return int4(i->w, i->h, i->d, 0); }
  operator int2() const { return int2(i->w, i->h); }
};
inline emu_dim get_image_dim(const emu_image* i) { return emu_dim{i}; }

inline int emu_fetch3(const emu_image* im, long x, long y, long z) {
  if (x < 0 || y < 0 || z < 0 || x >= im->w || y >= im->h || z >= im->d) return 0;  // border colour
  size_t idx = (size_t)x + (size_t)im->w * ((size_t)y + (size_t)im->h * (size_t)z);
  return im->fmt == EMU_S16 ? (int)((const int16_t*)im->data)[idx] : (int)((const int8_t*)im->data)[idx];
}
inline long emu_floor(float f) { return (f != f) ? 0 : (long)floorf(f); }
// read_imagei(image3d, sampler, float4): NEAREST, texel = floor(coord)
inline int4 read_imagei(const emu_image* im, sampler_t, float4 c) {
  return int4(emu_fetch3(im, emu_floor(c.x), emu_floor(c.y), emu_floor(c.z)), 0, 0, 1);
}
// read_imagei(image3d, sampler, int4) and the sampler-less form
inline int4 read_imagei(const emu_image* im, sampler_t, int4 c) { return int4(emu_fetch3(im, c.x, c.y, c.z), 0, 0, 1); }
inline int4 read_imagei(const emu_image* im, int4 c) { return int4(emu_fetch3(im, c.x, c.y, c.z), 0, 0, 1); }
// read_imageui(image2d RGBA8, sampler, float2)
inline uint4 read_imageui(const emu_image* im, sampler_t s, float2 c) {
  float u = c.x, v = c.y;
  if (s & CLK_NORMALIZED_COORDS_TRUE) { u = u * (float)im->w; v = v * (float)im->h; }
  long ix = emu_floor(u), iy = emu_floor(v);
  if (s & CLK_ADDRESS_CLAMP_TO_EDGE) {
    ix = ix < 0 ? 0 : (ix > im->w - 1 ? im->w - 1 : ix);
    iy = iy < 0 ? 0 : (iy > im->h - 1 ? im->h - 1 : iy);
  } else if (ix < 0 || iy < 0 || ix >= im->w || iy >= im->h) {
    return uint4(0, 0, 0, 0);
  }
  const uint8_t* p = (const uint8_t*)im->data + 4 * ((size_t)iy * im->w + ix);
  return uint4(p[0], p[1], p[2], p[3]);
}
// read_imageui(image2d RGBA8, sampler, int2): integer coordinates are used as they are; CLK_ADDRESS_CLAMP outside = border 0
inline uint4 read_imageui(const emu_image* im, sampler_t s, int2 c) {
  long ix = c.x, iy = c.y;
  if (s & CLK_ADDRESS_CLAMP_TO_EDGE) {
    ix = ix < 0 ? 0 : (ix > im->w - 1 ? im->w - 1 : ix);
    iy = iy < 0 ? 0 : (iy > im->h - 1 ? im->h - 1 : iy);
  } else if (ix < 0 || iy < 0 || ix >= im->w || iy >= im->h) {
    return uint4(0, 0, 0, 0);
  }
  const uint8_t* p = (const uint8_t*)im->data + 4 * ((size_t)iy * im->w + ix);
  return uint4(p[0], p[1], p[2], p[3]);
}
inline int emu_sat(long v, long lo, long hi) { return (int)(v < lo ? lo : (v > hi ? hi : v)); }
inline void write_imagei(emu_image* im, int4 c, int4 v) {
  if (c.x < 0 || c.y < 0 || c.x >= im->w || c.y >= im->h) return;
  if (im->fmt == EMU_U8x4) {  // tf_flush_color_frame writes int4 into an RGBA8 image (histogram.cl:66-68)
    uint8_t* p = (uint8_t*)im->data + 4 * ((size_t)c.y * im->w + c.x);
    p[0] = (uint8_t)emu_sat(v.x, 0, 255); p[1] = (uint8_t)emu_sat(v.y, 0, 255);
    p[2] = (uint8_t)emu_sat(v.z, 0, 255); p[3] = (uint8_t)emu_sat(v.w, 0, 255);
    return;
  }
  if (c.z < 0 || c.z >= im->d) return;
  size_t idx = (size_t)c.x + (size_t)im->w * ((size_t)c.y + (size_t)im->h * (size_t)c.z);
  if (im->fmt == EMU_S16) ((int16_t*)im->data)[idx] = (int16_t)emu_sat(v.x, -32768, 32767);
  else ((int8_t*)im->data)[idx] = (int8_t)emu_sat(v.x, -128, 127);
}
inline void write_imagei(emu_image* im, int2 c, int4 v) { write_imagei(im, int4(c.x, c.y, 0, 0), v); }
inline void write_imageui(emu_image* im, int2 c, uint4 v) {
  if (c.x < 0 || c.y < 0 || c.x >= im->w || c.y >= im->h) return;
  uint8_t* p = (uint8_t*)(im->wdata ? im->wdata : im->data) + 4 * ((size_t)c.y * im->w + c.x);
  p[0] = (uint8_t)min(v.x, 255u); p[1] = (uint8_t)min(v.y, 255u); p[2] = (uint8_t)min(v.z, 255u); p[3] = (uint8_t)min(v.w, 255u);
}

// ---- atomics ------------------------------------------------------------------------------------------------------------------
template <class T, class V> inline T atomic_add(T* p, V v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
template <class T, class V> inline T atomic_sub(T* p, V v) { return __atomic_fetch_sub(p, (T)v, __ATOMIC_RELAXED); }
template <class T> inline T atomic_inc(T* p) { return __atomic_fetch_add(p, (T)1, __ATOMIC_RELAXED); }
inline int atomic_min(int* p, int v) {
  int o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
inline int atomic_max(int* p, int v) {
  int o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}

// ---- the run-time generated transfer function (app/ui.cpp:160-168, app/tf_part.cpp:55-79) --------------------------------------
// The reference prepends generated source text; here the same clauses are evaluated from a table.
struct emu_tf_rect {
  float min_v, max_v, min_g, max_g;
  int32_t flags;  // 1: gradient clause present, 2: `return (value > min_v);`
  int32_t rgba[4];
};
struct emu_tf { const emu_tf_rect* r; int n; };
extern thread_local emu_tf emu_tf_cur;
inline bool is_event_gen(short value, short gradient, int4* color) {
  for (int i = 0; i < emu_tf_cur.n; ++i) {
    const emu_tf_rect& q = emu_tf_cur.r[i];
    if (q.flags & 2) return (value > q.min_v);
    if (value >= q.min_v && value <= q.max_v && (!(q.flags & 1) || (gradient > q.min_g && gradient < q.max_g))) {
      int4 tmp_color = {q.rgba[0], q.rgba[1], q.rgba[2], q.rgba[3]};
      *color = tmp_color;
      return true;
    }
  }
  return false;
}
// signed_distance_field.cl:19-20,38 passes a uint4* 
inline bool is_event_gen(short value, short gradient, uint4* color) {
  int4 c(color->x, color->y, color->z, color->w);
  bool r = is_event_gen(value, gradient, &c);
  *color = uint4(c.x, c.y, c.z, c.w);
  return r;
}

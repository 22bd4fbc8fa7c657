// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// Minimal OpenCL-C -> C++17 compatibility header.  It lets g++ compile the
// reference's OpenCL kernel sources *where they lie* under /root/reference
// (kernel_ASOC.c, kernel_ASOC_map.c, kernel_ASOC_sca.c plus the files they
// include) so that the reference algorithm itself can be executed on the host
// cores as the parity oracle (oracle/_ref/).  Nothing of the reference text is
// copied here: this file only supplies the language features OpenCL C has and
// C++ lacks (address-space qualifiers, vector types, built-ins, work-item ids).
//
// Semantics kept deliberately:
//   * float built-ins stay single precision (std:: overloads), double literals
//     promote to double exactly as in OpenCL C;
//   * no -ffast-math anywhere: octree links are stored as denormal floats;
//   * float3 here is 12 bytes (OpenCL's is 16) -- harnesses convert.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstddef>
#include <algorithm>

#define __kernel
#define __global
#define __constant const
#define constant   const
#define __local
#define __private

typedef unsigned int   uint;
typedef unsigned long  ulong;
typedef unsigned short ushort;
typedef unsigned char  uchar;

// ---- work-item identity: set by the OpenMP loop of the harness --------------
extern thread_local size_t clshim_gid;
extern thread_local size_t clshim_gsize;
static inline size_t get_global_id(int)   { return clshim_gid; }
static inline size_t get_global_size(int) { return clshim_gsize; }
static inline size_t get_local_id(int)    { return clshim_gid & 7; }

// ---- vector types ------------------------------------------------------------
template <typename T> struct clvec3 {
    T x, y, z;
    clvec3() = default;
    clvec3(T s) : x(s), y(s), z(s) {}
    clvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename U> clvec3(const clvec3<U> &o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    clvec3 &operator+=(const clvec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    clvec3 &operator-=(const clvec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    clvec3 &operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
    clvec3 &operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
    clvec3 operator-() const { return clvec3(-x, -y, -z); }
};
typedef clvec3<float>  float3;
typedef clvec3<double> double3;

template <typename T> static inline clvec3<T> operator+(clvec3<T> a, const clvec3<T> &b) { a += b; return a; }
template <typename T> static inline clvec3<T> operator-(clvec3<T> a, const clvec3<T> &b) { a -= b; return a; }
static inline float3  operator*(float s, const float3 &v)  { return float3(s * v.x, s * v.y, s * v.z); }
static inline float3  operator*(const float3 &v, float s)  { return float3(s * v.x, s * v.y, s * v.z); }
static inline float3  operator*(int s, const float3 &v)    { return float3(s * v.x, s * v.y, s * v.z); }
static inline float3  operator*(double s, const float3 &v) { return float3((float)(s * v.x), (float)(s * v.y), (float)(s * v.z)); }
static inline double3 operator*(double s, const double3 &v){ return double3(s * v.x, s * v.y, s * v.z); }
static inline double3 operator*(const double3 &v, double s){ return double3(s * v.x, s * v.y, s * v.z); }
static inline float3  operator/(const float3 &v, float s)  { return float3(v.x / s, v.y / s, v.z / s); }

struct uint2 { uint x, y; uint2() = default; uint2(uint a, uint b) : x(a), y(b) {} };
struct int2  { int  x, y; int2()  = default; int2(int a, int b)   : x(a), y(b) {} };

// ---- scalar built-ins ---------------------------------------------------------
using std::sqrt; using std::exp; using std::log; using std::log10; using std::pow;
using std::sin;  using std::cos; using std::acos; using std::atan2; using std::fabs;
using std::floor; using std::fmod; using std::ldexp; using std::round; using std::isfinite;
using std::expm1; using std::log1p;

static inline float  min(float a, float b)    { return a < b ? a : b; }
static inline float  max(float a, float b)    { return a > b ? a : b; }
static inline double min(double a, double b)  { return a < b ? a : b; }
static inline double max(double a, double b)  { return a > b ? a : b; }
static inline double min(double a, float b)   { return a < b ? a : (double)b; }
static inline double min(float a, double b)   { return a < b ? (double)a : b; }
static inline double max(double a, float b)   { return a > b ? a : (double)b; }
static inline double max(float a, double b)   { return a > b ? (double)a : b; }
static inline int    min(int a, int b)        { return a < b ? a : b; }
static inline int    max(int a, int b)        { return a > b ? a : b; }
static inline float  clamp(float v, float lo, float hi)    { return min(max(v, lo), hi); }
static inline double clamp(double v, double lo, double hi) { return min(max(v, lo), hi); }
static inline int    clamp(int v, int lo, int hi)          { return min(max(v, lo), hi); }
static inline float  clamp(float v, float lo, double hi)   { return min(max(v, lo), (float)hi); }
static inline float  clamp(float v, double lo, double hi)  { return min(max(v, (float)lo), (float)hi); }

static inline float  pown(float x, int n)  { return std::pow(x, (float)n); }
static inline double pown(double x, int n) { return std::pow(x, (double)n); }
static inline float  sincos(float a, float *c) { *c = std::cos(a); return std::sin(a); }
static inline uint   mad_hi(uint a, uint b, uint c) { return (uint)(((ulong)a * (ulong)b) >> 32) + c; }

static inline float native_sqrt(float x) { return std::sqrt(x); }
static inline float native_exp(float x)  { return std::exp(x); }
static inline float native_log(float x)  { return std::log(x); }
static inline float native_sin(float x)  { return std::sin(x); }
static inline float native_cos(float x)  { return std::cos(x); }

// ---- geometric built-ins --------------------------------------------------------
static inline float  dot(const float3 &a, const float3 &b)   { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float  length(const float3 &a)                 { return std::sqrt(dot(a, a)); }
static inline float  distance(const float3 &a, const float3 &b) { return length(a - b); }
static inline float3 normalize(const float3 &a)              { float l = length(a); return float3(a.x / l, a.y / l, a.z / l); }

// ---- atomics ----------------------------------------------------------------------
static inline unsigned int atomic_cmpxchg(volatile unsigned int *p, unsigned int cmp, unsigned int val) {
    return __sync_val_compare_and_swap(p, cmp, val);
}

// half precision storage is never enabled in the oracle builds (OPT_IS_HALF=0)
typedef unsigned short half;
static inline float vload_half(long, const half *) { return 0.0f; }

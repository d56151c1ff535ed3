// xs_math.cuh — float32 arithmetic with the reference's exact semantics.
//
// The reference binary (g++ -O3, x86-64 baseline) evaluates every expression in IEEE binary32 with
// round-to-nearest, no FMA contraction, correctly-rounded sqrt/div and half-away-from-zero rounding
// (SURVEY.md §7.4).  Every product/sum below therefore goes through __fmul_rn/__fadd_rn/__fdiv_rn/
// __fsqrt_rn, which nvcc never contracts or approximates regardless of -fmad / -prec-* / -ftz flags.
// Reference semantics mirrored: /root/reference/src/vec.hpp:8-142 (Vec3).
//
// The header also compiles as plain C++ (XS_HOST_SIM) so tests can run the same source on the CPU
// here; that build is test infrastructure only and is never part of libxs3d_b200.so.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define XS_HD __host__ __device__ __forceinline__
#define XS_D __device__ __forceinline__
#else
#define XS_HD inline
#define XS_D inline
#endif

namespace xs {

XS_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
XS_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
XS_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
XS_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
XS_HD float fsqrt(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
// std::round: nearest integer, halfway cases away from zero (vec.hpp:99).  |x - trunc(x)| is exact.
XS_HD float fround(float x) {
  float t = truncf(x);
  float f = fabsf(fsub(x, t));
  return (f >= 0.5f) ? fadd(t, copysignf(1.0f, x)) : t;
}

struct V3 {
  float x, y, z;
};
XS_HD V3 mk(float x, float y, float z) {
  V3 r;
  r.x = x; r.y = y; r.z = z;
  return r;
}
XS_HD V3 vadd(V3 a, V3 b) { return mk(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
XS_HD V3 vsub(V3 a, V3 b) { return mk(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
XS_HD V3 vmul(V3 a, V3 b) { return mk(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
XS_HD V3 vscale(V3 a, float s) { return mk(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
XS_HD V3 vdivs(V3 a, float s) { return mk(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
// vec.hpp:93 — (x*ox + y*oy) + z*oz
XS_HD float vdot(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
XS_HD float vnorm2(V3 a) { return vdot(a, a); }
XS_HD float vnorm(V3 a) { return fsqrt(vnorm2(a)); }
XS_HD V3 vcross(V3 a, V3 b) {
  return mk(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
            fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
XS_HD V3 vround(V3 a) { return mk(fround(a.x), fround(a.y), fround(a.z)); }
XS_HD float vcomp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
XS_HD bool vnull(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
// Vec3::close (vec.hpp:120): float norm2 promoted to double and compared with the double 1e-4.
// The largest float below 1e-4 is 0x38d1b717, so `(double)n2 < 1e-4`  <=>  `n2 <= 0x38d1b717`.
XS_HD bool vclose(V3 a, V3 b) {
  const float lim = 9.99999974737875163555145263671875e-05f;  // 0x38d1b717
  return vnorm2(vsub(a, b)) <= lim;
}

}  // namespace xs

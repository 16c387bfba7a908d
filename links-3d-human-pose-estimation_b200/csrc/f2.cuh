// Packed f32x2 arithmetic.  sm_100 issues FFMA2 / FMUL2 / FADD2: two fp32 operations per lane and instruction, which halves
// the issue slots of arithmetic that is identical for two joints (geometry kernels: two joints of a lane; metric kernels: two
// joints of a pose).  tests/hostsim compiles the same code with plain float pairs.
#pragma once
#include "devdefs.cuh"

namespace links {

// ---- packed f32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue two fp32 operations per lane and instruction)
typedef float2 F2;
__device__ __forceinline__ F2 f2_make(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ F2 f2_splat(float a) { return make_float2(a, a); }
#ifndef LINKS_HOSTSIM
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { return __fadd2_rn(a, b); }
#else
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif
// (qx, qy, qz) = R p  /  R^T p for two joints at once; R2[i] holds R[i] in both halves
__device__ __forceinline__ void f2_matvec(const F2 (&R)[9], F2 px, F2 py, F2 pz, F2& qx, F2& qy, F2& qz) {
  qx = f2_fma(R[2], pz, f2_fma(R[1], py, f2_mul(R[0], px)));
  qy = f2_fma(R[5], pz, f2_fma(R[4], py, f2_mul(R[3], px)));
  qz = f2_fma(R[8], pz, f2_fma(R[7], py, f2_mul(R[6], px)));
}
__device__ __forceinline__ void f2_matTvec(const F2 (&R)[9], F2 px, F2 py, F2 pz, F2& qx, F2& qy, F2& qz) {
  qx = f2_fma(R[6], pz, f2_fma(R[3], py, f2_mul(R[0], px)));
  qy = f2_fma(R[7], pz, f2_fma(R[4], py, f2_mul(R[1], px)));
  qz = f2_fma(R[8], pz, f2_fma(R[5], py, f2_mul(R[2], px)));
}


}  // namespace links

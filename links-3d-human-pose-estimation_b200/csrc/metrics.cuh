// Batched pose metrics: MPJPE / per-joint distances, threshold counts (PCK, AUC, CPS) and PA-MPJPE with an
// in-register 3x3 polar decomposition (scaled Newton iteration; one-sided Jacobi SVD as the fallback).  Replaces reference utils/metrics_batch.py:8-159 and the per-pose
// numpy loop around utils/metrics.py:35-171 (eval_h36m.py:83-93).
//
// Layout: a block of 64 threads owns 64 consecutive poses.  Their rows are one contiguous chunk of HBM per tensor that
// the bulk-copy engine drops into shared memory (stage.cuh; the next chunk is in flight while this one is scored);
// each thread then pulls its own pose into registers ONCE (row stride odd -> conflict free; even strides are read in
// a per-thread rotated joint order, which every formula here is invariant to) and does all math there.
#pragma once
#include "devdefs.cuh"
#include "stage.cuh"
#include "f2.cuh"

namespace links {

constexpr int kPosesPerBlock = 64;
constexpr int kMaxRow = 51;           // 3 * 17
constexpr int kRowStride = 52;        // static shared arrays hold kPosesPerBlock rows of up to 51 floats (+ slack)
constexpr int kChunkFloats = kPosesPerBlock * kMaxRow;      // one staged tensor chunk (13 056 B, a multiple of 128)
constexpr size_t metric_smem_bytes(int stages) { return static_cast<size_t>(2) * stages * kChunkFloats * sizeof(float); }

__device__ __forceinline__ double block_sum_double(double v, double* sh /*[2]*/) {
  // 64 threads = 2 warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LINKS_FULL_MASK, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return sh[0] + sh[1];
}

// A pose held in registers: c[a][k] = coordinate a of joint slot k.  JT = compile-time joint count (full unroll;
// JT = 0: runtime count <= 17, arrays sized for 17).  Slot k holds joint (k + rot) mod J.
template <int JT>
struct PoseRegs {
  static constexpr int kCap = JT ? JT : 17;
  float c[3][kCap];
};

template <int JT>
__device__ __forceinline__ int slot_joint(int k, int rot, int J) {
  int j = k + rot;
  if (JT == 16) return j & 15;
  return j >= J ? j - J : j;
}

// rot = 0 when the shared row stride (3J floats) is odd; otherwise the thread index (mod J) so that the 32 lanes of
// a warp hit 32 different banks.
template <int JT>
__device__ __forceinline__ int lane_rotation(int J) {
  if (JT != 0 && ((3 * JT) & 1)) return 0;
  if ((3 * J) & 1) return 0;
  return JT == 16 ? (threadIdx.x & 15) : static_cast<int>(threadIdx.x % static_cast<unsigned>(J));
}

template <int JT>
__device__ __forceinline__ void load_pose(const float* row, int J, int rot, PoseRegs<JT>& P) {
#pragma unroll
  for (int k = 0; k < PoseRegs<JT>::kCap; ++k) {
    if (JT == 0 && k >= J) break;
    const int j = rot ? slot_joint<JT>(k, rot, J) : k;
#pragma unroll
    for (int a = 0; a < 3; ++a) P.c[a][k] = row[a * J + j];
  }
}

// ---- MPJPE (metrics_batch.py:8-24): mean joint distance after root-centring (+ optional norm matching) ----------
// Joints (2p, 2p + 1) of a pose as one packed value; compile-time joint counts only (JT = 17 / 16).
template <int JT>
__device__ __forceinline__ F2 pose_pair(const PoseRegs<JT>& P, int a, int p) { return f2_make(P.c[a][2 * p], P.c[a][2 * p + 1]); }

template <int JT>
__device__ __forceinline__ float mpjpe_regs(const PoseRegs<JT>& R, const PoseRegs<JT>& P, const float (&r0)[3],
                                            const float (&p0)[3], int J, int rot, int use_scaling, float* dist_out,
                                            float* max_out) {
  if constexpr (JT >= 2) {
    // joint pairs in packed f32x2 arithmetic (half the issue slots); an odd last joint runs in scalar code
    constexpr int NP = JT / 2;
    const F2 m1 = f2_splat(-1.f);
    F2 np0[3], r02[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { np0[a] = f2_splat(-p0[a]); r02[a] = f2_splat(r0[a]); }
    float scale = 1.f;
    if (use_scaling) {
      F2 sp2 = f2_splat(0.f), sr2 = f2_splat(0.f);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const F2 x = f2_add(pose_pair<JT>(P, a, p), np0[a]);
          const F2 ny = f2_fma(pose_pair<JT>(R, a, p), m1, r02[a]);
          sp2 = f2_fma(x, x, sp2);
          sr2 = f2_fma(ny, ny, sr2);
        }
      }
      float sp = sp2.x + sp2.y, sr = sr2.x + sr2.y;
      if (JT & 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float x = P.c[a][JT - 1] - p0[a], y = R.c[a][JT - 1] - r0[a];
          sp += x * x;
          sr += y * y;
        }
      }
      scale = sqrtf(sr) / sqrtf(sp);
    }
    const F2 sc2 = f2_splat(scale);
    float acc = 0.f, mx = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      F2 d2 = f2_splat(0.f);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const F2 x = f2_add(pose_pair<JT>(P, a, p), np0[a]);
        const F2 ny = f2_fma(pose_pair<JT>(R, a, p), m1, r02[a]);      // -(R - r0)
        const F2 e = f2_fma(x, sc2, ny);
        d2 = f2_fma(e, e, d2);
      }
      const float da = sqrtf(d2.x), db = sqrtf(d2.y);
      if (dist_out) {
        dist_out[rot ? slot_joint<JT>(2 * p, rot, J) : 2 * p] = da;
        dist_out[rot ? slot_joint<JT>(2 * p + 1, rot, J) : 2 * p + 1] = db;
      }
      acc += da + db;
      mx = fmaxf(mx, fmaxf(da, db));
    }
    if (JT & 1) {
      float d2 = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float e = (P.c[a][JT - 1] - p0[a]) * scale - (R.c[a][JT - 1] - r0[a]);
        d2 += e * e;
      }
      const float d = sqrtf(d2);
      if (dist_out) dist_out[rot ? slot_joint<JT>(JT - 1, rot, J) : JT - 1] = d;
      acc += d;
      mx = fmaxf(mx, d);
    }
    if (max_out) *max_out = mx;
    return acc / static_cast<float>(J);
  } else {
  float scale = 1.f;
  if (use_scaling) {
    float sp = 0.f, sr = 0.f;
#pragma unroll
    for (int k = 0; k < PoseRegs<JT>::kCap; ++k) {
      if (JT == 0 && k >= J) break;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = P.c[a][k] - p0[a], y = R.c[a][k] - r0[a];
        sp += x * x;
        sr += y * y;
      }
    }
    scale = sqrtf(sr) / sqrtf(sp);
  }
  float acc = 0.f, mx = 0.f;
#pragma unroll
  for (int k = 0; k < PoseRegs<JT>::kCap; ++k) {
    if (JT == 0 && k >= J) break;
    float d2 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float e = (P.c[a][k] - p0[a]) * scale - (R.c[a][k] - r0[a]);
      d2 += e * e;
    }
    const float d = sqrtf(d2);
    if (dist_out) dist_out[rot ? slot_joint<JT>(k, rot, J) : k] = d;
    acc += d;
    mx = fmaxf(mx, d);
  }
  if (max_out) *max_out = mx;
  return acc / static_cast<float>(J);
  }
}

// ---- 3x3 one-sided Jacobi SVD -> polar factor Q = U V^T and sum of singular values ----------------------
// Returns false when columns p, q are already orthogonal to working precision (no rotation applied).
__device__ __forceinline__ bool jacobi_rotate(float (&B)[3][3], float (&V)[3][3], int p, int q) {
  const float alpha = B[0][p] * B[0][p] + B[1][p] * B[1][p] + B[2][p] * B[2][p];
  const float beta = B[0][q] * B[0][q] + B[1][q] * B[1][q] + B[2][q] * B[2][q];
  const float gamma = B[0][p] * B[0][q] + B[1][p] * B[1][q] + B[2][p] * B[2][q];
  if (fabsf(gamma) <= 1e-30f || gamma * gamma <= 1e-13f * (alpha * beta)) return false;
  const float zeta = (beta - alpha) / (2.f * gamma);
  const float t = (zeta >= 0.f ? 1.f : -1.f) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
  const float c = rsqrtf(1.f + t * t);
  const float s = c * t;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float bp = B[i][p], bq = B[i][q];
    B[i][p] = c * bp - s * bq;
    B[i][q] = s * bp + c * bq;
    const float vp = V[i][p], vq = V[i][q];
    V[i][p] = c * vp - s * vq;
    V[i][q] = s * vp + c * vq;
  }
  return true;
}

__device__ __forceinline__ void polar_svd3(const float (&A)[3][3], float (&Q)[3][3], float* sum_sigma) {
  float B[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { B[i][j] = A[i][j]; V[i][j] = (i == j) ? 1.f : 0.f; }
#pragma unroll 1
  for (int sweep = 0; sweep < 8; ++sweep) {   // quadratic convergence: 3-5 sweeps in practice
    bool any = jacobi_rotate(B, V, 0, 1);
    any |= jacobi_rotate(B, V, 0, 2);
    any |= jacobi_rotate(B, V, 1, 2);
    if (!any) break;
  }
  float sig[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) sig[k] = sqrtf(B[0][k] * B[0][k] + B[1][k] * B[1][k] + B[2][k] * B[2][k]);
  const float smax = fmaxf(sig[0], fmaxf(sig[1], sig[2]));
  float U[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float inv = sig[k] > 1e-12f * smax && sig[k] > 0.f ? 1.f / sig[k] : 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][k] = B[i][k] * inv;
  }
  // rank-deficient input (planar / collinear pose): complete the missing left vector with a cross product
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!(sig[k] > 1e-12f * smax && sig[k] > 0.f)) {
      const int a = (k + 1) % 3, b = (k + 2) % 3;
      U[0][k] = U[1][a] * U[2][b] - U[2][a] * U[1][b];
      U[1][k] = U[2][a] * U[0][b] - U[0][a] * U[2][b];
      U[2][k] = U[0][a] * U[1][b] - U[1][a] * U[0][b];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Q[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
  *sum_sigma = sig[0] + sig[1] + sig[2];
}

// sqrt(x) for x >= 0 through the reciprocal-square-root unit (<= 2 ulp; 0 -> 0)
__device__ __forceinline__ float approx_sqrt(float x) { return x > 0.f ? x * rsqrtf(x) : 0.f; }

// ---- polar factor by the scaled Newton iteration  Q <- (g Q + Q^-T / g) / 2,  g = (|Q^-T|_F / |Q|_F)^(1/2) --------
// Keeps the singular vectors and drives every singular value to 1, so it converges to the same U V^T as the SVD
// (det = -1 for mirrored inputs) in 5-6 steps of ~60 flops for pose covariances: ~4x fewer instructions and a far
// shorter dependent chain than the Jacobi sweeps.  Returns false (caller falls back to the SVD) for numerically
// rank-deficient input (planar / collinear poses) or if it has not converged after 10 steps.
__device__ __forceinline__ bool polar_newton3(const float (&A)[3][3], float (&Q)[3][3], float* sum_sigma) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Q[i][j] = A[i][j];
  bool ok = false;
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    float C[3][3];   // cofactors: Q^-T = C / det
    C[0][0] = Q[1][1] * Q[2][2] - Q[1][2] * Q[2][1]; C[0][1] = Q[1][2] * Q[2][0] - Q[1][0] * Q[2][2]; C[0][2] = Q[1][0] * Q[2][1] - Q[1][1] * Q[2][0];
    C[1][0] = Q[0][2] * Q[2][1] - Q[0][1] * Q[2][2]; C[1][1] = Q[0][0] * Q[2][2] - Q[0][2] * Q[2][0]; C[1][2] = Q[0][1] * Q[2][0] - Q[0][0] * Q[2][1];
    C[2][0] = Q[0][1] * Q[1][2] - Q[0][2] * Q[1][1]; C[2][1] = Q[0][2] * Q[1][0] - Q[0][0] * Q[1][2]; C[2][2] = Q[0][0] * Q[1][1] - Q[0][1] * Q[1][0];
    const float det = Q[0][0] * C[0][0] + Q[0][1] * C[0][1] + Q[0][2] * C[0][2];
    float nq = 0.f, nc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) { nq += Q[i][j] * Q[i][j]; nc += C[i][j] * C[i][j]; }
    // |det| = s1 s2 s3 and nq >= s1^2: a relative rank test on the smallest singular values
    if (!(det * det > 1e-14f * nq * nq * nq)) return false;
    // g only accelerates convergence (any g > 0 has the same fixed point) and a 2-ulp error in b perturbs the converged
    // factor by ~1e-7: the hardware approximations are exact enough and ~4x shorter than IEEE sqrt / divide
    const float g = approx_sqrt(approx_sqrt(__fdividef(nc, nq)) * __fdividef(1.f, fabsf(det)));
    const float a = 0.5f * g, b = __fdividef(0.5f, g * det);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float qn = a * Q[i][j] + b * C[i][j];
        d += (qn - Q[i][j]) * (qn - Q[i][j]);
        Q[i][j] = qn;
      }
    if (d < 3e-13f) { ok = true; break; }
  }
  float tr = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) tr += Q[i][j] * A[i][j];     // trace(Q^T A) = sum of singular values
  *sum_sigma = tr;
  return ok;
}

// ---- PA-MPJPE.  ONE covariance + ONE SVD serve both semantics: the polar factor Q = U V^T is invariant to the
// positive scale that distinguishes the two normalisations, and the trace scales linearly with it.
//   mode 0 (e_batch): metrics_batch.py:104-159 -- unit-RMS normalisation, R = diag(1,1,det) (U V^T), RMS match
//   mode 1 (e_best) : metrics.py:35-171 'best'  -- unit Frobenius norm, reflection allowed, optimal scale
// Split in two phases so that the 2 x 3J pose registers are dead while the SVD runs (the kernels re-read the pose
// from shared memory for phase 2).
struct PaFit {
  float mr[3], mp[3];   // means over joints
  float Q[3][3];        // polar factor
  float det;            // det(Q)
  float g1, g0;         // gains of mode 1 / mode 0
};

template <int JT>
__device__ __forceinline__ void pa_fit(const PoseRegs<JT>& R, const PoseRegs<JT>& P, int J, PaFit& f) {
  const float invJ = 1.f / static_cast<float>(J);
  float ssr = 0.f, ssp = 0.f;
  float A[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  if constexpr (JT >= 2) {
    // joint pairs in packed f32x2 arithmetic; an odd last joint runs in scalar code
    constexpr int NP = JT / 2;
    F2 smr[3], smp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { smr[a] = f2_splat(0.f); smp[a] = f2_splat(0.f); }
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int a = 0; a < 3; ++a) { smr[a] = f2_add(smr[a], pose_pair<JT>(R, a, p)); smp[a] = f2_add(smp[a], pose_pair<JT>(P, a, p)); }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      f.mr[a] = smr[a].x + smr[a].y;
      f.mp[a] = smp[a].x + smp[a].y;
      if (JT & 1) { f.mr[a] += R.c[a][JT - 1]; f.mp[a] += P.c[a][JT - 1]; }
      f.mr[a] *= invJ; f.mp[a] *= invJ;
    }
    F2 nmr[3], nmp[3], A2[3][3];
    F2 ssr2 = f2_splat(0.f), ssp2 = f2_splat(0.f);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      nmr[a] = f2_splat(-f.mr[a]); nmp[a] = f2_splat(-f.mp[a]);
#pragma unroll
      for (int b = 0; b < 3; ++b) A2[a][b] = f2_splat(0.f);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      F2 x[3], y[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) { x[a] = f2_add(pose_pair<JT>(R, a, p), nmr[a]); y[a] = f2_add(pose_pair<JT>(P, a, p), nmp[a]); }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ssr2 = f2_fma(x[a], x[a], ssr2);
        ssp2 = f2_fma(y[a], y[a], ssp2);
#pragma unroll
        for (int b = 0; b < 3; ++b) A2[a][b] = f2_fma(x[a], y[b], A2[a][b]);
      }
    }
    ssr = ssr2.x + ssr2.y; ssp = ssp2.x + ssp2.y;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) A[a][b] = A2[a][b].x + A2[a][b].y;
    if (JT & 1) {
      float x[3], y[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) { x[a] = R.c[a][JT - 1] - f.mr[a]; y[a] = P.c[a][JT - 1] - f.mp[a]; }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ssr += x[a] * x[a];
        ssp += y[a] * y[a];
#pragma unroll
        for (int b = 0; b < 3; ++b) A[a][b] += x[a] * y[b];
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) { f.mr[a] = 0.f; f.mp[a] = 0.f; }
#pragma unroll
    for (int k = 0; k < PoseRegs<JT>::kCap; ++k) {
      if (JT == 0 && k >= J) break;
#pragma unroll
      for (int a = 0; a < 3; ++a) { f.mr[a] += R.c[a][k]; f.mp[a] += P.c[a][k]; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { f.mr[a] *= invJ; f.mp[a] *= invJ; }
#pragma unroll
    for (int k = 0; k < PoseRegs<JT>::kCap; ++k) {
      if (JT == 0 && k >= J) break;
      float x[3], y[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) { x[a] = R.c[a][k] - f.mr[a]; y[a] = P.c[a][k] - f.mp[a]; }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ssr += x[a] * x[a];
        ssp += y[a] * y[a];
#pragma unroll
        for (int b = 0; b < 3; ++b) A[a][b] += x[a] * y[b];
      }
    }
  }
  const float nr1 = sqrtf(ssr), np1 = sqrtf(ssp);                           // unit Frobenius norm (mode 1)
  const float nr0 = sqrtf(ssr / (3.f * J)), np0 = sqrtf(ssp / (3.f * J));   // unit RMS (mode 0)
  const float inv = 1.f / (nr1 * np1);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) A[a][b] *= inv;
  float tr;
  if (!polar_newton3(A, f.Q, &tr)) polar_svd3(A, f.Q, &tr);
  f.det = f.Q[0][0] * (f.Q[1][1] * f.Q[2][2] - f.Q[1][2] * f.Q[2][1]) -
          f.Q[0][1] * (f.Q[1][0] * f.Q[2][2] - f.Q[1][2] * f.Q[2][0]) +
          f.Q[0][2] * (f.Q[1][0] * f.Q[2][1] - f.Q[1][1] * f.Q[2][0]);
  f.g1 = tr * nr1 / np1;      // mode 1: Z = normX * trace * (Y0 / normY) T
  f.g0 = nr0 / np0;           // mode 0: RMS match, last ROW of R scaled by det (metrics_batch.py:145-147)
}

// WANT: bit 0 -> e_batch (mode 0), bit 1 -> e_best (mode 1).  aligned (optional, natural joint order) receives the
// aligned pose of the single requested mode.
template <int JT, int WANT>
__device__ __forceinline__ void pa_errors(const PoseRegs<JT>& R, const PoseRegs<JT>& P, int J, int rot, const PaFit& f,
                                          float& e_best, float& e_batch, float* aligned) {
  float acc1 = 0.f, acc0 = 0.f;
  constexpr int kFirstScalar = (JT >= 2) ? (JT / 2) * 2 : 0;       // joints below it run pairwise in packed f32x2 arithmetic
  if constexpr (JT >= 2) {
    constexpr int NP = JT / 2;
    F2 Q2[3][3], nmr[3], nmp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      nmr[a] = f2_splat(-f.mr[a]); nmp[a] = f2_splat(-f.mp[a]);
#pragma unroll
      for (int b = 0; b < 3; ++b) Q2[a][b] = f2_splat(f.Q[a][b]);
    }
    const F2 ng1 = f2_splat(-f.g1), ng0 = f2_splat(-f.g0), ng0d = f2_splat(-f.g0 * f.det);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      F2 y[3], x[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) { y[a] = f2_add(pose_pair<JT>(P, a, p), nmp[a]); x[a] = f2_add(pose_pair<JT>(R, a, p), nmr[a]); }
      F2 d1 = f2_splat(0.f), d0 = f2_splat(0.f);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const F2 q = f2_fma(Q2[a][2], y[2], f2_fma(Q2[a][1], y[1], f2_mul(Q2[a][0], y[0])));
        if (WANT & 2) {
          const F2 e1 = f2_fma(q, ng1, x[a]);                      // x - g1 q
          if (WANT == 2 && aligned) {
            aligned[a * J + (rot ? slot_joint<JT>(2 * p, rot, J) : 2 * p)] = f.g1 * q.x + f.mr[a];
            aligned[a * J + (rot ? slot_joint<JT>(2 * p + 1, rot, J) : 2 * p + 1)] = f.g1 * q.y + f.mr[a];
          }
          d1 = f2_fma(e1, e1, d1);
        }
        if (WANT & 1) {
          const F2 e0 = f2_fma(q, a == 2 ? ng0d : ng0, x[a]);      // last ROW of R scaled by det (metrics_batch.py:145-147)
          if (WANT == 1 && aligned) {
            const float gz = a == 2 ? f.g0 * f.det : f.g0;
            aligned[a * J + (rot ? slot_joint<JT>(2 * p, rot, J) : 2 * p)] = gz * q.x + f.mr[a];
            aligned[a * J + (rot ? slot_joint<JT>(2 * p + 1, rot, J) : 2 * p + 1)] = gz * q.y + f.mr[a];
          }
          d0 = f2_fma(e0, e0, d0);
        }
      }
      if (WANT & 2) acc1 += approx_sqrt(d1.x) + approx_sqrt(d1.y);
      if (WANT & 1) acc0 += approx_sqrt(d0.x) + approx_sqrt(d0.y);
    }
  }
#pragma unroll
  for (int k = kFirstScalar; k < PoseRegs<JT>::kCap; ++k) {
    if (JT == 0 && k >= J) break;
    float y[3], x[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { y[a] = P.c[a][k] - f.mp[a]; x[a] = R.c[a][k] - f.mr[a]; }
    float d1 = 0.f, d0 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float q = f.Q[a][0] * y[0] + f.Q[a][1] * y[1] + f.Q[a][2] * y[2];
      if (WANT & 2) {
        const float z1 = f.g1 * q;
        if (WANT == 2 && aligned) aligned[a * J + (rot ? slot_joint<JT>(k, rot, J) : k)] = z1 + f.mr[a];
        const float e1 = x[a] - z1;
        d1 += e1 * e1;
      }
      if (WANT & 1) {
        const float z0 = f.g0 * (a == 2 ? f.det * q : q);
        if (WANT == 1 && aligned) aligned[a * J + (rot ? slot_joint<JT>(k, rot, J) : k)] = z0 + f.mr[a];
        const float e0 = x[a] - z0;
        d0 += e0 * e0;
      }
    }
    if (WANT & 2) acc1 += approx_sqrt(d1);
    if (WANT & 1) acc0 += approx_sqrt(d0);
  }
  const float invJ = 1.f / static_cast<float>(J);
  e_best = acc1 * invJ;
  e_batch = acc0 * invJ;
}

// Compiler barrier: forces the second read of a staged pose to be a real shared-memory load instead of keeping the
// first copy alive in registers across the SVD.
__device__ __forceinline__ void forget_registers() { asm volatile("" ::: "memory"); }

// =========================================================================================================
// All kernels walk 64-pose chunks with a grid-stride loop (grid = a few blocks per SM) and keep their partial sums in
// registers: one double atomic per block at the end instead of one per 64 poses (single-address atomics serialise in
// L2 and dominated the run time at 8 M poses).
template <int JT, int kStages>
__global__ void __launch_bounds__(kPosesPerBlock) mpjpe_kernel(
    const float* __restrict__ p_ref, const float* __restrict__ p, int M, int Jr, int root, int use_scaling,
    float* __restrict__ per_pose, float* __restrict__ per_pose_max, float* __restrict__ dist, double* sum) {
  LINKS_DYN_SMEM(float, s_dyn);                        // [2 tensors][kStages][kChunkFloats]
  float* const s_ref = s_dyn;
  float* const s_p = s_dyn + kStages * kChunkFloats;
  __shared__ __align__(8) uint64_t s_bar[kStages];
  __shared__ double s_red[2];
  const int J = JT ? JT : Jr;
  const int row_len = 3 * J;
  const int t = threadIdx.x;
  const int nchunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  const int rot = lane_rotation<JT>(J);
  const bool bulk_ok = ((kPosesPerBlock * row_len) & 3) == 0;
  double acc = 0.0;
  stage_bars_init(s_bar, kStages);
  chunk_pipeline<kStages>(
      nchunks, s_bar,
      [&](int chunk, int st) {
        const int pose0 = chunk * kPosesPerBlock;
        const int npos = min(kPosesPerBlock, M - pose0);
        const StageReq rq[2] = {
            {s_ref + st * kChunkFloats, p_ref + static_cast<size_t>(pose0) * row_len, static_cast<uint32_t>(npos * row_len)},
            {s_p + st * kChunkFloats, p + static_cast<size_t>(pose0) * row_len, static_cast<uint32_t>(npos * row_len)}};
        const bool bulk = bulk_ok && npos == kPosesPerBlock;
        stage_issue(rq, s_bar + st, bulk);
        return bulk;
      },
      [&](int chunk, int st) {
        const int pose0 = chunk * kPosesPerBlock;
        const int npos = min(kPosesPerBlock, M - pose0);
        if (t < npos) {
          const float* rr = s_ref + st * kChunkFloats + t * row_len;
          const float* pp = s_p + st * kChunkFloats + t * row_len;
          PoseRegs<JT> R, P;
          load_pose<JT>(rr, J, rot, R);
          load_pose<JT>(pp, J, rot, P);
          const float r0[3] = {rr[root], rr[J + root], rr[2 * J + root]};
          const float p0[3] = {pp[root], pp[J + root], pp[2 * J + root]};
          float mx;
          const float e = mpjpe_regs<JT>(R, P, r0, p0, J, rot, use_scaling,
                                         dist ? dist + static_cast<size_t>(pose0 + t) * J : nullptr, &mx);
          if (per_pose) per_pose[pose0 + t] = e;
          if (per_pose_max) per_pose_max[pose0 + t] = mx;
          acc += static_cast<double>(e);
        }
      });
  if (sum != nullptr) {   // uniform branch
    const double tot = block_sum_double(acc, s_red);
    if (t == 0) atomicAdd(sum, tot);
  }
}

template <int JT, int kStages>
__global__ void __launch_bounds__(kPosesPerBlock, 8) pmpjpe_kernel(
    const float* __restrict__ p_ref, const float* __restrict__ p, int M, int Jr, int mode,
    float* __restrict__ per_pose, float* __restrict__ aligned, double* sum) {
  LINKS_DYN_SMEM(float, s_dyn);                        // [2 tensors][kStages][kChunkFloats]
  float* const s_ref = s_dyn;
  float* const s_p = s_dyn + kStages * kChunkFloats;
  __shared__ __align__(8) uint64_t s_bar[kStages];
  __shared__ double s_red[2];
  const int J = JT ? JT : Jr;
  const int row_len = 3 * J;
  const int t = threadIdx.x;
  const int nchunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  const int rot = lane_rotation<JT>(J);
  const bool bulk_ok = ((kPosesPerBlock * row_len) & 3) == 0;
  double acc = 0.0;
  stage_bars_init(s_bar, kStages);
  chunk_pipeline<kStages>(
      nchunks, s_bar,
      [&](int chunk, int st) {
        const int pose0 = chunk * kPosesPerBlock;
        const int npos = min(kPosesPerBlock, M - pose0);
        const StageReq rq[2] = {
            {s_ref + st * kChunkFloats, p_ref + static_cast<size_t>(pose0) * row_len, static_cast<uint32_t>(npos * row_len)},
            {s_p + st * kChunkFloats, p + static_cast<size_t>(pose0) * row_len, static_cast<uint32_t>(npos * row_len)}};
        const bool bulk = bulk_ok && npos == kPosesPerBlock;
        stage_issue(rq, s_bar + st, bulk);
        return bulk;
      },
      [&](int chunk, int st) {
        const int pose0 = chunk * kPosesPerBlock;
        const int npos = min(kPosesPerBlock, M - pose0);
        if (t < npos) {
          const float* rr = s_ref + st * kChunkFloats + t * row_len;
          const float* pp = s_p + st * kChunkFloats + t * row_len;
          PaFit fit;
          {
            PoseRegs<JT> R, P;
            load_pose<JT>(rr, J, rot, R);
            load_pose<JT>(pp, J, rot, P);
            pa_fit<JT>(R, P, J, fit);
          }
          forget_registers();
          float eb, e0;
          {
            PoseRegs<JT> R, P;
            load_pose<JT>(rr, J, rot, R);
            load_pose<JT>(pp, J, rot, P);
            float* al = aligned ? aligned + static_cast<size_t>(pose0 + t) * row_len : nullptr;
            if (mode == 1) pa_errors<JT, 2>(R, P, J, rot, fit, eb, e0, al);     // block-uniform
            else pa_errors<JT, 1>(R, P, J, rot, fit, eb, e0, al);
          }
          const float e = mode == 1 ? eb : e0;
          if (per_pose) per_pose[pose0 + t] = e;
          acc += static_cast<double>(e);
        }
      });
  if (sum != nullptr) {
    const double tot = block_sum_double(acc, s_red);
    if (t == 0) atomicAdd(sum, tot);
  }
}

// counts[k] += #(values < thr[k]) (strict) or #(values <= thr[k]); thresholds ascending, T <= 512.
// metrics_batch.py:40 (PCK), :60-62 (AUC), :86-95 (get_all AUC / CPS).
__global__ void __launch_bounds__(256) threshold_counts_kernel(const float* __restrict__ values, size_t n,
                                                               const float* __restrict__ thr, int T, int strict,
                                                               unsigned long long* counts) {
  __shared__ unsigned int diff[513];
  __shared__ float sthr[512];
  for (int i = threadIdx.x; i <= T; i += blockDim.x) diff[i] = 0u;
  for (int i = threadIdx.x; i < T; i += blockDim.x) sthr[i] = thr[i];
  __syncthreads();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = values[i];
    // first k with (v < thr[k]) / (v <= thr[k]); NaN never counts
    int lo = 0, hi = T;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const bool ok = strict ? (v < sthr[mid]) : (v <= sthr[mid]);
      if (ok) hi = mid; else lo = mid + 1;
    }
    if (v == v) atomicAdd(&diff[lo], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int k = 0; k < T; ++k) {
      run += diff[k];
      if (run) atomicAdd(&counts[k], run);
    }
  }
}

// Eval fusion (eval_h36m.py:58-97): pred = [x*d, y*d, d] with d = depth_off + depth (no clamp, not centred);
// sums3 += (sum N-MPJPE(root 0, scaled), sum PA-MPJPE 'best', sum PA-MPJPE batch).  J = 17.
// The depth rows (17 of ld_depth floats used) are bulk-copied whole and compacted IN PLACE to stride 17 in shared memory
// (30 KB per block: 7 blocks per SM).
constexpr int kEvalMaxLd = 32;
__global__ void __launch_bounds__(kPosesPerBlock, 7) eval_lift_score_kernel(
    const float* __restrict__ poses_2d, const float* __restrict__ depth_off, int ld_depth,
    const float* __restrict__ gt, int M, float depth, double* sums3) {
  __shared__ __align__(128) float s_ref[kPosesPerBlock * kRowStride];
  __shared__ __align__(128) float s_2d[kPosesPerBlock * 34];
  __shared__ __align__(128) float s_draw[kPosesPerBlock * kEvalMaxLd];   // raw depth rows, compacted in place to stride 17
  __shared__ __align__(8) uint64_t s_bar[1];
  __shared__ double s_red[2];
  constexpr int J = 17;
  const int t = threadIdx.x;
  const int nchunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  const bool depth_bulk = ld_depth <= kEvalMaxLd && (ld_depth & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(depth_off) & 15u) == 0;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  stage_bars_init(s_bar, 1);
  chunk_pipeline<1>(
      nchunks, s_bar,
      [&](int chunk, int st) {
        (void)st;
        const int pose0 = chunk * kPosesPerBlock;
        const int npos = min(kPosesPerBlock, M - pose0);
        const bool bulk = npos == kPosesPerBlock && depth_bulk;
        if (bulk) {
          const StageReq rq[3] = {
              {s_ref, gt + static_cast<size_t>(pose0) * 51, static_cast<uint32_t>(npos * 51)},
              {s_2d, poses_2d + static_cast<size_t>(pose0) * 34, static_cast<uint32_t>(npos * 34)},
              {s_draw, depth_off + static_cast<size_t>(pose0) * ld_depth, static_cast<uint32_t>(npos * ld_depth)}};
          stage_issue(rq, s_bar, true);
        } else {
          const StageReq rq[2] = {
              {s_ref, gt + static_cast<size_t>(pose0) * 51, static_cast<uint32_t>(npos * 51)},
              {s_2d, poses_2d + static_cast<size_t>(pose0) * 34, static_cast<uint32_t>(npos * 34)}};
          stage_issue(rq, s_bar, false);
          for (int i = threadIdx.x; i < npos * J; i += blockDim.x) {
            const int r = i / J, j = i - r * J;
            s_draw[i] = depth_off[static_cast<size_t>(pose0 + r) * ld_depth + j];
          }
        }
        return bulk;
      },
      [&](int chunk, int st) {
        (void)st;
        const int pose0 = chunk * kPosesPerBlock;
        const int npos = min(kPosesPerBlock, M - pose0);
        if (npos == kPosesPerBlock && depth_bulk) {     // block-uniform
          // in place (the compact image overlaps the rows it is read from): through registers, between two barriers
          float tmp[J];
#pragma unroll
          for (int mm = 0; mm < J; ++mm) {
            const int i = threadIdx.x + kPosesPerBlock * mm;
            const int r = i / J, j = i - r * J;
            tmp[mm] = s_draw[r * ld_depth + j];
          }
          __syncthreads();
#pragma unroll
          for (int mm = 0; mm < J; ++mm) s_draw[threadIdx.x + kPosesPerBlock * mm] = tmp[mm];
          __syncthreads();
        }
        if (t < npos) {
          const float* q = s_2d + t * 34;
          const float* dd = s_draw + t * J;
          auto lift = [&](PoseRegs<17>& P) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
              const float d = dd[j] + depth;
              P.c[0][j] = q[j] * d;
              P.c[1][j] = q[J + j] * d;
              P.c[2][j] = d;
            }
          };
          PaFit fit;
          {
            PoseRegs<17> R, P;
            load_pose<17>(s_ref + t * 51, J, 0, R);
            lift(P);
            const float r0[3] = {R.c[0][0], R.c[1][0], R.c[2][0]};
            const float p0[3] = {P.c[0][0], P.c[1][0], P.c[2][0]};
            a0 += static_cast<double>(mpjpe_regs<17>(R, P, r0, p0, J, 0, 1, nullptr, nullptr));
          }
          forget_registers();
          {
            PoseRegs<17> R, P;
            load_pose<17>(s_ref + t * 51, J, 0, R);
            lift(P);
            pa_fit<17>(R, P, J, fit);
          }
          forget_registers();
          {
            PoseRegs<17> R, P;
            load_pose<17>(s_ref + t * 51, J, 0, R);
            lift(P);
            float eb, e0;
            pa_errors<17, 3>(R, P, J, 0, fit, eb, e0, nullptr);
            a1 += static_cast<double>(eb);
            a2 += static_cast<double>(e0);
          }
        }
      });
  const double t0 = block_sum_double(a0, s_red);
  __syncthreads();
  const double t1 = block_sum_double(a1, s_red);
  __syncthreads();
  const double t2 = block_sum_double(a2, s_red);
  if (t == 0) {
    atomicAdd(sums3 + 0, t0);
    atomicAdd(sums3 + 1, t1);
    atomicAdd(sums3 + 2, t2);
  }
}

}  // namespace links

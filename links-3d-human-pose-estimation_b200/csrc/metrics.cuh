// Batched pose metrics: MPJPE / per-joint distances, threshold counts (PCK, AUC, CPS) and PA-MPJPE with an
// in-register 3x3 one-sided Jacobi SVD.  Replaces reference utils/metrics_batch.py:8-159 and the per-pose
// numpy loop around utils/metrics.py:35-171 (eval_h36m.py:83-93).
//
// Layout: a block of 64 threads owns 64 consecutive poses.  Their rows are one contiguous chunk of HBM
// (64*3J floats per tensor) that is staged into shared memory with coalesced float4 loads; afterwards each
// thread works on its own pose out of shared memory (row stride odd -> conflict free) entirely in registers.
#pragma once
#include "devdefs.cuh"

namespace links {

constexpr int kPosesPerBlock = 64;
constexpr int kMaxRow = 51;           // 3 * 17
constexpr int kRowStride = 52;        // shared arrays hold kPosesPerBlock rows of up to 51 floats (+ float4 slack)

// Cooperative LINEAR copy of `count` floats starting at g (16-byte aligned chunk start) into shared memory: the
// shared rows keep the global row length as their stride (51 for 17 joints: odd, so lane = pose reads are bank-conflict
// free; 34 for the 2D poses: 2-way), which makes staging pure float4 traffic with no per-element index arithmetic.
__device__ __forceinline__ void stage_rows(const float* __restrict__ g, size_t count, float* s) {
  const size_t n4 = count >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* s4 = reinterpret_cast<float4*>(s);
  for (size_t i = threadIdx.x; i < n4; i += blockDim.x) s4[i] = g4[i];
  for (size_t ee = (n4 << 2) + threadIdx.x; ee < count; ee += blockDim.x) s[ee] = g[ee];
}

__device__ __forceinline__ double block_sum_double(double v, double* sh /*[2]*/) {
  // 64 threads = 2 warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LINKS_FULL_MASK, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return sh[0] + sh[1];
}

// ---- MPJPE (metrics_batch.py:8-24) for the pose in shared rows r (ref) and p (pred) --------------------
// Returns mean joint distance; optionally writes joint distances and the max.
__device__ __forceinline__ float mpjpe_row(const float* r, const float* p, int J, int root, int use_scaling,
                                           float* dist_out, float* max_out) {
  const float rx = r[root], ry = r[J + root], rz = r[2 * J + root];
  const float px = p[root], py = p[J + root], pz = p[2 * J + root];
  float scale = 1.f;
  if (use_scaling) {
    float sp = 0.f, sr = 0.f;
    for (int j = 0; j < J; ++j) {
      const float ax = p[j] - px, ay = p[J + j] - py, az = p[2 * J + j] - pz;
      const float bx = r[j] - rx, by = r[J + j] - ry, bz = r[2 * J + j] - rz;
      sp += ax * ax + ay * ay + az * az;
      sr += bx * bx + by * by + bz * bz;
    }
    scale = sqrtf(sr) / sqrtf(sp);
  }
  float acc = 0.f, mx = 0.f;
  for (int j = 0; j < J; ++j) {
    const float dx = (p[j] - px) * scale - (r[j] - rx);
    const float dy = (p[J + j] - py) * scale - (r[J + j] - ry);
    const float dz = (p[2 * J + j] - pz) * scale - (r[2 * J + j] - rz);
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    if (dist_out) dist_out[j] = d;
    acc += d;
    mx = fmaxf(mx, d);
  }
  if (max_out) *max_out = mx;
  return acc / static_cast<float>(J);
}

// ---- 3x3 one-sided Jacobi SVD -> polar factor Q = U V^T and sum of singular values ----------------------
__device__ __forceinline__ void jacobi_rotate(float (&B)[3][3], float (&V)[3][3], int p, int q) {
  const float alpha = B[0][p] * B[0][p] + B[1][p] * B[1][p] + B[2][p] * B[2][p];
  const float beta = B[0][q] * B[0][q] + B[1][q] * B[1][q] + B[2][q] * B[2][q];
  const float gamma = B[0][p] * B[0][q] + B[1][p] * B[1][q] + B[2][p] * B[2][q];
  if (fabsf(gamma) <= 1e-30f || fabsf(gamma) <= 1e-9f * sqrtf(alpha * beta)) return;
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
}

__device__ __forceinline__ void polar_svd3(const float (&A)[3][3], float (&Q)[3][3], float* sum_sigma) {
  float B[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { B[i][j] = A[i][j]; V[i][j] = (i == j) ? 1.f : 0.f; }
#pragma unroll 1
  for (int sweep = 0; sweep < 8; ++sweep) {
    jacobi_rotate(B, V, 0, 1);
    jacobi_rotate(B, V, 0, 2);
    jacobi_rotate(B, V, 1, 2);
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

// ---- PA-MPJPE for one pose.  mode 0: metrics_batch.py:104-159; mode 1: metrics.py:35-171 ('best') ---
__device__ __forceinline__ float pmpjpe_row(const float* r, const float* p, int J, int mode, float* aligned = nullptr) {
  const float invJ = 1.f / static_cast<float>(J);
  float mr[3] = {0.f, 0.f, 0.f}, mp[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < J; ++j) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { mr[a] += r[a * J + j]; mp[a] += p[a * J + j]; }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mr[a] *= invJ; mp[a] *= invJ; }
  float ssr = 0.f, ssp = 0.f;
  float A[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int j = 0; j < J; ++j) {
    float x[3], y[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { x[a] = r[a * J + j] - mr[a]; y[a] = p[a * J + j] - mp[a]; }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      ssr += x[a] * x[a];
      ssp += y[a] * y[a];
#pragma unroll
      for (int b = 0; b < 3; ++b) A[a][b] += x[a] * y[b];
    }
  }
  // normalise: mode 1 -> unit Frobenius norm; mode 0 -> unit RMS (the common factor cancels in Q)
  const float nr = mode == 1 ? sqrtf(ssr) : sqrtf(ssr / (3.f * J));
  const float np = mode == 1 ? sqrtf(ssp) : sqrtf(ssp / (3.f * J));
  const float inv = 1.f / (nr * np);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) A[a][b] *= inv;
  float Q[3][3], tr;
  polar_svd3(A, Q, &tr);
  float gain = 1.f;
  if (mode == 1) {
    gain = tr;   // optimal scale: Z = normX * trace * Y0 T + muX
  } else {
    // R = diag(1,1,det(UV^T)) @ (U V^T): scale the last ROW (metrics_batch.py:145-147)
    const float det = Q[0][0] * (Q[1][1] * Q[2][2] - Q[1][2] * Q[2][1]) - Q[0][1] * (Q[1][0] * Q[2][2] - Q[1][2] * Q[2][0]) +
                      Q[0][2] * (Q[1][0] * Q[2][1] - Q[1][1] * Q[2][0]);
#pragma unroll
    for (int b = 0; b < 3; ++b) Q[2][b] *= det;
  }
  float acc = 0.f;
  const float invp = 1.f / np;
  for (int j = 0; j < J; ++j) {
    float y[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) y[a] = (p[a * J + j] - mp[a]) * invp;
    float d2 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float z = gain * nr * (Q[a][0] * y[0] + Q[a][1] * y[1] + Q[a][2] * y[2]);
      const float e = (r[a * J + j] - mr[a]) - z;
      if (aligned) aligned[a * J + j] = z + mr[a];
      d2 += e * e;
    }
    acc += sqrtf(d2);
  }
  return acc * invJ;
}

// Both PA-MPJPE semantics of one pose from ONE covariance + ONE SVD: the polar factor Q = U V^T is invariant to the
// positive scale that distinguishes the two normalisations, and the trace scales linearly with it.
// Returns mode-1 ('best') error in e_best and mode-0 (metrics_batch) error in e_batch.
__device__ __forceinline__ void pmpjpe_row_both(const float* r, const float* p, int J, float& e_best, float& e_batch) {
  const float invJ = 1.f / static_cast<float>(J);
  float mr[3] = {0.f, 0.f, 0.f}, mp[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < J; ++j) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { mr[a] += r[a * J + j]; mp[a] += p[a * J + j]; }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mr[a] *= invJ; mp[a] *= invJ; }
  float ssr = 0.f, ssp = 0.f;
  float A[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int j = 0; j < J; ++j) {
    float x[3], y[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { x[a] = r[a * J + j] - mr[a]; y[a] = p[a * J + j] - mp[a]; }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      ssr += x[a] * x[a];
      ssp += y[a] * y[a];
#pragma unroll
      for (int b = 0; b < 3; ++b) A[a][b] += x[a] * y[b];
    }
  }
  const float nr1 = sqrtf(ssr), np1 = sqrtf(ssp);                       // unit Frobenius norm (mode 1)
  const float nr0 = sqrtf(ssr / (3.f * J)), np0 = sqrtf(ssp / (3.f * J));   // unit RMS (mode 0)
  const float inv = 1.f / (nr1 * np1);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) A[a][b] *= inv;
  float Q[3][3], tr;
  polar_svd3(A, Q, &tr);
  const float det = Q[0][0] * (Q[1][1] * Q[2][2] - Q[1][2] * Q[2][1]) - Q[0][1] * (Q[1][0] * Q[2][2] - Q[1][2] * Q[2][0]) +
                    Q[0][2] * (Q[1][0] * Q[2][1] - Q[1][1] * Q[2][0]);
  const float g1 = tr * nr1 / np1;      // mode 1: Z = normX * trace * (Y0 / normY) T
  const float g0 = nr0 / np0;           // mode 0: RMS match, last row of R scaled by det
  float acc1 = 0.f, acc0 = 0.f;
  for (int j = 0; j < J; ++j) {
    float y[3], x[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { y[a] = p[a * J + j] - mp[a]; x[a] = r[a * J + j] - mr[a]; }
    float d1 = 0.f, d0 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float q = Q[a][0] * y[0] + Q[a][1] * y[1] + Q[a][2] * y[2];
      const float e1 = x[a] - g1 * q;
      const float e0 = x[a] - g0 * (a == 2 ? det * q : q);
      d1 += e1 * e1;
      d0 += e0 * e0;
    }
    acc1 += sqrtf(d1);
    acc0 += sqrtf(d0);
  }
  e_best = acc1 * invJ;
  e_batch = acc0 * invJ;
}

// =========================================================================================================
// All three kernels walk 64-pose chunks with a grid-stride loop (grid = a few blocks per SM) and keep their partial
// sums in registers: one double atomic per block at the end instead of one per 64 poses (single-address atomics
// serialise in L2 and dominated the run time at 8 M poses).
__global__ void __launch_bounds__(kPosesPerBlock) mpjpe_kernel(
    const float* __restrict__ p_ref, const float* __restrict__ p, int M, int J, int root, int use_scaling,
    float* __restrict__ per_pose, float* __restrict__ per_pose_max, float* __restrict__ dist, double* sum) {
  __shared__ __align__(16) float s_ref[kPosesPerBlock * kRowStride];
  __shared__ __align__(16) float s_p[kPosesPerBlock * kRowStride];
  __shared__ double s_red[2];
  const int row_len = 3 * J;
  const int t = threadIdx.x;
  const int nchunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  double acc = 0.0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int pose0 = chunk * kPosesPerBlock;
    const int npos = min(kPosesPerBlock, M - pose0);
    stage_rows(p_ref + static_cast<size_t>(pose0) * row_len, static_cast<size_t>(npos) * row_len, s_ref);
    stage_rows(p + static_cast<size_t>(pose0) * row_len, static_cast<size_t>(npos) * row_len, s_p);
    __syncthreads();
    if (t < npos) {
      float mx;
      const float e = mpjpe_row(s_ref + t * row_len, s_p + t * row_len, J, root, use_scaling,
                                dist ? dist + static_cast<size_t>(pose0 + t) * J : nullptr, &mx);
      if (per_pose) per_pose[pose0 + t] = e;
      if (per_pose_max) per_pose_max[pose0 + t] = mx;
      acc += static_cast<double>(e);
    }
    __syncthreads();
  }
  if (sum != nullptr) {   // uniform branch
    const double tot = block_sum_double(acc, s_red);
    if (t == 0) atomicAdd(sum, tot);
  }
}

__global__ void __launch_bounds__(kPosesPerBlock) pmpjpe_kernel(
    const float* __restrict__ p_ref, const float* __restrict__ p, int M, int J, int mode,
    float* __restrict__ per_pose, float* __restrict__ aligned, double* sum) {
  __shared__ __align__(16) float s_ref[kPosesPerBlock * kRowStride];
  __shared__ __align__(16) float s_p[kPosesPerBlock * kRowStride];
  __shared__ double s_red[2];
  const int row_len = 3 * J;
  const int t = threadIdx.x;
  const int nchunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  double acc = 0.0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int pose0 = chunk * kPosesPerBlock;
    const int npos = min(kPosesPerBlock, M - pose0);
    stage_rows(p_ref + static_cast<size_t>(pose0) * row_len, static_cast<size_t>(npos) * row_len, s_ref);
    stage_rows(p + static_cast<size_t>(pose0) * row_len, static_cast<size_t>(npos) * row_len, s_p);
    __syncthreads();
    if (t < npos) {
      const float e = pmpjpe_row(s_ref + t * row_len, s_p + t * row_len, J, mode,
                                 aligned ? aligned + static_cast<size_t>(pose0 + t) * row_len : nullptr);
      if (per_pose) per_pose[pose0 + t] = e;
      acc += static_cast<double>(e);
    }
    __syncthreads();
  }
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
__global__ void __launch_bounds__(kPosesPerBlock) eval_lift_score_kernel(
    const float* __restrict__ poses_2d, const float* __restrict__ depth_off, int ld_depth,
    const float* __restrict__ gt, int M, float depth, double* sums3) {
  __shared__ __align__(16) float s_ref[kPosesPerBlock * kRowStride];
  __shared__ __align__(16) float s_p[kPosesPerBlock * kRowStride];     // per pose: x*d (17), y*d (17), d (17)
  __shared__ __align__(16) float s_2d[kPosesPerBlock * 34];
  __shared__ double s_red[2];
  const int J = 17;
  const int t = threadIdx.x;
  const int nchunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int pose0 = chunk * kPosesPerBlock;
    const int npos = min(kPosesPerBlock, M - pose0);
    stage_rows(gt + static_cast<size_t>(pose0) * 51, static_cast<size_t>(npos) * 51, s_ref);
    stage_rows(poses_2d + static_cast<size_t>(pose0) * 34, static_cast<size_t>(npos) * 34, s_2d);
    for (int i = threadIdx.x; i < npos * J; i += blockDim.x) {
      const int r = i / J, j = i - r * J;
      s_p[r * 51 + 34 + j] = depth_off[static_cast<size_t>(pose0 + r) * ld_depth + j] + depth;
    }
    __syncthreads();
    if (t < npos) {
      float* p = s_p + t * 51;
      const float* q = s_2d + t * 34;
      for (int j = 0; j < J; ++j) {
        const float d = p[34 + j];
        p[j] = q[j] * d;
        p[J + j] = q[J + j] * d;
      }
      const float* r = s_ref + t * 51;
      a0 += static_cast<double>(mpjpe_row(r, p, J, 0, 1, nullptr, nullptr));
      float eb, e0;
      pmpjpe_row_both(r, p, J, eb, e0);
      a1 += static_cast<double>(eb);
      a2 += static_cast<double>(e0);
    }
    __syncthreads();
  }
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

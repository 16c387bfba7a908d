// Geometry + self-supervised losses of the lifter training step, forward and hand-derived backward.
// Reference: train_leg_torso_lifter.py:153-272 (V = 1 pose variant) and train_left_right_lifter.py:150-423
// (V = 2 variants: 'left' / 'right' choice of combine_left_right_pred_1d, utils/helpers.py:40-53).
//
// Mapping: one warp per consecutive row pair (2k, 2k+1) -- the pairwise deformation loss (:250-254) couples
// exactly those rows; lane j < 17 owns joint j; per-row reductions are warp shuffles.  Nothing but the
// network outputs is read: P, R, Q, q are recomputed from (u, depth heads, angle heads, eps_x, u_y, stats).
#pragma once
#include "devdefs.cuh"

namespace links {

constexpr int kGeomWarps = 4;
constexpr int kJ = 17;

struct GeomArgs {
  LinksGeomMaps maps;
  const float* u;
  const float* head[2];
  const float* ang[2];
  const float* eps_x;
  const float* u_y;
  const float* stats;
  const float* head2[2];
  const float* dflow[2];
  const float* dlift[2];
  int N;
  float* loss_sums;
  __nv_bfloat16* g2[2];
  __nv_bfloat16* g2T[2];
  __nv_bfloat16* g1[2];
  __nv_bfloat16* g1T[2];
  int ldT, colT0;
  float* dgamma;
  float* da;
  float* red;
  float* qpart[2];
  float* qfull[2];
};

// The argument block (index maps, pointers) is copied into shared memory once per block: the maps are indexed by
// lane, and lane-divergent reads of the kernel-parameter constant bank serialise (one transaction per distinct
// address) -- together with per-block single-address atomics that made these kernels 10-50x slower than the HBM
// roofline at large N.  Blocks then walk rows with a grid-stride loop and issue ONE set of atomics at the end.
__device__ __forceinline__ void stage_args(GeomArgs* dst, const GeomArgs& src) {
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(&src);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(GeomArgs) / 4); i += blockDim.x) d32[i] = s32[i];
  __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LINKS_FULL_MASK, v, o);
  return v;
}

struct Vec3 { float x, y, z; };
__device__ __forceinline__ Vec3 shfl3(Vec3 v, int src) {
  Vec3 r;
  r.x = __shfl_sync(LINKS_FULL_MASK, v.x, src);
  r.y = __shfl_sync(LINKS_FULL_MASK, v.y, src);
  r.z = __shfl_sync(LINKS_FULL_MASK, v.z, src);
  return r;
}
__device__ __forceinline__ Vec3 mat_vec(const float (&R)[9], Vec3 p) {   // R p
  Vec3 q;
  q.x = R[0] * p.x + R[1] * p.y + R[2] * p.z;
  q.y = R[3] * p.x + R[4] * p.y + R[5] * p.z;
  q.z = R[6] * p.x + R[7] * p.y + R[8] * p.z;
  return q;
}
__device__ __forceinline__ Vec3 matT_vec(const float (&R)[9], Vec3 p) {  // R^T p
  Vec3 q;
  q.x = R[0] * p.x + R[3] * p.y + R[6] * p.z;
  q.y = R[1] * p.x + R[4] * p.y + R[7] * p.z;
  q.z = R[2] * p.x + R[5] * p.y + R[8] * p.z;
  return q;
}

// R = Rx(a) @ (Ry(b) @ Rx(g))   (utils/rotation_conversions.py:11-36; train_leg_torso_lifter.py:159-181)
__device__ __forceinline__ void make_rotation(float a, float b, float g, float (&R)[9]) {
  float sa, ca, sb, cb, sg, cg;
  sincosf(a, &sa, &ca);
  sincosf(b, &sb, &cb);
  sincosf(g, &sg, &cg);
  R[0] = cb;       R[1] = sb * sg;                 R[2] = sb * cg;
  R[3] = sa * sb;  R[4] = ca * cg - sa * cb * sg;  R[5] = -ca * sg - sa * cb * cg;
  R[6] = -ca * sb; R[7] = sa * cg + ca * cb * sg;  R[8] = -sa * sg + ca * cb * cg;
}

// Per-lane state of one (row, variant).
struct RowVar {
  float ux, uy;      // input 2D joint
  float mask;        // 1 where depth not clamped (:186)
  float d;           // clamped depth
  Vec3 P;            // root-centred lifted joint (:188-192)
  Vec3 Q;            // rotated joint (:195)
  float zq, qx, qy;  // projection (:198-199)
  // pass-2 side
  float mask2, d2;
  Vec3 P2;           // re-lifted joint (:235-238)
  Vec3 F;            // Q - P2
  float L3d;
  Vec3 S;            // R^T P2 (:242)
  float zs, rx, ry;
};

// bone table (utils/helpers.py:140-141) as lane-indexed lookups
__device__ __forceinline__ int bone_i(int b) {
  const int t[16] = {0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15};
  return t[b & 15];
}
__device__ __forceinline__ int bone_j(int b) {
  const int t[16] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  return t[b & 15];
}

__device__ __forceinline__ void row_forward(const GeomArgs& A, int v, int n, int lane, const float (&R)[9], RowVar& s) {
  const bool act = lane < kJ;
  const int j = act ? lane : 0;
  const float D = A.maps.depth;
  s.ux = act ? A.u[static_cast<size_t>(n) * 34 + j] : 0.f;
  s.uy = act ? A.u[static_cast<size_t>(n) * 34 + kJ + j] : 0.f;
  float delta = 0.f;
  if (act && j != 0) delta = A.head[A.maps.src_net[v][j]][static_cast<size_t>(n) * LINKS_HEAD_LD + A.maps.col[j]];
  float d = delta + D;
  s.mask = (d < 1.0f) ? 0.f : 1.f;
  d = (d < 1.0f) ? 1.0f : d;
  s.d = d;
  Vec3 T;
  T.x = s.ux * d; T.y = s.uy * d; T.z = d;
  const Vec3 T0 = shfl3(T, 0);
  s.P.x = act ? T.x - T0.x : 0.f;
  s.P.y = act ? T.y - T0.y : 0.f;
  s.P.z = act ? T.z - T0.z : 0.f;
  s.Q = mat_vec(R, s.P);
  s.zq = s.Q.z + D;
  s.qx = s.Q.x / s.zq;
  s.qy = s.Q.y / s.zq;
}

// pass-2 quantities and the per-row loss terms.  Returns (L3d, rep_rot, bl_prior) via out[3].
__device__ __forceinline__ void row_consistency(const GeomArgs& A, int v, int n, int lane, const float (&R)[9],
                                                RowVar& s, float (&out)[3], float& bl_mean, float& bl_len,
                                                float& bl_rho) {
  const bool act = lane < kJ;
  const int j = act ? lane : 0;
  const float D = A.maps.depth;
  float delta2 = 0.f;
  if (act && j != 0) delta2 = A.head2[A.maps.src_net[v][j]][static_cast<size_t>(n) * LINKS_HEAD_LD + A.maps.col[j]];
  float d2 = delta2 + D;
  s.mask2 = (d2 < 1.0f) ? 0.f : 1.f;
  d2 = (d2 < 1.0f) ? 1.0f : d2;
  s.d2 = d2;
  Vec3 T;
  T.x = s.qx * d2; T.y = s.qy * d2; T.z = d2;
  const Vec3 T0 = shfl3(T, 0);
  s.P2.x = act ? T.x - T0.x : 0.f;
  s.P2.y = act ? T.y - T0.y : 0.f;
  s.P2.z = act ? T.z - T0.z : 0.f;
  s.F.x = s.Q.x - s.P2.x; s.F.y = s.Q.y - s.P2.y; s.F.z = s.Q.z - s.P2.z;
  s.L3d = sqrtf(warp_sum(s.F.x * s.F.x + s.F.y * s.F.y + s.F.z * s.F.z));
  s.S = matT_vec(R, s.P2);
  s.zs = s.S.z + D;
  s.rx = s.S.x / s.zs;
  s.ry = s.S.y / s.zs;
  const float rep = warp_sum(act ? fabsf(s.rx - s.ux) + fabsf(s.ry - s.uy) : 0.f);
  // bone lengths: lane b < 16 owns bone b
  const bool bact = lane < 16;
  const Vec3 Pi = shfl3(s.P, bone_i(lane));
  const Vec3 Pj = shfl3(s.P, bone_j(lane));
  const float ex = Pi.x - Pj.x, ey = Pi.y - Pj.y, ez = Pi.z - Pj.z;
  bl_len = bact ? sqrtf(ex * ex + ey * ey + ez * ez) : 0.f;
  bl_mean = warp_sum(bl_len) * (1.f / 16.f);
  bl_rho = bl_len / bl_mean;
  const float c = A.maps.bone_rel[lane & 15];
  const float bl = warp_sum(bact ? (c - bl_rho) * (c - bl_rho) : 0.f);
  out[0] = s.L3d;
  out[1] = rep;
  out[2] = bl;
}

// =========================================================================================================
// forward: projected parts for the flows / pass-2 lifters
// =========================================================================================================
__global__ void __launch_bounds__(kGeomWarps * 32) geom_forward_kernel(const GeomArgs Ap) {
  __shared__ GeomArgs sA;
  stage_args(&sA, Ap);
  const GeomArgs& A = sA;
  const int lane = threadIdx.x & 31;
  // lane-constant map entries
  int part_net[2], part_idx[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    part_net[v] = lane < kJ ? A.maps.part_net[v][lane] : -1;
    part_idx[v] = lane < kJ ? A.maps.part_idx[v][lane] : 0;
  }
  const int stride = gridDim.x * kGeomWarps;
  for (int n = blockIdx.x * kGeomWarps + (threadIdx.x >> 5); n < A.N; n += stride) {   // one warp per row
    const float gamma = 0.5f * (A.ang[0][static_cast<size_t>(n) * LINKS_HEAD_LD] + A.ang[1][static_cast<size_t>(n) * LINKS_HEAD_LD]);
    const float a = -A.stats[0] + A.stats[1] * A.eps_x[n];
    const float b = (A.u_y[n] - 0.5f) * (1.99f * 3.14159265358979323846f);
    float R[9];
    make_rotation(a, b, gamma, R);
    for (int v = 0; v < A.maps.V; ++v) {
      RowVar s;
      row_forward(A, v, n, lane, R, s);
      if (lane < kJ) {
        if (A.qfull[v]) {
          A.qfull[v][static_cast<size_t>(n) * 34 + lane] = s.qx;
          A.qfull[v][static_cast<size_t>(n) * 34 + kJ + lane] = s.qy;
        }
        const int p = part_net[v];
        if (p >= 0) {
          const int nj = A.maps.n_joints[p];
          const int idx = part_idx[v];
          float* dst = A.qpart[p] + static_cast<size_t>(n) * (2 * nj);
          dst[idx] = s.qx;
          dst[nj + idx] = s.qy;
        }
      }
    }
  }
}

// =========================================================================================================
// losses + gradients.  kFull = false: loss sums and d/d(pass-2 heads) only (runs before the pass-2 backward);
// kFull = true : complete backward to the pass-1 heads, d gamma (direct) and d a (runs after it).
// =========================================================================================================
template <bool kFull>
__global__ void __launch_bounds__(kGeomWarps * 32) geom_lossgrad_kernel(const GeomArgs Ap) {
  __shared__ float s_part[kGeomWarps][6];
  __shared__ GeomArgs sA;
  stage_args(&sA, Ap);
  const GeomArgs& A = sA;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool act = lane < kJ;
  const float invN = 1.f / static_cast<float>(A.N);
  const int npairs = A.N / 2;
  const float c3d = A.maps.w_3d * invN, c2d = A.maps.w_2d * invN, cbl = A.maps.w_bl * invN;
  const float cv = npairs > 0 ? A.maps.w_vel / static_cast<float>(npairs) : 0.f;
  const float D = A.maps.depth;

  float sums[4] = {0.f, 0.f, 0.f, 0.f};   // L3d, rep, pair, bl (raw sums, lane-uniform)
  float red_da = 0.f, red_eda = 0.f;

  const int total_pairs = (A.N + 1) / 2;
  for (int pair = blockIdx.x * kGeomWarps + warp; pair < total_pairs; pair += gridDim.x * kGeomWarps) {
    const int nA = 2 * pair, nB = 2 * pair + 1;
    const bool vB = nB < A.N;                                    // warp-uniform
    float R[2][9], gam[2], eps[2];
    int rows[2] = {nA, nB};
    const int nrows = vB ? 2 : 1;
    #pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r >= nrows) break;
      const int n = rows[r];
      gam[r] = 0.5f * (A.ang[0][static_cast<size_t>(n) * LINKS_HEAD_LD] + A.ang[1][static_cast<size_t>(n) * LINKS_HEAD_LD]);
      eps[r] = A.eps_x[n];
      const float a = -A.stats[0] + A.stats[1] * eps[r];
      const float b = (A.u_y[n] - 0.5f) * (1.99f * 3.14159265358979323846f);
      make_rotation(a, b, gam[r], R[r]);
    }
    float dR[2][9];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < 9; ++k) dR[r][k] = 0.f;
    float g1acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [row][net] d/d(pass-1 head) for this lane's column
    float g2acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};

    for (int v = 0; v < A.maps.V; ++v) {
      RowVar s[2];
      float bl_mean[2], bl_len[2], bl_rho[2];
      #pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r >= nrows) break;
        float o[3];
        row_forward(A, v, rows[r], lane, R[r], s[r]);
        row_consistency(A, v, rows[r], lane, R[r], s[r], o, bl_mean[r], bl_len[r], bl_rho[r]);
        sums[0] += o[0]; sums[1] += o[1]; sums[3] += o[2];
      }
      // pairwise deformation (:250-254)
      Vec3 E; E.x = E.y = E.z = 0.f;
      float pnorm = 0.f;
      if (vB) {
        E.x = (s[0].P.x - s[1].P.x) - (s[0].S.x - s[1].S.x);
        E.y = (s[0].P.y - s[1].P.y) - (s[0].S.y - s[1].S.y);
        E.z = (s[0].P.z - s[1].P.z) - (s[0].S.z - s[1].S.z);
        pnorm = sqrtf(warp_sum(E.x * E.x + E.y * E.y + E.z * E.z));
        sums[2] += pnorm;
      }
      const float ge = pnorm > 0.f ? cv / pnorm : 0.f;
      #pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r >= nrows) break;
        RowVar& t = s[r];
        const float sgn = r == 0 ? 1.f : -1.f;
        // ---- d/dS: reprojection L1 (:247) + pair term
        const float drx = act ? c2d * ((t.rx > t.ux) ? 1.f : ((t.rx < t.ux) ? -1.f : 0.f)) : 0.f;
        const float dry = act ? c2d * ((t.ry > t.uy) ? 1.f : ((t.ry < t.uy) ? -1.f : 0.f)) : 0.f;
        Vec3 dS;
        dS.x = drx / t.zs - sgn * ge * E.x;
        dS.y = dry / t.zs - sgn * ge * E.y;
        dS.z = -(drx * t.rx + dry * t.ry) / t.zs - sgn * ge * E.z;
        // ---- d/dP2 = R dS - c3d F / L3d
        const float g3 = t.L3d > 0.f ? c3d / t.L3d : 0.f;
        Vec3 dP2 = mat_vec(R[r], dS);
        dP2.x -= g3 * t.F.x; dP2.y -= g3 * t.F.y; dP2.z -= g3 * t.F.z;
        if (!act) { dP2.x = dP2.y = dP2.z = 0.f; }
        // root centring: dT_j = dP2_j - [j==0] sum_i dP2_i
        const float sx = warp_sum(dP2.x), sy = warp_sum(dP2.y), sz = warp_sum(dP2.z);
        Vec3 dT2 = dP2;
        if (lane == 0) { dT2.x -= sx; dT2.y -= sy; dT2.z -= sz; }
        const float dd2 = dT2.x * t.qx + dT2.y * t.qy + dT2.z;
        const float ddelta2 = (act && lane != 0) ? t.mask2 * dd2 : 0.f;
        const int net = A.maps.src_net[v][act ? lane : 0];
        g2acc[r][0] += net == 0 ? ddelta2 : 0.f;
        g2acc[r][1] += net == 1 ? ddelta2 : 0.f;
        if (kFull) {
          // ---- d/dq: through P2 = (q d2) and the external consumers (flows, pass-2 lifters)
          float dqx = dT2.x * t.d2, dqy = dT2.y * t.d2;
          const int p = act ? A.maps.part_net[v][lane] : -1;
          if (p >= 0) {
            const int nj = A.maps.n_joints[p];
            const int idx = A.maps.part_idx[v][lane];
            const size_t n = rows[r];
            dqx += A.dflow[p][n * (2 * nj) + idx] + A.dlift[p][n * LINKS_HEAD_LD + idx];
            dqy += A.dflow[p][n * (2 * nj) + nj + idx] + A.dlift[p][n * LINKS_HEAD_LD + nj + idx];
          }
          // ---- d/dQ
          Vec3 dQ;
          dQ.x = g3 * t.F.x + dqx / t.zq;
          dQ.y = g3 * t.F.y + dqy / t.zq;
          dQ.z = g3 * t.F.z - (dqx * t.qx + dqy * t.qy) / t.zq;
          if (!act) { dQ.x = dQ.y = dQ.z = 0.f; }
          // ---- d/dP = R^T dQ + pair + bones
          Vec3 dP = matT_vec(R[r], dQ);
          dP.x += sgn * ge * E.x; dP.y += sgn * ge * E.y; dP.z += sgn * ge * E.z;
          {
            const bool bact = lane < 16;
            const float c = A.maps.bone_rel[lane & 15];
            const float h = bact ? -2.f * (c - bl_rho[r]) * cbl : 0.f;
            const float hl = warp_sum(h * bl_len[r]);
            const float dl = bact ? h / bl_mean[r] - hl / (16.f * bl_mean[r] * bl_mean[r]) : 0.f;
            const Vec3 Pi = shfl3(t.P, bone_i(lane));
            const Vec3 Pj = shfl3(t.P, bone_j(lane));
            const float w = (bact && bl_len[r] > 0.f) ? dl / bl_len[r] : 0.f;
            Vec3 dv;
            dv.x = w * (Pi.x - Pj.x); dv.y = w * (Pi.y - Pj.y); dv.z = w * (Pi.z - Pj.z);
#pragma unroll
            for (int bb = 0; bb < 16; ++bb) {
              const Vec3 x = shfl3(dv, bb);
              if (lane == bone_i(bb)) { dP.x += x.x; dP.y += x.y; dP.z += x.z; }
              if (lane == bone_j(bb)) { dP.x -= x.x; dP.y -= x.y; dP.z -= x.z; }
            }
          }
          if (!act) { dP.x = dP.y = dP.z = 0.f; }
          // ---- d/dR from Q = R P and S = R^T P2
          const float dq3[3] = {dQ.x, dQ.y, dQ.z}, p3[3] = {t.P.x, t.P.y, t.P.z};
          const float p23[3] = {t.P2.x, t.P2.y, t.P2.z}, ds3[3] = {dS.x, dS.y, dS.z};
#pragma unroll
          for (int ra = 0; ra < 3; ++ra)
#pragma unroll
            for (int cb = 0; cb < 3; ++cb)
              dR[r][ra * 3 + cb] += warp_sum(act ? dq3[ra] * p3[cb] + p23[ra] * ds3[cb] : 0.f);
          // ---- root centring + lift
          const float px = warp_sum(dP.x), py = warp_sum(dP.y), pz = warp_sum(dP.z);
          Vec3 dT = dP;
          if (lane == 0) { dT.x -= px; dT.y -= py; dT.z -= pz; }
          const float dd = dT.x * t.ux + dT.y * t.uy + dT.z;
          const float ddelta = (act && lane != 0) ? t.mask * dd : 0.f;
          g1acc[r][0] += net == 0 ? ddelta : 0.f;
          g1acc[r][1] += net == 1 ? ddelta : 0.f;
        }
      }
    }
    // ---- write head gradients (lane j -> column col[j] of every net that feeds joint j in some variant)
    #pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r >= nrows) break;
      const size_t n = rows[r];
      if (act) {
        const int c = A.maps.col[lane];
        bool feeds[2] = {false, false};
        for (int v = 0; v < A.maps.V; ++v) feeds[A.maps.src_net[v][lane]] = true;
#pragma unroll
        for (int net = 0; net < 2; ++net) {
          if (!feeds[net]) continue;
          if (!kFull) {
            const __nv_bfloat16 h = __float2bfloat16_rn(g2acc[r][net]);
            A.g2[net][n * 64 + c] = h;
            if (A.g2T[net]) A.g2T[net][static_cast<size_t>(c) * A.ldT + A.colT0 + n] = h;
          } else {
            const __nv_bfloat16 h = __float2bfloat16_rn(g1acc[r][net]);
            A.g1[net][n * 64 + c] = h;
            if (A.g1T[net]) A.g1T[net][static_cast<size_t>(c) * A.ldT + A.colT0 + n] = h;
          }
        }
      }
      if (kFull) {
        const float* Rr = R[r];
        const float* d = dR[r];
        // dR/da: row1' = -row2, row2' = row1 ; dR/dgamma: col1' = col2, col2' = -col1
        const float dav = -(d[3] * Rr[6] + d[4] * Rr[7] + d[5] * Rr[8]) + (d[6] * Rr[3] + d[7] * Rr[4] + d[8] * Rr[5]);
        const float dgv = (d[1] * Rr[2] + d[4] * Rr[5] + d[7] * Rr[8]) - (d[2] * Rr[1] + d[5] * Rr[4] + d[8] * Rr[7]);
        if (lane == 0) {
          A.da[n] = dav;
          A.dgamma[n] = dgv;
        }
        red_da += dav;
        red_eda += eps[r] * dav;
      }
    }
  }
  // ---- block reduction of the scalar sums
  if (lane == 0) {
    s_part[warp][0] = sums[0]; s_part[warp][1] = sums[1]; s_part[warp][2] = sums[2]; s_part[warp][3] = sums[3];
    s_part[warp][4] = red_da;  s_part[warp][5] = red_eda;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kGeomWarps; ++w) tot += s_part[w][threadIdx.x];
    if (!kFull) {
      if (threadIdx.x < 4) atomicAdd(A.loss_sums + threadIdx.x, tot);
    } else {
      if (threadIdx.x >= 4) atomicAdd(A.red + (threadIdx.x - 4), tot);
    }
  }
}

// (mean, unbiased std) of gamma = (ang0+ang1)/2 -- one block (train_leg_torso_lifter.py:153,168)
__global__ void __launch_bounds__(1024) elev_stats_kernel(const float* __restrict__ ang0, const float* __restrict__ ang1,
                                                          int N, float* __restrict__ stats) {
  __shared__ float sh[32];
  __shared__ float s_mean;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  float acc = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    acc += 0.5f * (ang0[static_cast<size_t>(i) * LINKS_HEAD_LD] + ang1[static_cast<size_t>(i) * LINKS_HEAD_LD]);
  acc = warp_sum(acc);
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarps ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) s_mean = t / static_cast<float>(N);
  }
  __syncthreads();
  const float mean = s_mean;
  float var = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float g = 0.5f * (ang0[static_cast<size_t>(i) * LINKS_HEAD_LD] + ang1[static_cast<size_t>(i) * LINKS_HEAD_LD]) - mean;
    var += g * g;
  }
  var = warp_sum(var);
  __syncthreads();
  if (lane == 0) sh[warp] = var;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarps ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) {
      stats[0] = mean;
      stats[1] = sqrtf(t / static_cast<float>(N - 1));
    }
  }
}

// Phase B of the backward: batch-statistic terms of d gamma, angle-head gradients (props = (a0+a1)/2).
//   a_n = -mu + sigma*eps_n  =>  dL/dmu = -sum da, dL/dsigma = sum eps*da,
//   dmu/dgamma_m = 1/N, dsigma/dgamma_m = (gamma_m - mu)/((N-1) sigma)
__global__ void geom_backward_angles_kernel(const float* __restrict__ ang0, const float* __restrict__ ang1,
                                            const float* __restrict__ stats, const float* __restrict__ dgamma,
                                            const float* __restrict__ red, int N, __nv_bfloat16* __restrict__ g0,
                                            __nv_bfloat16* __restrict__ g1, __nv_bfloat16* __restrict__ gT0,
                                            __nv_bfloat16* __restrict__ gT1, int ldT, int colT0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float mu = stats[0], sigma = stats[1];
  const float gam = 0.5f * (ang0[static_cast<size_t>(n) * LINKS_HEAD_LD] + ang1[static_cast<size_t>(n) * LINKS_HEAD_LD]);
  const float dmu = -red[0], dsig = red[1];
  float dg = dgamma[n] + dmu / static_cast<float>(N);
  if (sigma > 0.f) dg += dsig * (gam - mu) / (static_cast<float>(N - 1) * sigma);
  const __nv_bfloat16 h = __float2bfloat16_rn(0.5f * dg);
  g0[static_cast<size_t>(n) * 64] = h;
  g1[static_cast<size_t>(n) * 64] = h;
  if (gT0) gT0[colT0 + n] = h;
  if (gT1) gT1[colT0 + n] = h;
}

}  // namespace links

// Geometry + self-supervised losses of the lifter training step, forward and hand-derived backward.
// Reference: train_leg_torso_lifter.py:153-272 (V = 1 pose variant) and train_left_right_lifter.py:150-423
// (V = 2 variants: 'left' / 'right' choice of combine_left_right_pred_1d, utils/helpers.py:40-53).
//
// Mapping: one warp per consecutive row pair (2k, 2k+1) -- the pairwise deformation loss (:250-254) couples exactly
// those rows.  Lanes 0-15 own row 2k, lanes 16-31 row 2k+1; lane s of a half owns joint s+1 AND bone s (whose child is
// joint s+1).  The root joint needs no lane: after root-centring (:188-192) its lifted, rotated, re-lifted and
// back-rotated positions are identically zero, its depth offset is forced to 0 (:183) so it receives no gradient, and
// its only loss contribution is the constant |u_root| of the reprojection term.  Per-row reductions are 4-step
// half-warp shuffles, the pair exchange is one xor-16 shuffle, the three sin/cos pairs of a row are evaluated by three
// different lanes in one call.  Nothing but the network outputs is read: P, R, Q, q are recomputed from
// (u, depth heads, angle heads, eps_x, u_y, stats).
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
// address).  Blocks walk row pairs with a grid-stride loop and issue ONE set of atomics at the end.
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
// sum over the 16 lanes of a half-warp (every lane of the half receives it)
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(LINKS_FULL_MASK, v, o);
  return v;
}

// 1/x by the hardware approximation (<= 2 ulp): the projections divide by depths ~ 10, far from its range limits
__device__ __forceinline__ float fast_rcp(float x) { return __fdividef(1.f, x); }

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

// bone table (utils/helpers.py:140-141): bone b joins parent kBoneParent[b] and child b+1 (4 bits per entry);
// bit j of kHasNextBone: joint j is the parent of bone j; joint 8 is also the parent of bones 10 and 13, the root
// of bones 0, 3 and 6.
constexpr unsigned long long kBoneParent = 0xfe8cb89870540210ull;
constexpr unsigned kHasNextBone = 0xdbb6u;

// Lane-constant indexing state (one joint, one bone).
struct LaneMaps {
  int lane, sub, hbase, j;     // j = sub + 1
  int net[2], col;             // depth-head source of joint j per variant
  int pnet[2], pidx[2];        // which part (flow / pass-2 lifter input) receives joint j per variant, and where
  int parent_src;              // lane holding the parent joint of bone `sub`
  bool parent_is_root, has_next;
  float crel;                  // bone_rel[sub]
};
__device__ __forceinline__ void lane_maps(const GeomArgs& A, LaneMaps& m) {
  m.lane = threadIdx.x & 31;
  m.sub = m.lane & 15;
  m.hbase = m.lane & 16;
  m.j = m.sub + 1;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    m.net[v] = A.maps.src_net[v][m.j];
    m.pnet[v] = A.maps.part_net[v][m.j];
    m.pidx[v] = A.maps.part_idx[v][m.j];
  }
  m.col = A.maps.col[m.j];
  const int parent = static_cast<int>((kBoneParent >> (4 * m.sub)) & 15ull);
  m.parent_is_root = parent == 0;
  m.parent_src = m.hbase | (parent > 0 ? parent - 1 : 0);
  m.has_next = (kHasNextBone >> m.j) & 1u;
  m.crel = A.maps.bone_rel[m.sub];
}

// R = Rx(a) @ (Ry(b) @ Rx(g))   (utils/rotation_conversions.py:11-36; train_leg_torso_lifter.py:159-181).
// Lanes 0, 1, 2 of each half evaluate sincos(a), sincos(b), sincos(g) in one call; the results are broadcast.
__device__ __forceinline__ void make_rotation(const LaneMaps& m, float a, float b, float g, float (&R)[9]) {
  const int k = m.sub % 3;
  const float arg = k == 0 ? a : (k == 1 ? b : g);
  float s, c;
  sincosf(arg, &s, &c);
  const float sa = __shfl_sync(LINKS_FULL_MASK, s, m.hbase), ca = __shfl_sync(LINKS_FULL_MASK, c, m.hbase);
  const float sb = __shfl_sync(LINKS_FULL_MASK, s, m.hbase | 1), cb = __shfl_sync(LINKS_FULL_MASK, c, m.hbase | 1);
  const float sg = __shfl_sync(LINKS_FULL_MASK, s, m.hbase | 2), cg = __shfl_sync(LINKS_FULL_MASK, c, m.hbase | 2);
  R[0] = cb;       R[1] = sb * sg;                 R[2] = sb * cg;
  R[3] = sa * sb;  R[4] = ca * cg - sa * cb * sg;  R[5] = -ca * sg - sa * cb * cg;
  R[6] = -ca * sb; R[7] = sa * cg + ca * cb * sg;  R[8] = -sa * sg + ca * cb * cg;
}

// Everything a lane reads from global memory for one row: loaded one grid-stride iteration AHEAD of its use so that
// the load latency hides behind the math of the current row pair (these kernels run a few hundred dependent
// instructions per pair on few resident warps).  kLevel: 0 forward, 1 + pass-2 heads, 2 + external d/dq.
template <int V, int kLevel>
struct RawRow {
  int n;
  float ux, uy, u0x, u0y;      // this lane's 2D joint, root joint
  float ang0, ang1, eps, uyaw;
  float delta[V];              // pass-1 depth-head output of this lane's joint, per variant
  float delta2[V];             // pass-2 depth-head output
  float xqx[V], xqy[V];        // external d/d(projected joint): flow + pass-2 lifter input gradients
};
template <int V, int kLevel>
__device__ __forceinline__ void load_raw(const GeomArgs& A, const LaneMaps& m, int n, RawRow<V, kLevel>& w) {
  w.n = n;
  const size_t nn = static_cast<size_t>(n);
  const float* u = A.u + nn * 34;
  w.ux = u[m.j];
  w.uy = u[kJ + m.j];
  w.u0x = u[0];
  w.u0y = u[kJ];
  w.ang0 = A.ang[0][nn * LINKS_HEAD_LD];
  w.ang1 = A.ang[1][nn * LINKS_HEAD_LD];
  w.eps = A.eps_x[n];
  w.uyaw = A.u_y[n];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    w.delta[v] = A.head[m.net[v]][nn * LINKS_HEAD_LD + m.col];
    w.delta2[v] = 0.f; w.xqx[v] = 0.f; w.xqy[v] = 0.f;
    if (kLevel >= 1) w.delta2[v] = A.head2[m.net[v]][nn * LINKS_HEAD_LD + m.col];
    if (kLevel >= 2) {
      const int p = m.pnet[v];
      if (p >= 0) {
        const int nj = A.maps.n_joints[p];
        const int idx = m.pidx[v];
        w.xqx[v] = A.dflow[p][nn * (2 * nj) + idx] + A.dlift[p][nn * LINKS_HEAD_LD + idx];
        w.xqy[v] = A.dflow[p][nn * (2 * nj) + nj + idx] + A.dlift[p][nn * LINKS_HEAD_LD + nj + idx];
      }
    }
  }
}

// Per-row inputs shared by all variants.
struct RowIn {
  size_t n;
  float ux, uy;        // this lane's 2D joint
  float u0x, u0y;      // root joint
  float gamma, eps;
  float R[9];
};
template <int V, int kLevel>
__device__ __forceinline__ void derive_row(const GeomArgs& A, const LaneMaps& m, const RawRow<V, kLevel>& w, RowIn& r) {
  r.n = static_cast<size_t>(w.n);
  r.ux = w.ux; r.uy = w.uy; r.u0x = w.u0x; r.u0y = w.u0y;
  r.gamma = 0.5f * (w.ang0 + w.ang1);
  r.eps = w.eps;
  const float a = -A.stats[0] + A.stats[1] * r.eps;
  const float b = (w.uyaw - 0.5f) * (1.99f * 3.14159265358979323846f);
  make_rotation(m, a, b, r.gamma, r.R);
}

// Per-lane state of one (row, variant).
struct RowVar {
  float mask;        // 1 where depth not clamped (:186)
  float d;           // clamped depth
  Vec3 P;            // root-centred lifted joint (:188-192)
  Vec3 Q;            // rotated joint (:195)
  float izq, qx, qy; // projection (:198-199); izq = 1 / (Q.z + depth)
  // pass-2 side
  float mask2, d2;
  Vec3 P2;           // re-lifted joint (:235-238)
  Vec3 F;            // Q - P2
  float L3d;
  Vec3 S;            // R^T P2 (:242)
  float izs, rx, ry; // izs = 1 / (S.z + depth)
  // bones
  Vec3 e;            // P_parent - P_child of this lane's bone
  float bl_len, bl_imean, bl_rho;   // bone length, 1 / mean bone length, their ratio
};

__device__ __forceinline__ void row_forward(const GeomArgs& A, float delta, const RowIn& r, RowVar& s) {
  const float D = A.maps.depth;
  const float d0 = D < 1.0f ? 1.0f : D;                   // root depth: offset forced to 0 (:183), then clamped
  float d = delta + D;
  s.mask = (d < 1.0f) ? 0.f : 1.f;
  d = (d < 1.0f) ? 1.0f : d;
  s.d = d;
  s.P.x = r.ux * d - r.u0x * d0;
  s.P.y = r.uy * d - r.u0y * d0;
  s.P.z = d - d0;
  s.Q = mat_vec(r.R, s.P);
  s.izq = fast_rcp(s.Q.z + D);
  s.qx = s.Q.x * s.izq;
  s.qy = s.Q.y * s.izq;
}

// pass-2 quantities and the per-row loss terms (L3d, rep_rot, bl_prior) via out[3] (uniform over the half-warp).
__device__ __forceinline__ void row_consistency(const GeomArgs& A, const LaneMaps& m, float delta2, const RowIn& r,
                                                RowVar& s, float (&out)[3]) {
  const float D = A.maps.depth;
  const float d0 = D < 1.0f ? 1.0f : D;
  float d2 = delta2 + D;
  s.mask2 = (d2 < 1.0f) ? 0.f : 1.f;
  d2 = (d2 < 1.0f) ? 1.0f : d2;
  s.d2 = d2;
  // the root projects to (0, 0): its re-lifted position is (0, 0, d0)
  s.P2.x = s.qx * d2;
  s.P2.y = s.qy * d2;
  s.P2.z = d2 - d0;
  s.F.x = s.Q.x - s.P2.x; s.F.y = s.Q.y - s.P2.y; s.F.z = s.Q.z - s.P2.z;
  s.L3d = sqrtf(half_sum(s.F.x * s.F.x + s.F.y * s.F.y + s.F.z * s.F.z));
  s.S = matT_vec(r.R, s.P2);
  s.izs = fast_rcp(s.S.z + D);
  s.rx = s.S.x * s.izs;
  s.ry = s.S.y * s.izs;
  const float rep = half_sum(fabsf(s.rx - r.ux) + fabsf(s.ry - r.uy)) + (fabsf(r.u0x) + fabsf(r.u0y));
  // bone lengths: this lane's bone joins its joint (child) to the parent joint
  Vec3 Pp = shfl3(s.P, m.parent_src);
  if (m.parent_is_root) { Pp.x = 0.f; Pp.y = 0.f; Pp.z = 0.f; }
  s.e.x = Pp.x - s.P.x; s.e.y = Pp.y - s.P.y; s.e.z = Pp.z - s.P.z;
  s.bl_len = sqrtf(s.e.x * s.e.x + s.e.y * s.e.y + s.e.z * s.e.z);
  s.bl_imean = fast_rcp(half_sum(s.bl_len) * (1.f / 16.f));
  s.bl_rho = s.bl_len * s.bl_imean;
  const float bl = half_sum((m.crel - s.bl_rho) * (m.crel - s.bl_rho));
  out[0] = s.L3d;
  out[1] = rep;
  out[2] = bl;
}

// =========================================================================================================
// forward: projected parts for the flows / pass-2 lifters
// =========================================================================================================
template <int V>
__global__ void __launch_bounds__(kGeomWarps * 32) geom_forward_kernel(const GeomArgs Ap) {
  __shared__ GeomArgs sA;
  stage_args(&sA, Ap);
  const GeomArgs& A = sA;
  LaneMaps m;
  lane_maps(A, m);
  const int half = m.lane >> 4;
  const int npairs = (A.N + 1) / 2;
  const int stride = gridDim.x * kGeomWarps;
  // The loop bounds depend on blockIdx only (block-uniform trip count): the compiler can then prove that the warp is
  // converged at every shuffle and emits plain SHFLs; a warp past the end works on a clamped row with writes masked.
  const int warp = threadIdx.x >> 5;
  auto row_of = [&](int pr) {                    // row this half-warp reads for pair `pr` (clamped to a valid row)
    const int n = 2 * pr + half;
    return n < A.N ? n : A.N - 1;
  };
  RawRow<V, 0> raw, raw_next;
  load_raw(A, m, row_of(blockIdx.x * kGeomWarps + warp), raw);
  for (int base = blockIdx.x * kGeomWarps; base < npairs; base += stride, raw = raw_next) {
    const int pair = base + warp;
    if (base + stride < npairs) load_raw(A, m, row_of(pair + stride), raw_next);
    const bool valid = 2 * pair + half < A.N;    // uniform over the half-warp
    RowIn r;
    derive_row(A, m, raw, r);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      RowVar s;
      row_forward(A, raw.delta[v], r, s);
      if (!valid) continue;
      if (A.qfull[v]) {
        float* qf = A.qfull[v] + r.n * 34;
        qf[m.j] = s.qx;
        qf[kJ + m.j] = s.qy;
        if (m.sub == 0) { qf[0] = 0.f; qf[kJ] = 0.f; }          // root: projects to (0, 0)
      }
      const int p = m.pnet[v];
      if (p >= 0) {
        const int nj = A.maps.n_joints[p];
        float* dst = A.qpart[p] + r.n * (2 * nj);
        dst[m.pidx[v]] = s.qx;
        dst[nj + m.pidx[v]] = s.qy;
      }
      if (m.sub == 0) {
        const int p0 = A.maps.part_net[v][0];
        if (p0 >= 0) {
          const int nj = A.maps.n_joints[p0];
          float* dst = A.qpart[p0] + r.n * (2 * nj);
          dst[A.maps.part_idx[v][0]] = 0.f;
          dst[nj + A.maps.part_idx[v][0]] = 0.f;
        }
      }
    }
  }
}

// =========================================================================================================
// losses + gradients.  kFull = false: loss sums and d/d(pass-2 heads) only (runs before the pass-2 backward);
// kFull = true : complete backward to the pass-1 heads, d gamma (direct) and d a (runs after it).
// V (number of pose variants, = maps.V) is a template parameter so that the variant loop unrolls: the per-variant lane
// maps stay in registers and every shuffle sits in straight-line, provably convergent code.
// =========================================================================================================
template <bool kFull, int V>
__global__ void __launch_bounds__(kGeomWarps * 32) geom_lossgrad_kernel(const GeomArgs Ap) {
  __shared__ float s_part[kGeomWarps][6];
  __shared__ GeomArgs sA;
  stage_args(&sA, Ap);
  const GeomArgs& A = sA;
  LaneMaps m;
  lane_maps(A, m);
  const int lane = m.lane;
  const int warp = threadIdx.x >> 5;
  const int half = lane >> 4;
  const float invN = 1.f / static_cast<float>(A.N);
  const int npairs = A.N / 2;
  const float c3d = A.maps.w_3d * invN, c2d = A.maps.w_2d * invN, cbl = A.maps.w_bl * invN;
  const float cv = npairs > 0 ? A.maps.w_vel / static_cast<float>(npairs) : 0.f;
  // which depth heads ever feed this lane's joint / the root joint (-> which gradient columns this lane writes)
  bool feeds[2] = {false, false}, feeds_root[2] = {false, false};
#pragma unroll
  for (int v = 0; v < V; ++v) {
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      feeds[net] = feeds[net] || m.net[v] == net;
      feeds_root[net] = feeds_root[net] || A.maps.src_net[v][0] == net;
    }
  }
  const int col_root = A.maps.col[0];

  float sums[4] = {0.f, 0.f, 0.f, 0.f};   // L3d, rep, pair, bl (raw sums of this half-warp's rows)
  float red_da = 0.f, red_eda = 0.f;

  const int total_pairs = (A.N + 1) / 2;
  const int stride = gridDim.x * kGeomWarps;
  // block-uniform trip count (see geom_forward_kernel): shuffles sit in provably convergent code
  auto row_of = [&](int pr) {
    const int n = 2 * pr + half;
    return n < A.N ? n : A.N - 1;
  };
  RawRow<V, kFull ? 2 : 1> raw, raw_next;
  load_raw(A, m, row_of(blockIdx.x * kGeomWarps + warp), raw);
  for (int base = blockIdx.x * kGeomWarps; base < total_pairs; base += stride, raw = raw_next) {
    const int pair = base + warp;
    if (base + stride < total_pairs) load_raw(A, m, row_of(pair + stride), raw_next);
    const bool vB = 2 * pair + 1 < A.N;                           // warp-uniform: the pair is complete
    const bool valid = 2 * pair + half < A.N;                     // uniform over the half-warp
    RowIn r;
    derive_row(A, m, raw, r);
    float dRm[9];                                                 // lane-partial d/dR
#pragma unroll
    for (int k = 0; k < 9; ++k) dRm[k] = 0.f;
    float g1acc[2] = {0.f, 0.f};   // [net] d/d(pass-1 head) for this lane's column
    float g2acc[2] = {0.f, 0.f};

#pragma unroll
    for (int v = 0; v < V; ++v) {
      RowVar t;
      float o[3];
      row_forward(A, raw.delta[v], r, t);
      row_consistency(A, m, raw.delta2[v], r, t, o);
      if (valid) { sums[0] += o[0]; sums[1] += o[1]; sums[3] += o[2]; }
      // pairwise deformation (:250-254): E = (P - P') - (S - S'), ' = the other row of the pair
      Vec3 E; E.x = E.y = E.z = 0.f;
      float pnorm = 0.f;
      {
        Vec3 dPS; dPS.x = t.P.x - t.S.x; dPS.y = t.P.y - t.S.y; dPS.z = t.P.z - t.S.z;
        const float ox = __shfl_xor_sync(LINKS_FULL_MASK, dPS.x, 16);
        const float oy = __shfl_xor_sync(LINKS_FULL_MASK, dPS.y, 16);
        const float oz = __shfl_xor_sync(LINKS_FULL_MASK, dPS.z, 16);
        if (vB) { E.x = dPS.x - ox; E.y = dPS.y - oy; E.z = dPS.z - oz; }
        pnorm = sqrtf(half_sum(E.x * E.x + E.y * E.y + E.z * E.z));
        if (half == 0) sums[2] += pnorm;
      }
      const float ge = pnorm > 0.f ? cv / pnorm : 0.f;
      // ---- d/dS: reprojection L1 (:247) + pair term
      const float drx = c2d * ((t.rx > r.ux) ? 1.f : ((t.rx < r.ux) ? -1.f : 0.f));
      const float dry = c2d * ((t.ry > r.uy) ? 1.f : ((t.ry < r.uy) ? -1.f : 0.f));
      Vec3 dS;
      dS.x = drx * t.izs - ge * E.x;
      dS.y = dry * t.izs - ge * E.y;
      dS.z = -(drx * t.rx + dry * t.ry) * t.izs - ge * E.z;
      // ---- d/dP2 = R dS - c3d F / L3d   (root centring only feeds the root's own, constant, depth)
      const float g3 = t.L3d > 0.f ? c3d / t.L3d : 0.f;
      Vec3 dP2 = mat_vec(r.R, dS);
      dP2.x -= g3 * t.F.x; dP2.y -= g3 * t.F.y; dP2.z -= g3 * t.F.z;
      const float ddelta2 = t.mask2 * (dP2.x * t.qx + dP2.y * t.qy + dP2.z);
      const int net = m.net[v];
      g2acc[0] += net == 0 ? ddelta2 : 0.f;
      g2acc[1] += net == 1 ? ddelta2 : 0.f;
      if (kFull) {
        // ---- d/dq: through P2 = (q d2) and the external consumers (flows, pass-2 lifters)
        const float dqx = dP2.x * t.d2 + raw.xqx[v], dqy = dP2.y * t.d2 + raw.xqy[v];
        // ---- d/dQ
        Vec3 dQ;
        dQ.x = g3 * t.F.x + dqx * t.izq;
        dQ.y = g3 * t.F.y + dqy * t.izq;
        dQ.z = g3 * t.F.z - (dqx * t.qx + dqy * t.qy) * t.izq;
        // ---- d/dP = R^T dQ + pair + bones
        Vec3 dP = matT_vec(r.R, dQ);
        dP.x += ge * E.x; dP.y += ge * E.y; dP.z += ge * E.z;
        {
          const float h = -2.f * (m.crel - t.bl_rho) * cbl;
          const float hl = half_sum(h * t.bl_len);
          const float dl = (h - hl * (1.f / 16.f) * t.bl_imean) * t.bl_imean;
          const float w = t.bl_len > 0.f ? dl * fast_rcp(t.bl_len) : 0.f;
          Vec3 dv;                                   // d/d(P_parent) of this lane's bone; d/d(P_child) = -dv
          dv.x = w * t.e.x; dv.y = w * t.e.y; dv.z = w * t.e.z;
          dP.x -= dv.x; dP.y -= dv.y; dP.z -= dv.z;
          const Vec3 nx = shfl3(dv, lane + 1);       // bone j (child j+1) lives in the next lane
          if (m.has_next) { dP.x += nx.x; dP.y += nx.y; dP.z += nx.z; }
          const Vec3 b10 = shfl3(dv, m.hbase | 10), b13 = shfl3(dv, m.hbase | 13);
          if (m.j == 8) { dP.x += b10.x + b13.x; dP.y += b10.y + b13.y; dP.z += b10.z + b13.z; }
        }
        // ---- d/dR from Q = R P and S = R^T P2 (lane-partial; reduced once per row below)
        const float dq3[3] = {dQ.x, dQ.y, dQ.z}, p3[3] = {t.P.x, t.P.y, t.P.z};
        const float p23[3] = {t.P2.x, t.P2.y, t.P2.z}, ds3[3] = {dS.x, dS.y, dS.z};
#pragma unroll
        for (int ra = 0; ra < 3; ++ra)
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) dRm[ra * 3 + cb] += dq3[ra] * p3[cb] + p23[ra] * ds3[cb];
        // ---- lift
        const float ddelta = t.mask * (dP.x * r.ux + dP.y * r.uy + dP.z);
        g1acc[0] += net == 0 ? ddelta : 0.f;
        g1acc[1] += net == 1 ? ddelta : 0.f;
      }
    }
    // ---- write head gradients (lane -> column col[j] of every net that feeds joint j in some variant; the root's
    //      columns receive zeros)
    if (valid) {
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        __nv_bfloat16* g = kFull ? A.g1[net] : A.g2[net];
        __nv_bfloat16* gT = kFull ? A.g1T[net] : A.g2T[net];
        if (feeds[net]) {
          const __nv_bfloat16 h = __float2bfloat16_rn(kFull ? g1acc[net] : g2acc[net]);
          g[r.n * 64 + m.col] = h;
          if (gT) gT[static_cast<size_t>(m.col) * A.ldT + A.colT0 + r.n] = h;
        }
        if (m.sub == 0 && feeds_root[net]) {
          const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
          g[r.n * 64 + col_root] = z;
          if (gT) gT[static_cast<size_t>(col_root) * A.ldT + A.colT0 + r.n] = z;
        }
      }
    }
    if (kFull) {
      const float* Rr = r.R;
      const float* d = dRm;
      // dR/da: row1' = -row2, row2' = row1 ; dR/dgamma: col1' = col2, col2' = -col1
      const float dav = half_sum(-(d[3] * Rr[6] + d[4] * Rr[7] + d[5] * Rr[8]) + (d[6] * Rr[3] + d[7] * Rr[4] + d[8] * Rr[5]));
      const float dgv = half_sum((d[1] * Rr[2] + d[4] * Rr[5] + d[7] * Rr[8]) - (d[2] * Rr[1] + d[5] * Rr[4] + d[8] * Rr[7]));
      if (valid) {
        if (m.sub == 0) {
          A.da[r.n] = dav;
          A.dgamma[r.n] = dgv;
        }
        red_da += dav;
        red_eda += r.eps * dav;
      }
    }
  }
  // ---- block reduction of the scalar sums (values are uniform over each half-warp: add the two halves)
#pragma unroll
  for (int k = 0; k < 4; ++k) sums[k] += __shfl_xor_sync(LINKS_FULL_MASK, sums[k], 16);
  red_da += __shfl_xor_sync(LINKS_FULL_MASK, red_da, 16);
  red_eda += __shfl_xor_sync(LINKS_FULL_MASK, red_eda, 16);
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

// Data-parallel "global elevation statistics" (train_leg_torso_lifter.py:168 evaluated over the GLOBAL batch instead of the
// rank's shard): per-rank sums (sum gamma, sum gamma^2) in double -> all-reduce -> finalize.
__global__ void __launch_bounds__(1024) elev_sums_kernel(const float* __restrict__ ang0, const float* __restrict__ ang1,
                                                         int N, double* __restrict__ sums) {
  __shared__ double sh[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double g = 0.5 * (static_cast<double>(ang0[static_cast<size_t>(i) * LINKS_HEAD_LD]) +
                            static_cast<double>(ang1[static_cast<size_t>(i) * LINKS_HEAD_LD]));
    a += g;
    b += g * g;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(LINKS_FULL_MASK, a, o);
    b += __shfl_xor_sync(LINKS_FULL_MASK, b, o);
  }
  if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < nwarps; ++w) { ta += sh[0][w]; tb += sh[1][w]; }
    sums[0] = ta;
    sums[1] = tb;
  }
}
__global__ void elev_finalize_kernel(const double* __restrict__ sums, double n_total, float* __restrict__ stats) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double mean = sums[0] / n_total;
  double var = (sums[1] - n_total * mean * mean) / (n_total - 1.0);
  if (var < 0.0) var = 0.0;
  stats[0] = static_cast<float>(mean);
  stats[1] = static_cast<float>(sqrt(var));
}

// Phase B of the backward: batch-statistic terms of d gamma, angle-head gradients (props = (a0+a1)/2).
//   a_n = -mu + sigma*eps_n  =>  dL/dmu = -sum da, dL/dsigma = sum eps*da,
//   dmu/dgamma_m = 1/N, dsigma/dgamma_m = (gamma_m - mu)/((N-1) sigma)
__global__ void geom_backward_angles_kernel(const float* __restrict__ ang0, const float* __restrict__ ang1,
                                            const float* __restrict__ stats, const float* __restrict__ dgamma,
                                            const float* __restrict__ red, int N, __nv_bfloat16* __restrict__ g0,
                                            __nv_bfloat16* __restrict__ g1, __nv_bfloat16* __restrict__ gT0,
                                            __nv_bfloat16* __restrict__ gT1, int ldT, int colT0, int n_stat) {
  // n_stat: rows behind the statistic (= N, or the global row count when stats / red were reduced over the ranks)
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float mu = stats[0], sigma = stats[1];
  const float gam = 0.5f * (ang0[static_cast<size_t>(n) * LINKS_HEAD_LD] + ang1[static_cast<size_t>(n) * LINKS_HEAD_LD]);
  const float dmu = -red[0], dsig = red[1];
  float dg = dgamma[n] + dmu / static_cast<float>(n_stat);
  if (sigma > 0.f) dg += dsig * (gam - mu) / (static_cast<float>(n_stat - 1) * sigma);
  const __nv_bfloat16 h = __float2bfloat16_rn(0.5f * dg);
  g0[static_cast<size_t>(n) * 64] = h;
  g1[static_cast<size_t>(n) * 64] = h;
  if (gT0) gT0[colT0 + n] = h;
  if (gT1) gT1[colT0 + n] = h;
}

}  // namespace links

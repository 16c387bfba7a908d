// Geometry + self-supervised losses of the lifter training step, forward and hand-derived backward.
// Reference: train_leg_torso_lifter.py:153-272 (V = 1 pose variant) and train_left_right_lifter.py:150-423
// (V = 2 variants: 'left' / 'right' choice of combine_left_right_pred_1d, utils/helpers.py:40-53).
//
// Mapping: four lanes per row, eight consecutive rows (four row pairs) per warp.  Lane q of a row's quad owns joints
// 4q+1 .. 4q+4 and the bones whose children they are ("slots" 0-3): the per-joint math of a slot is straight-line code over
// registers, the row-uniform work (rotation, reductions) is shared by 4 lanes instead of being replicated in 16, per-row
// reductions are two xor-shuffles, the pairwise deformation loss (:250-254) couples rows 2k and 2k+1 = lanes l and l^4.
// The root joint needs no slot: after root-centring (:188-192) its lifted, rotated, re-lifted and back-rotated positions
// are identically zero, its depth offset is forced to 0 (:183) so it receives no gradient, and its only loss contribution
// is the constant |u_root| of the reprojection term.  Nothing but the network outputs is read: P, R, Q, q are recomputed
// from (u, depth heads, angle heads, eps_x, u_y, stats).
#pragma once
#include "devdefs.cuh"

namespace links {

constexpr int kGeomWarps = 4;
constexpr int kGeomRows = 8;    // rows per warp and grid-stride iteration
constexpr int kGeomMaxRows = (1 << 25) - 8;   // element offsets (row * 64 + column) are formed in 32 bits
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

// 1/x and sqrt(x) by the hardware approximations (MUFU.RCP / MUFU.SQRT, <= 2 ulp, one instruction each, no denormal or
// range fix-up code): the projections divide by depths ~ 10, the roots are of sums of squares -- far from the range limits
__device__ __forceinline__ float fast_rcp(float x) {
#ifndef LINKS_HOSTSIM
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.f / x;
#endif
}
__device__ __forceinline__ float fast_sqrt(float x) {
#ifndef LINKS_HOSTSIM
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
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

// Per-block lookup tables (shared memory, built once per block): for every (variant, joint) the ELEMENT of row 0 that
// a lane reads or writes, so that the row loop forms an address with one 32 x 32 + 64 bit multiply-add from one 8-byte
// table entry instead of chasing net / column / part indices through the maps.  Entries without a source point at a zero
// word with pitch 0 (loads stay unconditional); entries without a destination are null (stores are predicated).
__device__ float g_geom_zero[4];
struct GeomTabs {
  const float* hp[2][kJ];     // pass-1 depth-head output of joint j in variant v
  const float* h2p[2][kJ];    // pass-2 depth-head output
  const float* dfx[2][kJ];    // d/d(projected x) from the part flow; y sits njy floats further
  const float* dlx[2][kJ];    // ... from the pass-2 lifter input gradient
  float* qp[2][kJ];           // projected x in the part's flow / pass-2 input row; y sits njy floats further
  __nv_bfloat16* gp[2][kJ];   // [net][j]: head-gradient element, null when net never feeds joint j
  int pitch_f[2][kJ], pitch_l[2][kJ], njy[2][kJ];
};
template <int V>
__device__ __forceinline__ void build_tabs(const GeomArgs& A, GeomTabs& T, bool full) {
  for (int i = threadIdx.x; i < 2 * kJ; i += blockDim.x) {
    const int v = i / kJ, j = i - v * kJ;
    // gradient destinations are per NET (first index), everything else per VARIANT
    {
      bool fed = false;
#pragma unroll
      for (int w = 0; w < V; ++w) fed = fed || A.maps.src_net[w][j] == v;
      __nv_bfloat16* g = full ? A.g1[v] : A.g2[v];
      T.gp[v][j] = (fed && g) ? g + A.maps.col[j] : nullptr;
    }
    if (v >= V) continue;
    const int net = A.maps.src_net[v][j], col = A.maps.col[j];
    T.hp[v][j] = A.head[net] + col;
    T.h2p[v][j] = A.head2[net] ? A.head2[net] + col : g_geom_zero;
    const int pn = A.maps.part_net[v][j];
    const bool has = pn >= 0;
    const int pi = has ? pn : 0;
    const int nj = A.maps.n_joints[pi], idx = A.maps.part_idx[v][j];
    const bool ext = has && A.dflow[pi] && A.dlift[pi];
    T.dfx[v][j] = ext ? A.dflow[pi] + idx : g_geom_zero;
    T.dlx[v][j] = ext ? A.dlift[pi] + idx : g_geom_zero;
    T.pitch_f[v][j] = ext ? 2 * nj : 0;
    T.pitch_l[v][j] = ext ? LINKS_HEAD_LD : 0;
    T.njy[v][j] = has ? nj : 0;
    T.qp[v][j] = (has && A.qpart[pi]) ? A.qpart[pi] + idx : nullptr;
  }
  __syncthreads();
}

// bone table (utils/helpers.py:140-141): bone b joins parent kBoneParent[b] and child b+1 (4 bits per entry)
constexpr unsigned long long kBoneParent = 0xfe8cb89870540210ull;
__host__ __device__ constexpr int bone_parent(int child) { return static_cast<int>((kBoneParent >> (4 * (child - 1))) & 15ull); }
// The quad mapping below hard-wires where a slot finds the parent joint of its bone; tie it to the table.
//   slot 0 (joints 1, 5, 9, 13): root for quad lane 0, else slot 3 of the previous lane (joints 4, 8, 12)
//   slot 1 (2, 6, 10, 14): own slot 0, except joint 14 <- joint 8 (lane 1, slot 3)
//   slot 2 (3, 7, 11, 15): own slot 1, except joint 7 <- root and joint 11 <- joint 8
//   slot 3 (4, 8, 12, 16): own slot 2, except joint 4 <- root
static_assert(bone_parent(1) == 0 && bone_parent(5) == 4 && bone_parent(9) == 8 && bone_parent(13) == 12, "slot 0 parents");
static_assert(bone_parent(2) == 1 && bone_parent(6) == 5 && bone_parent(10) == 9 && bone_parent(14) == 8, "slot 1 parents");
static_assert(bone_parent(3) == 2 && bone_parent(7) == 0 && bone_parent(11) == 8 && bone_parent(15) == 14, "slot 2 parents");
static_assert(bone_parent(4) == 0 && bone_parent(8) == 7 && bone_parent(12) == 11 && bone_parent(16) == 15, "slot 3 parents");

// Lane-constant indexing state.
struct Quad {
  int lane, q, gb, j0;   // q = lane & 3; gb = first lane of this row's quad; j0 = 4 q + 1 = joint of slot 0
};
__device__ __forceinline__ void quad_init(Quad& m) {
  m.lane = threadIdx.x & 31;
  m.q = m.lane & 3;
  m.gb = m.lane & ~3;
  m.j0 = 4 * m.q + 1;
}
__device__ __forceinline__ float quad_sum(float v) {          // sum over the 4 lanes of a row (every lane receives it)
  v += __shfl_xor_sync(LINKS_FULL_MASK, v, 1);
  v += __shfl_xor_sync(LINKS_FULL_MASK, v, 2);
  return v;
}

// R = Rx(a) @ (Ry(b) @ Rx(g))   (utils/rotation_conversions.py:11-36; train_leg_torso_lifter.py:159-181).
// Lanes 0, 1, 2 of a row's quad evaluate sincos(a), sincos(b), sincos(g) in one call; the results are broadcast.
__device__ __forceinline__ void make_rotation(const Quad& m, float a, float b, float g, float (&R)[9]) {
  const float arg = m.q == 0 ? a : (m.q == 1 ? b : g);
  float s, c;
  sincosf(arg, &s, &c);
  const float sa = __shfl_sync(LINKS_FULL_MASK, s, m.gb), ca = __shfl_sync(LINKS_FULL_MASK, c, m.gb);
  const float sb = __shfl_sync(LINKS_FULL_MASK, s, m.gb | 1), cb = __shfl_sync(LINKS_FULL_MASK, c, m.gb | 1);
  const float sg = __shfl_sync(LINKS_FULL_MASK, s, m.gb | 2), cg = __shfl_sync(LINKS_FULL_MASK, c, m.gb | 2);
  R[0] = cb;       R[1] = sb * sg;                 R[2] = sb * cg;
  R[3] = sa * sb;  R[4] = ca * cg - sa * cb * sg;  R[5] = -ca * sg - sa * cb * cg;
  R[6] = -ca * sb; R[7] = sa * cg + ca * cb * sg;  R[8] = -sa * sg + ca * cb * cg;
}

// Per-row inputs shared by all variants: this lane's four 2D joints, the root joint, the rotation.
struct RowIn {
  int n;               // row (element offsets are formed in 32 bits: the host checks N * 64 < 2^31)
  float ux[4], uy[4];
  float u0x, u0y;
  float eps;
  float R[9];
};
__device__ __forceinline__ void load_row(const GeomArgs& A, const Quad& m, int n, RowIn& r) {
  r.n = n;
  const float* u = A.u + n * 34;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r.ux[k] = __ldg(u + m.j0 + k);
    r.uy[k] = __ldg(u + kJ + m.j0 + k);
  }
  r.u0x = __ldg(u);
  r.u0y = __ldg(u + kJ);
  const float ang0 = __ldg(A.ang[0] + n * LINKS_HEAD_LD), ang1 = __ldg(A.ang[1] + n * LINKS_HEAD_LD);
  r.eps = __ldg(A.eps_x + n);
  const float uyaw = __ldg(A.u_y + n);
  const float gamma = 0.5f * (ang0 + ang1);
  const float a = -A.stats[0] + A.stats[1] * r.eps;
  const float b = (uyaw - 0.5f) * (1.99f * 3.14159265358979323846f);
  make_rotation(m, a, b, gamma, r.R);
}

// lift (:183-192), rotate (:195), project (:198-199) one joint
struct JointFwd {
  float mask;        // 1 where the depth is not clamped (:186)
  Vec3 P;            // root-centred lifted joint
  Vec3 Q;            // rotated joint
  float izq, qx, qy; // izq = 1 / (Q.z + depth)
};
__device__ __forceinline__ void joint_forward(float D, float d0, float delta, float ux, float uy, float u0x, float u0y,
                                              const float (&R)[9], JointFwd& s) {
  float d = delta + D;
  s.mask = (d < 1.0f) ? 0.f : 1.f;
  d = (d < 1.0f) ? 1.0f : d;
  s.P.x = ux * d - u0x * d0;
  s.P.y = uy * d - u0y * d0;
  s.P.z = d - d0;
  s.Q = mat_vec(R, s.P);
  s.izq = fast_rcp(s.Q.z + D);
  s.qx = s.Q.x * s.izq;
  s.qy = s.Q.y * s.izq;
}

// =========================================================================================================
// forward: projected parts for the flows / pass-2 lifters
// =========================================================================================================
template <int V>
__global__ void __launch_bounds__(kGeomWarps * 32) geom_forward_kernel(const GeomArgs Ap) {
  __shared__ GeomArgs sA;
  __shared__ GeomTabs sT;
  stage_args(&sA, Ap);
  const GeomArgs& A = sA;
  build_tabs<V>(A, sT, false);
  const GeomTabs& T = sT;
  Quad m;
  quad_init(m);
  const int warp = threadIdx.x >> 5, rl = m.lane >> 2;
  const float D = A.maps.depth;
  const float d0 = D < 1.0f ? 1.0f : D;                   // root depth: offset forced to 0 (:183), then clamped
  const int n_iters = (A.N + kGeomRows - 1) / kGeomRows;
  const int stride = gridDim.x * kGeomWarps;
  // The loop bounds depend on blockIdx only (block-uniform trip count): the compiler can then prove that the warp is
  // converged at every shuffle and emits plain SHFLs; rows past the end work on a clamped row with writes masked.
  for (int base = blockIdx.x * kGeomWarps; base < n_iters; base += stride) {
    const int n_raw = kGeomRows * (base + warp) + rl;
    const bool valid = n_raw < A.N;                      // uniform over the quad
    RowIn r;
    load_row(A, m, valid ? n_raw : A.N - 1, r);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float delta[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) delta[k] = __ldg(T.hp[v][m.j0 + k] + r.n * LINKS_HEAD_LD);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = m.j0 + k;
        JointFwd s;
        joint_forward(D, d0, delta[k], r.ux[k], r.uy[k], r.u0x, r.u0y, r.R, s);
        if (!valid) continue;
        if (A.qfull[v]) {
          float* qf = A.qfull[v] + r.n * 34;
          qf[j] = s.qx;
          qf[kJ + j] = s.qy;
        }
        float* dst = T.qp[v][j];
        if (dst) {
          const int nj = T.njy[v][j];
          dst += r.n * (2 * nj);
          dst[0] = s.qx;
          dst[nj] = s.qy;
        }
      }
      if (valid && m.q == 0) {                             // root: projects to (0, 0)
        if (A.qfull[v]) {
          float* qf = A.qfull[v] + r.n * 34;
          qf[0] = 0.f; qf[kJ] = 0.f;
        }
        float* dst = T.qp[v][0];
        if (dst) {
          const int nj = T.njy[v][0];
          dst += r.n * (2 * nj);
          dst[0] = 0.f;
          dst[nj] = 0.f;
        }
      }
    }
  }
}

// =========================================================================================================
// losses + gradients.  kFull = false: loss sums and d/d(pass-2 heads) only (runs before the pass-2 backward);
// kFull = true : complete backward to the pass-1 heads, d gamma (direct) and d a (runs after it).
// V (number of pose variants, = maps.V) is a template parameter so that the variant loop unrolls; kT: also write the
// transposed copies g*T (kept for the C ABI; the training step does not use them).
// Per (row, variant): phase 1 = forward quantities of the lane's four joints + the row reductions (|F|, reprojection,
// pair distance, bone lengths); phase 2 = bone-prior terms (need the mean bone length); phase 3 = backward per joint.
// d/da and d/dgamma do not go through a 3x3 d/dR: with R = Rx(a) Ry(b) Rx(g), dR/da = [x]x R and dR/dg = R [x]x, so
//   dL/da = sum_j  dQ . (x ^ Q)  -  (R dS) . (x ^ P2),      dL/dg = sum_j (R^T dQ) . (x ^ P)  -  dS . (x ^ S)
// which reuses R dS and R^T dQ of the backward chain (x ^ v = (0, -v.z, v.y)).
// =========================================================================================================
template <bool kFull, int V, bool kT = false>
__global__ void __launch_bounds__(kGeomWarps * 32) geom_lossgrad_kernel(const GeomArgs Ap) {
  __shared__ float s_part[kGeomWarps][6];
  __shared__ GeomArgs sA;
  __shared__ GeomTabs sT;
  stage_args(&sA, Ap);
  const GeomArgs& A = sA;
  build_tabs<V>(A, sT, kFull);
  const GeomTabs& T = sT;
  Quad m;
  quad_init(m);
  const int warp = threadIdx.x >> 5, rl = m.lane >> 2;
  const float invN = 1.f / static_cast<float>(A.N);
  const int npairs = A.N / 2;
  const float c3d = A.maps.w_3d * invN, c2d = A.maps.w_2d * invN, cbl = A.maps.w_bl * invN;
  const float cv = npairs > 0 ? A.maps.w_vel / static_cast<float>(npairs) : 0.f;
  const float D = A.maps.depth;
  const float d0 = D < 1.0f ? 1.0f : D;
  float sums[4] = {0.f, 0.f, 0.f, 0.f};   // L3d, rep, pair, bl (raw sums; lane 0 of each quad accumulates its rows)
  float red_da = 0.f, red_eda = 0.f;

  const int n_iters = (A.N + kGeomRows - 1) / kGeomRows;
  const int stride = gridDim.x * kGeomWarps;
  // block-uniform trip count (see geom_forward_kernel): shuffles sit in provably convergent code
  for (int base = blockIdx.x * kGeomWarps; base < n_iters; base += stride) {
    const int n_raw = kGeomRows * (base + warp) + rl;
    const bool valid = n_raw < A.N;                               // uniform over the quad
    const bool vB = (n_raw | 1) < A.N;                            // uniform over the pair's 8 lanes: the pair is complete
    RowIn r;
    load_row(A, m, valid ? n_raw : A.N - 1, r);
    float g1acc[4][2], g2acc[4][2];                               // [slot][net] d/d(head) of the slot's column
#pragma unroll
    for (int k = 0; k < 4; ++k) { g1acc[k][0] = g1acc[k][1] = g2acc[k][0] = g2acc[k][1] = 0.f; }
    float da_acc = 0.f, dg_acc = 0.f;                             // lane-partial d/da, d/dgamma

#pragma unroll
    for (int v = 0; v < V; ++v) {
      // ---- everything this (row, variant) reads, issued together
      float delta[4], delta2[4], xqx[4], xqy[4];
      bool net1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = m.j0 + k;
        net1[k] = A.maps.src_net[v][j] != 0;
        delta[k] = __ldg(T.hp[v][j] + r.n * LINKS_HEAD_LD);
        delta2[k] = __ldg(T.h2p[v][j] + r.n * LINKS_HEAD_LD);
        xqx[k] = 0.f; xqy[k] = 0.f;
        if (kFull) {
          const float* df = T.dfx[v][j] + r.n * T.pitch_f[v][j];
          const float* dl = T.dlx[v][j] + r.n * T.pitch_l[v][j];
          const int nj = T.njy[v][j];
          xqx[k] = __ldg(df) + __ldg(dl);
          xqy[k] = __ldg(df + nj) + __ldg(dl + nj);
        }
      }
      // ---- phase 1
      JointFwd f[4];
      float mask2[4], d2[4], izs[4], rx[4], ry[4], len[4];
      Vec3 P2[4], F[4], S[4], E[4], e[4];
      float f2 = 0.f, rep = 0.f, e2 = 0.f, lsum = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        joint_forward(D, d0, delta[k], r.ux[k], r.uy[k], r.u0x, r.u0y, r.R, f[k]);
        // re-lift (:228-238); the root projects to (0, 0): its re-lifted position is (0, 0, d0)
        float dd = delta2[k] + D;
        mask2[k] = (dd < 1.0f) ? 0.f : 1.f;
        dd = (dd < 1.0f) ? 1.0f : dd;
        d2[k] = dd;
        P2[k].x = f[k].qx * dd; P2[k].y = f[k].qy * dd; P2[k].z = dd - d0;
        F[k].x = f[k].Q.x - P2[k].x; F[k].y = f[k].Q.y - P2[k].y; F[k].z = f[k].Q.z - P2[k].z;
        f2 += F[k].x * F[k].x + F[k].y * F[k].y + F[k].z * F[k].z;
        // rotate back and re-project (:242-247)
        S[k] = matT_vec(r.R, P2[k]);
        izs[k] = fast_rcp(S[k].z + D);
        rx[k] = S[k].x * izs[k];
        ry[k] = S[k].y * izs[k];
        rep += fabsf(rx[k] - r.ux[k]) + fabsf(ry[k] - r.uy[k]);
        // pairwise deformation (:250-254): E = (P - P') - (S - S'), ' = the other row of the pair
        Vec3 dPS; dPS.x = f[k].P.x - S[k].x; dPS.y = f[k].P.y - S[k].y; dPS.z = f[k].P.z - S[k].z;
        const float ox = __shfl_xor_sync(LINKS_FULL_MASK, dPS.x, 4);
        const float oy = __shfl_xor_sync(LINKS_FULL_MASK, dPS.y, 4);
        const float oz = __shfl_xor_sync(LINKS_FULL_MASK, dPS.z, 4);
        E[k].x = vB ? dPS.x - ox : 0.f; E[k].y = vB ? dPS.y - oy : 0.f; E[k].z = vB ? dPS.z - oz : 0.f;
        e2 += E[k].x * E[k].x + E[k].y * E[k].y + E[k].z * E[k].z;
      }
      // bone vectors (parent - child): see the parent table above
      {
        const Vec3 prevP = shfl3(f[3].P, (m.lane + 31) & 31);     // slot 3 of the previous lane: joints 4, 8, 12
        const Vec3 P8 = shfl3(f[3].P, m.gb | 1);                  // joint 8
        Vec3 par[4];
        par[0] = prevP; if (m.q == 0) { par[0].x = par[0].y = par[0].z = 0.f; }
        par[1] = m.q == 3 ? P8 : f[0].P;
        par[2] = m.q == 2 ? P8 : f[1].P; if (m.q == 1) { par[2].x = par[2].y = par[2].z = 0.f; }
        par[3] = f[2].P; if (m.q == 0) { par[3].x = par[3].y = par[3].z = 0.f; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          e[k].x = par[k].x - f[k].P.x; e[k].y = par[k].y - f[k].P.y; e[k].z = par[k].z - f[k].P.z;
          len[k] = fast_sqrt(e[k].x * e[k].x + e[k].y * e[k].y + e[k].z * e[k].z);
          lsum += len[k];
        }
      }
      f2 = quad_sum(f2); rep = quad_sum(rep); e2 = quad_sum(e2); lsum = quad_sum(lsum);
      const float L3d = fast_sqrt(f2);
      const float pnorm = fast_sqrt(e2);
      const float imean = fast_rcp(lsum * (1.f / 16.f));
      // ---- phase 2: bone prior (:256-259)
      float h[4], bl = 0.f, hl = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = A.maps.bone_rel[m.j0 - 1 + k] - len[k] * imean;
        bl += t * t;
        h[k] = -2.f * t * cbl;
        hl += h[k] * len[k];
      }
      if (!kFull) {
        bl = quad_sum(bl);
        if (valid && m.q == 0) {
          sums[0] += L3d;
          sums[1] += rep + (fabsf(r.u0x) + fabsf(r.u0y));
          sums[3] += bl;
          if ((rl & 1) == 0) sums[2] += pnorm;
        }
      } else {
        hl = quad_sum(hl);
      }
      const float ge = pnorm > 0.f ? cv / pnorm : 0.f;
      const float g3 = L3d > 0.f ? c3d / L3d : 0.f;
      // ---- phase 3: backward per joint
      Vec3 dP[4], dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // d/dS: reprojection L1 (:247) + pair term
        const float drx = c2d * ((rx[k] > r.ux[k]) ? 1.f : ((rx[k] < r.ux[k]) ? -1.f : 0.f));
        const float dry = c2d * ((ry[k] > r.uy[k]) ? 1.f : ((ry[k] < r.uy[k]) ? -1.f : 0.f));
        Vec3 dS;
        dS.x = drx * izs[k] - ge * E[k].x;
        dS.y = dry * izs[k] - ge * E[k].y;
        dS.z = -(drx * rx[k] + dry * ry[k]) * izs[k] - ge * E[k].z;
        // d/dP2 = R dS - c3d F / L3d   (root centring only feeds the root's own, constant, depth)
        const Vec3 RdS = mat_vec(r.R, dS);
        Vec3 dP2;
        dP2.x = RdS.x - g3 * F[k].x; dP2.y = RdS.y - g3 * F[k].y; dP2.z = RdS.z - g3 * F[k].z;
        const float ddelta2 = mask2[k] * (dP2.x * f[k].qx + dP2.y * f[k].qy + dP2.z);
        g2acc[k][0] += net1[k] ? 0.f : ddelta2;
        g2acc[k][1] += net1[k] ? ddelta2 : 0.f;
        if (kFull) {
          // d/dq: through P2 = (q d2) and the external consumers (flows, pass-2 lifters)
          const float dqx = dP2.x * d2[k] + xqx[k], dqy = dP2.y * d2[k] + xqy[k];
          Vec3 dQ;
          dQ.x = g3 * F[k].x + dqx * f[k].izq;
          dQ.y = g3 * F[k].y + dqy * f[k].izq;
          dQ.z = g3 * F[k].z - (dqx * f[k].qx + dqy * f[k].qy) * f[k].izq;
          // d/dP = R^T dQ + pair + own bone (the children's bones are added below)
          const Vec3 RtdQ = matT_vec(r.R, dQ);
          const float dl = (h[k] - hl * (1.f / 16.f) * imean) * imean;
          const float w = len[k] > 0.f ? dl * fast_rcp(len[k]) : 0.f;
          dv[k].x = w * e[k].x; dv[k].y = w * e[k].y; dv[k].z = w * e[k].z;     // d/d(P_parent); d/d(P_child) = -dv
          dP[k].x = RtdQ.x + ge * E[k].x - dv[k].x;
          dP[k].y = RtdQ.y + ge * E[k].y - dv[k].y;
          dP[k].z = RtdQ.z + ge * E[k].z - dv[k].z;
          da_acc += (dQ.z * f[k].Q.y - dQ.y * f[k].Q.z) + (P2[k].z * RdS.y - P2[k].y * RdS.z);
          dg_acc += (f[k].P.y * RtdQ.z - f[k].P.z * RtdQ.y) + (dS.y * S[k].z - dS.z * S[k].y);
        }
      }
      if (kFull) {
        // bones: the parent joint receives +dv of each child bone
        const Vec3 nx = shfl3(dv[0], (m.lane + 1) & 31);           // bone of the next lane's slot 0 hangs on my slot 3
        const Vec3 b11 = shfl3(dv[2], m.gb | 2), b14 = shfl3(dv[1], m.gb | 3);   // bones of joints 11 and 14 hang on joint 8
        if (m.q != 3) { dP[0].x += dv[1].x; dP[0].y += dv[1].y; dP[0].z += dv[1].z; }
        if (m.q != 1 && m.q != 2) { dP[1].x += dv[2].x; dP[1].y += dv[2].y; dP[1].z += dv[2].z; }
        if (m.q != 0) { dP[2].x += dv[3].x; dP[2].y += dv[3].y; dP[2].z += dv[3].z; }
        if (m.q != 3) { dP[3].x += nx.x; dP[3].y += nx.y; dP[3].z += nx.z; }
        if (m.q == 1) { dP[3].x += b11.x + b14.x; dP[3].y += b11.y + b14.y; dP[3].z += b11.z + b14.z; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float ddelta = f[k].mask * (dP[k].x * r.ux[k] + dP[k].y * r.uy[k] + dP[k].z);     // lift
          g1acc[k][0] += net1[k] ? 0.f : ddelta;
          g1acc[k][1] += net1[k] ? ddelta : 0.f;
        }
      }
    }
    // ---- write head gradients (slot -> column col[j] of every net that feeds joint j in some variant; the root's
    //      columns receive zeros)
    if (valid) {
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        __nv_bfloat16* gT = kT ? (kFull ? A.g1T[net] : A.g2T[net]) : nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat16* g = T.gp[net][m.j0 + k];
          if (g) {
            const __nv_bfloat16 hv = __float2bfloat16_rn(kFull ? g1acc[k][net] : g2acc[k][net]);
            g[r.n * 64] = hv;
            if (kT && gT) gT[static_cast<size_t>(A.maps.col[m.j0 + k]) * A.ldT + A.colT0 + r.n] = hv;
          }
        }
        if (m.q == 0) {
          __nv_bfloat16* g = T.gp[net][0];
          if (g) {
            const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
            g[r.n * 64] = z;
            if (kT && gT) gT[static_cast<size_t>(A.maps.col[0]) * A.ldT + A.colT0 + r.n] = z;
          }
        }
      }
    }
    if (kFull) {
      const float dav = quad_sum(da_acc), dgv = quad_sum(dg_acc);
      if (valid && m.q == 0) {
        A.da[r.n] = dav;
        A.dgamma[r.n] = dgv;
        red_da += dav;
        red_eda += r.eps * dav;
      }
    }
  }
  // ---- block reduction of the scalar sums (lane 0 of every quad holds its rows' partial sums)
#pragma unroll
  for (int k = 0; k < 4; ++k) sums[k] = warp_sum(sums[k]);
  red_da = warp_sum(red_da);
  red_eda = warp_sum(red_eda);
  if (m.lane == 0) {
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

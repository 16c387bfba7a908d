// Geometry + self-supervised losses of the lifter training step, forward and hand-derived backward.
// Reference: train_leg_torso_lifter.py:153-272 (V = 1 pose variant) and train_left_right_lifter.py:150-423
// (V = 2 variants: 'left' / 'right' choice of combine_left_right_pred_1d, utils/helpers.py:40-53).
//
// Mapping: four lanes per row, eight consecutive rows (four row pairs) per warp.  Lane q of a row's quad owns joints
// q+1, q+5, q+9, q+13 and the bones whose children they are ("slots" 0-3): the per-joint math of a slot is straight-line code over
// registers, the row-uniform work (rotation, reductions) is shared by 4 lanes instead of being replicated in 16, per-row
// reductions are two xor-shuffles, the pairwise deformation loss (:250-254) couples rows 2k and 2k+1 = lanes l and l^4.
// The root joint needs no slot: after root-centring (:188-192) its lifted, rotated, re-lifted and back-rotated positions
// are identically zero, its depth offset is forced to 0 (:183) so it receives no gradient, and its only loss contribution
// is the constant |u_root| of the reprojection term.  Nothing but the network outputs is read: P, R, Q, q are recomputed
// from (u, depth heads, angle heads, eps_x, u_y, stats).
#pragma once
#include "devdefs.cuh"
#include "f2.cuh"

namespace links {

constexpr int kGeomWarps = 4;
constexpr int kGeomRows = 8;    // rows per warp and grid-stride iteration of the loss / backward kernels (four lanes per row)
constexpr int kGeomMaxRows = (1 << 25) - 8;   // element offsets (row * 64 + column) are formed in 32 bits
constexpr int kJ = 17;

// ---- staging plan (host-built, see geom_plan): which slices of which tensors a warp copies into shared memory for the 8
// rows of one iteration, and where the logical tensors sit inside that staging buffer.
constexpr int kGeomMaxRegions = 14;
constexpr int kGeomSp = 144;        // bytes between the rows of a strided slice in shared memory: 128 + 16, so that the 8 rows
                                    // of a warp start in 8 different bank groups
constexpr int kGeomMaxCopyIters = 32, kGeomMaxOutIters = 12;   // 16-byte copies per lane and iteration (input / output lists)
struct GeomRegion {
  char* g;              // global address of row 0 of the slice (16-byte aligned)
  int pitch;            // bytes between rows in global memory
  int cpr;              // strided: 16-byte chunks staged per row; contiguous (the 8-row block is one range): 0
  int soff;             // byte offset of the slice in the warp's staging buffer
  int spitch;           // strided: bytes between rows in the staging buffer (kGeomSp)
};
struct GeomStage {
  int rows;                                   // rows a warp stages per iteration: 8 (four lanes per row) or 32 (one lane per row)
  int n_in, n_out;
  GeomRegion in[kGeomMaxRegions], out[4];
  int in_bytes, out_bytes;                    // per warp: one input buffer (there are two), the output buffer
  int in_chunks, out_chunks;                  // 16-byte copies per warp iteration
  // byte offsets (row 0) of the logical tensors inside the input / output staging buffers
  int u_off, eps_off, uy_off;                 // contiguous: row pitch 136 / 4 / 4
  int head_off[2], ang_off[2], head2_off[2], dlift_off[2];     // strided (row pitch kGeomSp)
  int dflow_off[2];                           // contiguous: row pitch 8 n_joints
  int qpart_off[2];                           // output, contiguous: row pitch 8 n_joints
  int g_off[2];                               // output, strided
  int zero_off;                               // input buffers: 16 bytes that are never written (= 0); used with pitch 0
  int trash_off;                              // output buffer: bytes nobody copies out (strided form, or pitch 0)
};

struct GeomArgs {
  LinksGeomMaps maps;
  GeomStage st;
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

// ---------------------------------------------------------------------------------------------------------
// Staging.  The tensors of this step are row-major with 56 ... 136-byte rows of which a lane needs single words picked by
// the joint maps: read or written directly that is one 4-byte (2-byte) access per word -- 8 rows x 4 lanes of a warp hit up
// to 32 different sectors per instruction and the kernels were bound by the load/store unit, not by HBM (ncu, round 2).
// Instead a warp moves the 8 rows of its iteration as whole 16-byte chunks (coalesced cp.async into shared memory, one
// iteration ahead; outputs back with 16-byte stores) and the joint-wise gathers / scatters run against shared memory.
// ---------------------------------------------------------------------------------------------------------
struct GeomCopy {        // one 16-byte copy of the per-iteration list
  char* g0;              // global address for iteration 0
  int stride;            // bytes per iteration (8 rows)
  int soff;              // offset in the staging buffer
};
// Per-block lookup tables (shared memory, built once per block): staging-buffer byte offsets (row 0) of what a lane reads
// or writes for joint j in variant v; the lane adds its row offset (rl * kGeomSp for strided slices, rl * pitch for contiguous
// ones).  Entries without a source point at zero bytes (pitch 0); entries without a destination are -1.
struct GeomTabs {
  int hs[2][kJ];        // pass-1 depth-head output (strided)
  int h2s[2][kJ];       // pass-2 depth-head output (strided)
  int dls[2][kJ];       // d/d(projected x) from the pass-2 lifter input gradient (strided); y sits njy4 bytes further
  int dfs[2][kJ];       // ... from the part flow (contiguous, pitch dfp); y sits njy4 bytes further
  int dfp[2][kJ];       // row pitch (bytes) of the joint's part: flow gradient rows and projected-part output rows
  int njy4[2][kJ];
  int qps[2][kJ];       // projected x in the output buffer (contiguous, pitch dfp); scratch if the joint has no part
  int gs[2][kJ];        // [net][j]: head-gradient element in the output buffer (strided); scratch if net never feeds j
};
#ifndef LINKS_HOSTSIM
__device__ __forceinline__ void geom_cp16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(sdst))), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void geom_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending> __device__ __forceinline__ void geom_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }
#else
__device__ __forceinline__ void geom_cp16(void* sdst, const void* gsrc) { memcpy(sdst, gsrc, 16); }
__device__ __forceinline__ void geom_cp_commit() {}
template <int kPending> __device__ __forceinline__ void geom_cp_wait() {}
#endif

// chunk c of a region -> (row, byte offset in the row) for strided regions; contiguous regions are one 8-row range
__device__ __forceinline__ void build_copy_list(const GeomRegion* R, int n, int rows, GeomCopy* list) {
  int base = 0;
  for (int i = 0; i < n; ++i) {
    const GeomRegion r = R[i];
    const int nchunk = r.cpr ? rows * r.cpr : (rows * r.pitch) / 16;
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
      GeomCopy e;
      if (r.cpr) {
        const int row = c / r.cpr, cc = c - row * r.cpr;
        e.g0 = r.g + static_cast<size_t>(row) * r.pitch + cc * 16;
        e.soff = r.soff + row * r.spitch + cc * 16;
      } else {
        e.g0 = r.g + c * 16;
        e.soff = r.soff + c * 16;
      }
      e.stride = rows * r.pitch;
      list[base + c] = e;
    }
    base += nchunk;
  }
}

template <int V>
__device__ __forceinline__ void build_tabs(const GeomArgs& A, GeomTabs& T, bool full) {
  const GeomStage& S = A.st;
  for (int i = threadIdx.x; i < 2 * kJ; i += blockDim.x) {
    const int v = i / kJ, j = i - v * kJ;
    // gradient destinations are per NET (first index), everything else per VARIANT
    {
      bool fed = false;
#pragma unroll
      for (int w = 0; w < V; ++w) fed = fed || A.maps.src_net[w][j] == v;
      __nv_bfloat16* g = full ? A.g1[v] : A.g2[v];
      T.gs[v][j] = (fed && g) ? S.g_off[v] + 2 * A.maps.col[j] : S.trash_off;
    }
    if (v >= V) continue;
    const int net = A.maps.src_net[v][j], col = A.maps.col[j];
    T.hs[v][j] = S.head_off[net] + 4 * col;
    T.h2s[v][j] = A.head2[net] ? S.head2_off[net] + 4 * col : S.zero_off;
    const int pn = A.maps.part_net[v][j];
    const bool has = pn >= 0;
    const int pi = has ? pn : 0;
    const int nj = A.maps.n_joints[pi], idx = A.maps.part_idx[v][j];
    const bool ext = has && A.dflow[pi] && A.dlift[pi];
    T.dfs[v][j] = ext ? S.dflow_off[pi] + 4 * idx : S.zero_off;
    T.dls[v][j] = ext ? S.dlift_off[pi] + 4 * idx : S.zero_off;
    T.dfp[v][j] = (ext || (has && A.qpart[pi])) ? 8 * nj : 0;
    T.njy4[v][j] = ext || (has && A.qpart[pi]) ? 4 * nj : 0;
    T.qps[v][j] = (has && A.qpart[pi]) ? S.qpart_off[pi] + 4 * idx : S.trash_off;
  }
}

// Read-only table words.  The tables are written once before the row loop; stores into the staging buffers (char*) would
// otherwise force a reload after every store.  A non-volatile asm without a memory clobber is a pure function of its
// address: the compiler may hoist it out of the loop or re-issue it, as it sees fit.  `tok` (GeomSmem::tok, always 0) is
// produced by a volatile asm AFTER the block barrier that publishes the tables and is an (unused) input of every table load, so
// none of them can be scheduled above that barrier.
__device__ __forceinline__ int tab_ld(const int* p, uint32_t tok) {
#ifndef LINKS_HOSTSIM
  int v;
  asm("ld.shared.b32 %0, [%1];   // ordered behind %2" : "=r"(v) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(tok));
  return v;
#else
  (void)tok;
  return *p;
#endif
}
// entry `field`[v][j0 + k] of the tables / root entry `field`[v][0] / map entry maps.`field`[v][j0 + k]
#define GEOM_TAB(field, v, k) tab_ld(&T.field[v][m.j0 + (k) * m.js], M.tok)
#define GEOM_TAB0(field, v) tab_ld(&T.field[v][0], M.tok)
#define GEOM_MAP(field, v, k) tab_ld(&A.maps.field[v][m.j0 + (k) * m.js], M.tok)

// Shared-memory frame of a block: [GeomArgs][GeomTabs][input copy list][output copy list][per warp: in0 | in1 | out]
struct GeomSmem {
  GeomArgs* A;
  GeomTabs* T;
  GeomCopy *cin, *cout;
  char* wbuf;           // this warp's staging buffers
  uint32_t tok;         // 0; orders the pure table loads behind the set-up barrier (tab_ld)
};
__device__ __forceinline__ size_t geom_rup16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }
__host__ __device__ inline size_t geom_smem_bytes(const GeomStage& S) {
  size_t b = ((sizeof(GeomArgs) + 15) & ~size_t(15)) + ((sizeof(GeomTabs) + 15) & ~size_t(15));
  b += static_cast<size_t>(S.in_chunks + S.out_chunks) * sizeof(GeomCopy);
  b += static_cast<size_t>(kGeomWarps) * (2 * S.in_bytes + S.out_bytes);
  return b;
}
template <int V>
__device__ __forceinline__ void geom_setup(const GeomArgs& Ap, bool full, GeomSmem& M) {
  LINKS_DYN_SMEM(char, dyn);
  M.A = reinterpret_cast<GeomArgs*>(dyn);
  size_t off = geom_rup16(sizeof(GeomArgs));
  M.T = reinterpret_cast<GeomTabs*>(dyn + off);
  off += geom_rup16(sizeof(GeomTabs));
  stage_args(M.A, Ap);                                   // ends with a block barrier
  const GeomStage& S = M.A->st;
  M.cin = reinterpret_cast<GeomCopy*>(dyn + off);
  off += static_cast<size_t>(S.in_chunks) * sizeof(GeomCopy);
  M.cout = reinterpret_cast<GeomCopy*>(dyn + off);
  off += static_cast<size_t>(S.out_chunks) * sizeof(GeomCopy);
  const int per_warp = 2 * S.in_bytes + S.out_bytes;
  // zero every staging buffer once: the padding of strided slices and the columns of the gradient rows that no joint
  // writes stay zero for the whole kernel
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(dyn + off);
    for (int i = threadIdx.x; i < kGeomWarps * per_warp / 4; i += blockDim.x) z[i] = 0u;
  }
  M.wbuf = dyn + off + static_cast<size_t>(threadIdx.x >> 5) * per_warp;
  build_tabs<V>(*M.A, *M.T, full);
  build_copy_list(S.in, S.n_in, S.rows, M.cin);
  build_copy_list(S.out, S.n_out, S.rows, M.cout);
  __syncthreads();
#ifndef LINKS_HOSTSIM
  asm volatile("mov.u32 %0, 0;" : "=r"(M.tok) : : "memory");
#else
  M.tok = 0;
#endif
}

// Start the copies of the rows [row0, row0 + 8) into `buf`.  Full blocks: the prepared 16-byte list.  The ragged last
// block: word by word, rows past the end are not touched.
__device__ __forceinline__ void geom_stage_in(const GeomSmem& M, int it, int N, char* buf) {
  const GeomStage& S = M.A->st;
  const int lane = threadIdx.x & 31;
  const int row0 = it * S.rows;
  if (row0 + S.rows <= N) {
    const int nc = S.in_chunks;
    for (int c0 = lane; c0 < nc; c0 += 32 * 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {                         // four independent copies in flight per lane, predicated
        const int c = c0 + 32 * i;
        if (c < nc) {
          const GeomCopy e = M.cin[c];
          geom_cp16(buf + e.soff, e.g0 + static_cast<size_t>(it) * e.stride);
        }
      }
    }
  } else if (row0 < N) {
    const int rows = N - row0;
    for (int i = 0; i < S.n_in; ++i) {
      const GeomRegion r = S.in[i];
      const char* g = r.g + static_cast<size_t>(row0) * r.pitch;
      if (r.cpr) {
        for (int w = lane; w < rows * r.cpr * 4; w += 32) {
          const int row = w / (r.cpr * 4), ww = w - row * (r.cpr * 4);
          *reinterpret_cast<uint32_t*>(buf + r.soff + row * r.spitch + 4 * ww) =
              *reinterpret_cast<const uint32_t*>(g + static_cast<size_t>(row) * r.pitch + 4 * ww);
        }
      } else {
        for (int w = lane; w < rows * r.pitch / 4; w += 32)
          *reinterpret_cast<uint32_t*>(buf + r.soff + 4 * w) = *reinterpret_cast<const uint32_t*>(g + 4 * w);
      }
    }
  }
  geom_cp_commit();
}
// Write the output staging buffer of rows [row0, row0 + 8) back (16-byte stores; ragged last block word by word).
__device__ __forceinline__ void geom_stage_out(const GeomSmem& M, int it, int N, const char* buf) {
  const GeomStage& S = M.A->st;
  const int lane = threadIdx.x & 31;
  const int row0 = it * S.rows;
  if (row0 + S.rows <= N) {
    const int nc = S.out_chunks;
    for (int c0 = lane; c0 < nc; c0 += 32 * 3) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = c0 + 32 * i;
        if (c < nc) {
          const GeomCopy e = M.cout[c];
          *reinterpret_cast<uint4*>(e.g0 + static_cast<size_t>(it) * e.stride) = *reinterpret_cast<const uint4*>(buf + e.soff);
        }
      }
    }
  } else if (row0 < N) {
    const int rows = N - row0;
    for (int i = 0; i < S.n_out; ++i) {
      const GeomRegion r = S.out[i];
      char* g = r.g + static_cast<size_t>(row0) * r.pitch;
      if (r.cpr) {
        for (int w = lane; w < rows * r.cpr * 4; w += 32) {
          const int row = w / (r.cpr * 4), ww = w - row * (r.cpr * 4);
          *reinterpret_cast<uint32_t*>(g + static_cast<size_t>(row) * r.pitch + 4 * ww) =
              *reinterpret_cast<const uint32_t*>(buf + r.soff + row * r.spitch + 4 * ww);
        }
      } else {
        for (int w = lane; w < rows * r.pitch / 4; w += 32)
          *reinterpret_cast<uint32_t*>(g + 4 * w) = *reinterpret_cast<const uint32_t*>(buf + r.soff + 4 * w);
      }
    }
  }
}
__device__ __forceinline__ float lds_f(const char* base, int off) { return *reinterpret_cast<const float*>(base + off); }
// Read-only table word.  The tables are written once before the row loop; stores into the staging buffers (char*) would
// otherwise force a reload after every store.  A non-volatile asm without a memory clobber is a pure function of its
// address: the compiler may hoist it out of the loop or re-issue it under register pressure, as it sees fit.
// bone table (utils/helpers.py:140-141): bone b joins parent kBoneParent[b] and child b+1 (4 bits per entry)
constexpr unsigned long long kBoneParent = 0xfe8cb89870540210ull;
__host__ __device__ constexpr int bone_parent(int child) { return static_cast<int>((kBoneParent >> (4 * (child - 1))) & 15ull); }
// The quad mapping below hard-wires where a slot finds the parent joint of its bone; tie it to the table.  Lane q of a
// row's quad owns joints q+1, q+5, q+9, q+13 (slots 0-3).  For most bones the parent is the SAME slot of the previous lane
// (for lane 0: the previous slot of lane 3), so one rotating exchange per slot serves them:
//   slot 0 (joints 1, 2, 3, 4)    : root, joint 1, joint 2, root
//   slot 1 (joints 5, 6, 7, 8)    : joint 4, joint 5, root, joint 7
//   slot 2 (joints 9, 10, 11, 12) : joint 8, joint 9, joint 8 (lane 3, slot 1: broadcast), joint 11
//   slot 3 (joints 13, 14, 15, 16): joint 12, joint 8 (broadcast), joint 14, joint 15
static_assert(bone_parent(1) == 0 && bone_parent(2) == 1 && bone_parent(3) == 2 && bone_parent(4) == 0, "slot 0 parents");
static_assert(bone_parent(5) == 4 && bone_parent(6) == 5 && bone_parent(7) == 0 && bone_parent(8) == 7, "slot 1 parents");
static_assert(bone_parent(9) == 8 && bone_parent(10) == 9 && bone_parent(11) == 8 && bone_parent(12) == 11, "slot 2 parents");
static_assert(bone_parent(13) == 12 && bone_parent(14) == 8 && bone_parent(15) == 14 && bone_parent(16) == 15, "slot 3 parents");

// Lane-constant indexing state.
struct Quad {
  int lane, q, gb, j0, js;   // q = lane & 3; gb = first lane of this row's quad; slot k of the lane owns joint j0 + k * js
};
// Joints q+1, q+5, q+9, q+13: the four lanes of a row read four CONSECUTIVE words of a staged row for every slot, which
// together with the 4-word bank shift per staged row makes the gathers of a warp conflict-free (a lane owning four
// consecutive joints costs 4-way bank conflicts on every gather: forward kernel 0.60 -> 0.73 of HBM peak).
__device__ __forceinline__ void quad_init(Quad& m) {
  m.lane = threadIdx.x & 31;
  m.q = m.lane & 3;
  m.gb = m.lane & ~3;
  m.j0 = m.q + 1;
  m.js = 4;
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
// `cur`: the warp's staged input rows; rl = row slot of this lane's quad; st0 / st1 = (mean, std) of the elevation
__device__ __forceinline__ void load_row(const GeomStage& S, const Quad& m, const char* cur, int rl, int n, float st0, float st1,
                                         RowIn& r) {
  r.n = n;
  const char* u = cur + S.u_off + rl * (2 * kJ * 4);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r.ux[k] = lds_f(u, 4 * (m.j0 + k * m.js));
    r.uy[k] = lds_f(u, 4 * (kJ + m.j0 + k * m.js));
  }
  r.u0x = lds_f(u, 0);
  r.u0y = lds_f(u, 4 * kJ);
  const float ang0 = lds_f(cur, S.ang_off[0] + rl * kGeomSp), ang1 = lds_f(cur, S.ang_off[1] + rl * kGeomSp);
  r.eps = lds_f(cur, S.eps_off + 4 * rl);
  const float uyaw = lds_f(cur, S.uy_off + 4 * rl);
  const float gamma = 0.5f * (ang0 + ang1);
  const float a = -st0 + st1 * r.eps;
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
template <int V, bool kQ = false>      // kQ: also store the full projected poses qfull[v] (a debugging output of the step)
__global__ void __launch_bounds__(kGeomWarps * 32) geom_forward_kernel(const GeomArgs Ap) {
  GeomSmem M;
  geom_setup<V>(Ap, false, M);
  const GeomArgs& A = *M.A;
  const GeomTabs& T = *M.T;
  const GeomStage& S = A.st;
  Quad m;
  quad_init(m);
  const int warp = threadIdx.x >> 5, rl = m.lane >> 2;
  const int rs = rl * kGeomSp;
  const float D = A.maps.depth;
  const float d0 = D < 1.0f ? 1.0f : D;                   // root depth: offset forced to 0 (:183), then clamped
  const float st0 = A.stats[0], st1 = A.stats[1];
  const int n_iters = (A.N + kGeomRows - 1) / kGeomRows;
  const int stride = gridDim.x * kGeomWarps;
  char* const in0 = M.wbuf;
  char* const in1 = M.wbuf + S.in_bytes;
  char* const outb = M.wbuf + 2 * S.in_bytes;
  // The loop bounds depend on blockIdx only (block-uniform trip count): the compiler can then prove that the warp is
  // converged at every shuffle and emits plain SHFLs; a warp past the end works on stale rows with every write masked.
  int buf = 0;
  geom_stage_in(M, blockIdx.x * kGeomWarps + warp, A.N, in0);
  for (int base = blockIdx.x * kGeomWarps; base < n_iters; base += stride, buf ^= 1) {
    const int it = base + warp;
    const char* cur = buf ? in1 : in0;
    geom_stage_in(M, it + stride, A.N, buf ? in0 : in1);     // next iteration's rows (an empty group past the end)
    geom_cp_wait<1>();
    __syncwarp();
    const int n_raw = kGeomRows * it + rl;
    const bool valid = n_raw < A.N;                      // uniform over the quad
    RowIn r;
    load_row(S, m, cur, rl, n_raw, st0, st1, r);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float delta[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) delta[k] = lds_f(cur, GEOM_TAB(hs, v, k) + rs);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = m.j0 + k * m.js;
        JointFwd s;
        joint_forward(D, d0, delta[k], r.ux[k], r.uy[k], r.u0x, r.u0y, r.R, s);
        if (kQ && valid && A.qfull[v]) {
          float* qf = A.qfull[v] + r.n * 34;
          qf[j] = s.qx;
          qf[kJ + j] = s.qy;
        }
        char* dst = outb + GEOM_TAB(qps, v, k) + rl * GEOM_TAB(dfp, v, k);
        *reinterpret_cast<float*>(dst) = s.qx;
        *reinterpret_cast<float*>(dst + GEOM_TAB(njy4, v, k)) = s.qy;
      }
      if (m.q == 0) {                                      // root: projects to (0, 0)
        if (kQ && valid && A.qfull[v]) {
          float* qf = A.qfull[v] + r.n * 34;
          qf[0] = 0.f; qf[kJ] = 0.f;
        }
        char* dst = outb + GEOM_TAB0(qps, v) + rl * GEOM_TAB0(dfp, v);
        *reinterpret_cast<float*>(dst) = 0.f;
        *reinterpret_cast<float*>(dst + GEOM_TAB0(njy4, v)) = 0.f;
      }
    }
    __syncwarp();
    geom_stage_out(M, it, A.N, outb);
    __syncwarp();
  }
  geom_cp_wait<0>();
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
__global__ void __launch_bounds__(kGeomWarps * 32, kFull ? 3 : 4) geom_lossgrad_kernel(const GeomArgs Ap) {
  __shared__ float s_part[kGeomWarps][6];
  GeomSmem M;
  geom_setup<V>(Ap, kFull, M);
  const GeomArgs& A = *M.A;
  const GeomTabs& T = *M.T;
  const GeomStage& S = A.st;
  Quad m;
  quad_init(m);
  const int warp = threadIdx.x >> 5, rl = m.lane >> 2;
  const int rs = rl * kGeomSp;
  const float st0 = A.stats[0], st1 = A.stats[1];
  char* const in0 = M.wbuf;
  char* const in1 = M.wbuf + S.in_bytes;
  char* const outb = M.wbuf + 2 * S.in_bytes;
  const float invN = 1.f / static_cast<float>(A.N);
  const int npairs = A.N / 2;
  const float c3d = A.maps.w_3d * invN, c2d = A.maps.w_2d * invN, cbl = A.maps.w_bl * invN;
  const float cv = npairs > 0 ? A.maps.w_vel / static_cast<float>(npairs) : 0.f;
  const float D = A.maps.depth;
  const float d0 = D < 1.0f ? 1.0f : D;
  const F2 D2 = f2_splat(D), nd0 = f2_splat(-d0), m1 = f2_splat(-1.f);
  float sums[4] = {0.f, 0.f, 0.f, 0.f};   // L3d, rep, pair, bl (raw sums; lane 0 of each quad accumulates its rows)
  float red_da = 0.f, red_eda = 0.f;

  const int n_iters = (A.N + kGeomRows - 1) / kGeomRows;
  const int stride = gridDim.x * kGeomWarps;
  // block-uniform trip count (see geom_forward_kernel): shuffles sit in provably convergent code
  int buf = 0;
  geom_stage_in(M, blockIdx.x * kGeomWarps + warp, A.N, in0);
  for (int base = blockIdx.x * kGeomWarps; base < n_iters; base += stride, buf ^= 1) {
    const int it = base + warp;
    const char* cur = buf ? in1 : in0;
    geom_stage_in(M, it + stride, A.N, buf ? in0 : in1);     // next iteration's rows (an empty group past the end)
    geom_cp_wait<1>();
    __syncwarp();
    const int n_raw = kGeomRows * it + rl;
    const bool valid = n_raw < A.N;                               // uniform over the quad
    const bool vB = (n_raw | 1) < A.N;                            // uniform over the pair's 8 lanes: the pair is complete
    RowIn r;
    load_row(S, m, cur, rl, n_raw, st0, st1, r);
    float g1acc[4][2], g2acc[4][2];                               // [slot][net] d/d(head) of the slot's column
#pragma unroll
    for (int k = 0; k < 4; ++k) { g1acc[k][0] = g1acc[k][1] = g2acc[k][0] = g2acc[k][1] = 0.f; }
    float da_acc = 0.f, dg_acc = 0.f;                             // lane-partial d/da, d/dgamma
    F2 dap = f2_splat(0.f), dan = f2_splat(0.f), dgp = f2_splat(0.f), dgn = f2_splat(0.f);   // ... as positive / negative parts
    F2 R2[9];                                                     // the rotation, both halves
#pragma unroll
    for (int i = 0; i < 9; ++i) R2[i] = f2_splat(r.R[i]);
    const F2 nx0 = f2_splat(-r.u0x * d0), ny0 = f2_splat(-r.u0y * d0);
    const F2 vb2 = f2_splat(vB ? 1.f : 0.f);

#pragma unroll
    for (int v = 0; v < V; ++v) {
      // ---- everything this (row, variant) reads, issued together
      float delta[4], delta2[4], xqx[4], xqy[4];
      bool net1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        net1[k] = GEOM_MAP(src_net, v, k) != 0;
        delta[k] = lds_f(cur, GEOM_TAB(hs, v, k) + rs);
        delta2[k] = lds_f(cur, GEOM_TAB(h2s, v, k) + rs);
        xqx[k] = 0.f; xqy[k] = 0.f;
        if (kFull) {
          const char* df = cur + GEOM_TAB(dfs, v, k) + rl * GEOM_TAB(dfp, v, k);
          const char* dl = cur + GEOM_TAB(dls, v, k) + rs;
          const int ny = GEOM_TAB(njy4, v, k);
          xqx[k] = lds_f(df, 0) + lds_f(dl, 0);
          xqy[k] = lds_f(df, ny) + lds_f(dl, ny);
        }
      }
      // ---- phase 1.  Two slots per packed f32x2 value (pair p = slots 2p, 2p + 1): the per-joint arithmetic issues as
      //      FFMA2 / FMUL2 / FADD2, half the instructions of the scalar form; selects, MUFU and shuffles stay per element.
      F2 mask[2], Px[2], Py[2], Pz[2], Qx[2], Qy[2], Qz[2], izq[2], qx[2], qy[2];
      F2 mask2[2], d2[2], P2x[2], P2y[2], P2z[2], Fx[2], Fy[2], Fz[2], Sx[2], Sy[2], Sz[2], izs[2], rx[2], ry[2];
      F2 dxr[2], dyr[2];                                           // reprojection residuals rx - ux, ry - uy
      F2 Ex[2], Ey[2], Ez[2], ex[2], ey[2], ez[2], len[2], ux2[2], uy2[2];
      F2 f2a = f2_splat(0.f), e2a = f2_splat(0.f);
      float rep = 0.f, lsum = 0.f;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        ux2[p] = f2_make(r.ux[2 * p], r.ux[2 * p + 1]);
        uy2[p] = f2_make(r.uy[2 * p], r.uy[2 * p + 1]);
        // lift (:183-192), rotate (:195), project (:198-199)
        F2 d = f2_add(f2_make(delta[2 * p], delta[2 * p + 1]), D2);
        mask[p] = f2_make(d.x < 1.0f ? 0.f : 1.f, d.y < 1.0f ? 0.f : 1.f);
        d = f2_make(d.x < 1.0f ? 1.0f : d.x, d.y < 1.0f ? 1.0f : d.y);
        Px[p] = f2_fma(ux2[p], d, nx0);
        Py[p] = f2_fma(uy2[p], d, ny0);
        Pz[p] = f2_add(d, nd0);
        f2_matvec(R2, Px[p], Py[p], Pz[p], Qx[p], Qy[p], Qz[p]);
        {
          const F2 t = f2_add(Qz[p], D2);
          izq[p] = f2_make(fast_rcp(t.x), fast_rcp(t.y));
        }
        qx[p] = f2_mul(Qx[p], izq[p]);
        qy[p] = f2_mul(Qy[p], izq[p]);
        // re-lift (:228-238); the root projects to (0, 0): its re-lifted position is (0, 0, d0)
        F2 dd = f2_add(f2_make(delta2[2 * p], delta2[2 * p + 1]), D2);
        mask2[p] = f2_make(dd.x < 1.0f ? 0.f : 1.f, dd.y < 1.0f ? 0.f : 1.f);
        dd = f2_make(dd.x < 1.0f ? 1.0f : dd.x, dd.y < 1.0f ? 1.0f : dd.y);
        d2[p] = dd;
        P2x[p] = f2_mul(qx[p], dd);
        P2y[p] = f2_mul(qy[p], dd);
        P2z[p] = f2_add(dd, nd0);
        Fx[p] = f2_fma(P2x[p], m1, Qx[p]);
        Fy[p] = f2_fma(P2y[p], m1, Qy[p]);
        Fz[p] = f2_fma(P2z[p], m1, Qz[p]);
        f2a = f2_fma(Fx[p], Fx[p], f2a); f2a = f2_fma(Fy[p], Fy[p], f2a); f2a = f2_fma(Fz[p], Fz[p], f2a);
        // rotate back and re-project (:242-247)
        f2_matTvec(R2, P2x[p], P2y[p], P2z[p], Sx[p], Sy[p], Sz[p]);
        {
          const F2 t = f2_add(Sz[p], D2);
          izs[p] = f2_make(fast_rcp(t.x), fast_rcp(t.y));
        }
        rx[p] = f2_mul(Sx[p], izs[p]);
        ry[p] = f2_mul(Sy[p], izs[p]);
        dxr[p] = f2_fma(ux2[p], m1, rx[p]);
        dyr[p] = f2_fma(uy2[p], m1, ry[p]);
        rep += (fabsf(dxr[p].x) + fabsf(dxr[p].y)) + (fabsf(dyr[p].x) + fabsf(dyr[p].y));
        // pairwise deformation (:250-254): E = (P - P') - (S - S'), ' = the other row of the pair
        const F2 tx = f2_fma(Sx[p], m1, Px[p]), ty = f2_fma(Sy[p], m1, Py[p]), tz = f2_fma(Sz[p], m1, Pz[p]);
        const F2 ox = f2_make(__shfl_xor_sync(LINKS_FULL_MASK, tx.x, 4), __shfl_xor_sync(LINKS_FULL_MASK, tx.y, 4));
        const F2 oy = f2_make(__shfl_xor_sync(LINKS_FULL_MASK, ty.x, 4), __shfl_xor_sync(LINKS_FULL_MASK, ty.y, 4));
        const F2 oz = f2_make(__shfl_xor_sync(LINKS_FULL_MASK, tz.x, 4), __shfl_xor_sync(LINKS_FULL_MASK, tz.y, 4));
        Ex[p] = f2_mul(f2_fma(ox, m1, tx), vb2);
        Ey[p] = f2_mul(f2_fma(oy, m1, ty), vb2);
        Ez[p] = f2_mul(f2_fma(oz, m1, tz), vb2);
        e2a = f2_fma(Ex[p], Ex[p], e2a); e2a = f2_fma(Ey[p], Ey[p], e2a); e2a = f2_fma(Ez[p], Ez[p], e2a);
      }
      // bone vectors (parent - child): see the parent table above
      {
        Vec3 Ps[4];
        Ps[0].x = Px[0].x; Ps[0].y = Py[0].x; Ps[0].z = Pz[0].x;
        Ps[1].x = Px[0].y; Ps[1].y = Py[0].y; Ps[1].z = Pz[0].y;
        Ps[2].x = Px[1].x; Ps[2].y = Py[1].x; Ps[2].z = Pz[1].x;
        Ps[3].x = Px[1].y; Ps[3].y = Py[1].y; Ps[3].z = Pz[1].y;
        const int src = m.gb | ((m.q + 3) & 3);                  // previous lane of the quad (lane 0 <- lane 3)
        Vec3 par[4];
        par[0] = shfl3(Ps[0], src);
#pragma unroll
        for (int k = 1; k < 4; ++k) par[k] = shfl3(m.q == 3 ? Ps[k - 1] : Ps[k], src);   // lane 3 hands lane 0 its previous slot
        const Vec3 P8 = shfl3(Ps[1], m.gb | 3);                  // joint 8
        if (m.q == 0 || m.q == 3) { par[0].x = par[0].y = par[0].z = 0.f; }              // joints 1, 4 hang on the root
        if (m.q == 2) { par[1].x = par[1].y = par[1].z = 0.f; par[2] = P8; }             // joint 7 <- root, joint 11 <- joint 8
        if (m.q == 1) par[3] = P8;                                                        // joint 14 <- joint 8
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          ex[p] = f2_fma(Px[p], m1, f2_make(par[2 * p].x, par[2 * p + 1].x));
          ey[p] = f2_fma(Py[p], m1, f2_make(par[2 * p].y, par[2 * p + 1].y));
          ez[p] = f2_fma(Pz[p], m1, f2_make(par[2 * p].z, par[2 * p + 1].z));
          const F2 l2 = f2_fma(ez[p], ez[p], f2_fma(ey[p], ey[p], f2_mul(ex[p], ex[p])));
          len[p] = f2_make(fast_sqrt(l2.x), fast_sqrt(l2.y));
          lsum += len[p].x + len[p].y;
        }
      }
      const float f2 = quad_sum(f2a.x + f2a.y), e2 = quad_sum(e2a.x + e2a.y);
      rep = quad_sum(rep); lsum = quad_sum(lsum);
      const float L3d = fast_sqrt(f2);
      const float pnorm = fast_sqrt(e2);
      const float imean = fast_rcp(lsum * (1.f / 16.f));
      // ---- phase 2: bone prior (:256-259)
      F2 h[2];
      F2 bla = f2_splat(0.f), hla = f2_splat(0.f);
      {
        const F2 nim = f2_splat(-imean), m2c = f2_splat(-2.f * cbl);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const F2 crel = f2_make(__uint_as_float(static_cast<unsigned>(tab_ld(reinterpret_cast<const int*>(&A.maps.bone_rel[m.j0 - 1 + (2 * p) * m.js]), M.tok))),
                                  __uint_as_float(static_cast<unsigned>(tab_ld(reinterpret_cast<const int*>(&A.maps.bone_rel[m.j0 - 1 + (2 * p + 1) * m.js]), M.tok))));
          const F2 t = f2_fma(len[p], nim, crel);
          bla = f2_fma(t, t, bla);
          h[p] = f2_mul(t, m2c);
          hla = f2_fma(h[p], len[p], hla);
        }
      }
      float hl = 0.f;
      if (!kFull) {
        const float bl = quad_sum(bla.x + bla.y);
        if (valid && m.q == 0) {
          sums[0] += L3d;
          sums[1] += rep + (fabsf(r.u0x) + fabsf(r.u0y));
          sums[3] += bl;
          if ((rl & 1) == 0) sums[2] += pnorm;
        }
      } else {
        hl = quad_sum(hla.x + hla.y);
      }
      const float ge = pnorm > 0.f ? cv / pnorm : 0.f;
      const float g3 = L3d > 0.f ? c3d / L3d : 0.f;
      const F2 ge2 = f2_splat(ge), nge2 = f2_splat(-ge), g32 = f2_splat(g3), ng32 = f2_splat(-g3);
      // ---- phase 3: backward per joint pair
      F2 dPx[2], dPy[2], dPz[2], dvx[2], dvy[2], dvz[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        // d/dS: reprojection L1 (:247) + pair term
        const F2 drx = f2_make(dxr[p].x > 0.f ? c2d : (dxr[p].x < 0.f ? -c2d : 0.f), dxr[p].y > 0.f ? c2d : (dxr[p].y < 0.f ? -c2d : 0.f));
        const F2 dry = f2_make(dyr[p].x > 0.f ? c2d : (dyr[p].x < 0.f ? -c2d : 0.f), dyr[p].y > 0.f ? c2d : (dyr[p].y < 0.f ? -c2d : 0.f));
        const F2 dSx = f2_fma(Ex[p], nge2, f2_mul(drx, izs[p]));
        const F2 dSy = f2_fma(Ey[p], nge2, f2_mul(dry, izs[p]));
        const F2 sxy = f2_fma(dry, ry[p], f2_mul(drx, rx[p]));
        const F2 dSz = f2_mul(f2_fma(Ez[p], ge2, f2_mul(sxy, izs[p])), m1);
        // d/dP2 = R dS - c3d F / L3d   (root centring only feeds the root's own, constant, depth)
        F2 Rx, Ry, Rz;
        f2_matvec(R2, dSx, dSy, dSz, Rx, Ry, Rz);
        const F2 dP2x = f2_fma(Fx[p], ng32, Rx), dP2y = f2_fma(Fy[p], ng32, Ry), dP2z = f2_fma(Fz[p], ng32, Rz);
        const F2 dd2 = f2_mul(mask2[p], f2_fma(dP2y, qy[p], f2_fma(dP2x, qx[p], dP2z)));
        g2acc[2 * p][0] += net1[2 * p] ? 0.f : dd2.x;          g2acc[2 * p][1] += net1[2 * p] ? dd2.x : 0.f;
        g2acc[2 * p + 1][0] += net1[2 * p + 1] ? 0.f : dd2.y;  g2acc[2 * p + 1][1] += net1[2 * p + 1] ? dd2.y : 0.f;
        if (kFull) {
          // d/dq: through P2 = (q d2) and the external consumers (flows, pass-2 lifters)
          const F2 dqx = f2_fma(dP2x, d2[p], f2_make(xqx[2 * p], xqx[2 * p + 1]));
          const F2 dqy = f2_fma(dP2y, d2[p], f2_make(xqy[2 * p], xqy[2 * p + 1]));
          const F2 dQx = f2_fma(dqx, izq[p], f2_mul(Fx[p], g32));
          const F2 dQy = f2_fma(dqy, izq[p], f2_mul(Fy[p], g32));
          const F2 sq = f2_fma(dqy, qy[p], f2_mul(dqx, qx[p]));
          const F2 dQz = f2_fma(f2_mul(sq, m1), izq[p], f2_mul(Fz[p], g32));
          // d/dP = R^T dQ + pair + own bone (the children's bones are added below)
          F2 Tx, Ty, Tz;
          f2_matTvec(R2, dQx, dQy, dQz, Tx, Ty, Tz);
          const float c_hl = -hl * (1.f / 16.f) * imean * imean;
          const F2 dl = f2_fma(h[p], f2_splat(imean), f2_splat(c_hl));
          const F2 w = f2_make(len[p].x > 0.f ? dl.x * fast_rcp(len[p].x) : 0.f, len[p].y > 0.f ? dl.y * fast_rcp(len[p].y) : 0.f);
          dvx[p] = f2_mul(w, ex[p]); dvy[p] = f2_mul(w, ey[p]); dvz[p] = f2_mul(w, ez[p]);   // d/d(P_parent); d/d(P_child) = -dv
          dPx[p] = f2_fma(dvx[p], m1, f2_fma(Ex[p], ge2, Tx));
          dPy[p] = f2_fma(dvy[p], m1, f2_fma(Ey[p], ge2, Ty));
          dPz[p] = f2_fma(dvz[p], m1, f2_fma(Ez[p], ge2, Tz));
          dap = f2_fma(dQz, Qy[p], dap);   dap = f2_fma(P2z[p], Ry, dap);
          dan = f2_fma(dQy, Qz[p], dan);   dan = f2_fma(P2y[p], Rz, dan);
          dgp = f2_fma(Py[p], Tz, dgp);    dgp = f2_fma(dSy, Sz[p], dgp);
          dgn = f2_fma(Pz[p], Ty, dgn);    dgn = f2_fma(dSz, Sy[p], dgn);
        }
      }
      if (kFull) {
        // bones: the parent joint receives +dv of each child bone (slot k = element k & 1 of pair k >> 1).  Mirror of the
        // forward exchange: the child of (lane q, slot k) is slot k of the NEXT lane (for lane 3: slot k + 1 of lane 0).
        Vec3 dvs[4];
        dvs[0].x = dvx[0].x; dvs[0].y = dvy[0].x; dvs[0].z = dvz[0].x;
        dvs[1].x = dvx[0].y; dvs[1].y = dvy[0].y; dvs[1].z = dvz[0].y;
        dvs[2].x = dvx[1].x; dvs[2].y = dvy[1].x; dvs[2].z = dvz[1].x;
        dvs[3].x = dvx[1].y; dvs[3].y = dvy[1].y; dvs[3].z = dvz[1].y;
        const int nxt = m.gb | ((m.q + 1) & 3);
        Vec3 W[4];
#pragma unroll
        for (int k = 0; k < 3; ++k) W[k] = shfl3(m.q == 0 ? dvs[k + 1] : dvs[k], nxt);   // lane 0 hands lane 3 its next slot
        W[3] = shfl3(dvs[3], nxt);
        const Vec3 b11 = shfl3(dvs[2], m.gb | 2), b14 = shfl3(dvs[3], m.gb | 1);          // bones of joints 11, 14 hang on joint 8
        // not a child: joint 4 (root) for joint 3; joints 7 (root), 11 (joint 8) for joints 6, 10; joint 14 (joint 8) for joint 13;
        // joint 16 has no child
        if (m.q != 2) { dPx[0].x += W[0].x; dPy[0].x += W[0].y; dPz[0].x += W[0].z; }
        if (m.q != 1) { dPx[0].y += W[1].x; dPy[0].y += W[1].y; dPz[0].y += W[1].z; }
        if (m.q != 1) { dPx[1].x += W[2].x; dPy[1].x += W[2].y; dPz[1].x += W[2].z; }
        if (m.q == 1 || m.q == 2) { dPx[1].y += W[3].x; dPy[1].y += W[3].y; dPz[1].y += W[3].z; }
        if (m.q == 3) { dPx[0].y += b11.x + b14.x; dPy[0].y += b11.y + b14.y; dPz[0].y += b11.z + b14.z; }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const F2 dd1 = f2_mul(mask[p], f2_fma(dPy[p], uy2[p], f2_fma(dPx[p], ux2[p], dPz[p])));     // lift
          g1acc[2 * p][0] += net1[2 * p] ? 0.f : dd1.x;          g1acc[2 * p][1] += net1[2 * p] ? dd1.x : 0.f;
          g1acc[2 * p + 1][0] += net1[2 * p + 1] ? 0.f : dd1.y;  g1acc[2 * p + 1][1] += net1[2 * p + 1] ? dd1.y : 0.f;
        }
      }
    }
    if (kFull) { da_acc = (dap.x + dap.y) - (dan.x + dan.y); dg_acc = (dgp.x + dgp.y) - (dgn.x + dgn.y); }
    // ---- head gradients (slot -> column col[j] of every net that feeds joint j in some variant) into the staged output
    //      rows; every other column of those rows -- the root's included -- stays zero
#pragma unroll
    for (int net = 0; net < 2; ++net) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int go = GEOM_TAB(gs, net, k);              // scratch if `net` never feeds this joint
        const __nv_bfloat16 hv = __float2bfloat16_rn(kFull ? g1acc[k][net] : g2acc[k][net]);
        *reinterpret_cast<__nv_bfloat16*>(outb + go + rs) = hv;
        if (kT) {
          __nv_bfloat16* gT = kFull ? A.g1T[net] : A.g2T[net];
          if (gT && valid && go != S.trash_off)
            gT[static_cast<size_t>(A.maps.col[m.j0 + k * m.js]) * A.ldT + A.colT0 + r.n] = hv;
        }
      }
      if (kT && m.q == 0 && T.gs[net][0] != S.trash_off) {
        __nv_bfloat16* gT = kFull ? A.g1T[net] : A.g2T[net];
        if (gT && valid) gT[static_cast<size_t>(A.maps.col[0]) * A.ldT + A.colT0 + r.n] = __float2bfloat16_rn(0.f);
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
    __syncwarp();
    geom_stage_out(M, it, A.N, outb);
    __syncwarp();
  }
  geom_cp_wait<0>();
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

// ---------------------------------------------------------------------------------------------------------
// Host side: the staging plan of one launch.  level 0: forward, 1: losses + d/d(pass-2 heads), 2: full backward.
// Strided tensors ([N, 32] fp32 head / input-gradient rows, [N, 64] bf16 gradient rows: 128-byte pitch) are grouped:
// slices whose used columns fall into one 128-byte window share one staged region, so the four heads of a pass that the
// lifter engine packs into one row (MlpSet head_groups) cost one 128-byte row; only the 16-byte chunks that hold used
// columns are moved.  Returns LINKS_E_ALIGN if a contiguous tensor is not 16-byte aligned.
// ---------------------------------------------------------------------------------------------------------
inline int geom_plan(GeomArgs& A, int level) {
  GeomStage& S = A.st;
  memset(&S, 0, sizeof(S));
  const int V = A.maps.V;
  const int rows = kGeomRows;
  S.rows = rows;
  int maxcol[2] = {0, 0};
  for (int v = 0; v < V; ++v)
    for (int j = 0; j < kJ; ++j) {
      const int net = A.maps.src_net[v][j];
      if (A.maps.col[j] > maxcol[net]) maxcol[net] = A.maps.col[j];
    }
  struct Item { const char* p; int used; int* slot; };
  Item items[10];
  int n_items = 0;
  auto push = [&](const void* p, int used, int* slot) {
    if (p == nullptr) return;
    items[n_items].p = static_cast<const char*>(p); items[n_items].used = used; items[n_items].slot = slot;
    ++n_items;
  };
  for (int h = 0; h < 2; ++h) {
    push(A.head[h], 4 * (maxcol[h] + 1), &S.head_off[h]);
    push(A.ang[h], 4, &S.ang_off[h]);
    if (level >= 1) push(A.head2[h], 4 * (maxcol[h] + 1), &S.head2_off[h]);
    if (level >= 2) push(A.dlift[h], 8 * A.maps.n_joints[h], &S.dlift_off[h]);
  }
  for (int i = 1; i < n_items; ++i) {            // insertion sort by address
    Item t = items[i];
    int k = i;
    while (k > 0 && items[k - 1].p > t.p) { items[k] = items[k - 1]; --k; }
    items[k] = t;
  }
  int in_off = 0, n_in = 0, chunks = 0;
  for (int i = 0; i < n_items;) {
    const char* lo = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(items[i].p) & ~static_cast<uintptr_t>(15));
    const char* hi = items[i].p + items[i].used;
    int k = i + 1;
    while (k < n_items && items[k].p + items[k].used - lo <= 128) {
      if (items[k].p + items[k].used > hi) hi = items[k].p + items[k].used;
      ++k;
    }
    if (hi - lo > 128 || n_in >= kGeomMaxRegions) return LINKS_E_RANGE;
    GeomRegion& R = S.in[n_in++];
    R.g = const_cast<char*>(lo);
    R.pitch = 4 * LINKS_HEAD_LD;
    R.cpr = static_cast<int>((hi - lo + 15) / 16);
    R.soff = in_off;
    R.spitch = kGeomSp;
    for (int q = i; q < k; ++q) *items[q].slot = in_off + static_cast<int>(items[q].p - lo);
    in_off += rows * R.spitch;
    chunks += rows * R.cpr;
    i = k;
  }
  // bytes [128, 144) of every row slot of a strided slice are never a copy destination: they stay zero and serve, with the
  // strided row offset or with pitch 0, as the source of entries that have none
  S.zero_off = S.in[0].soff + 128;
  int rc = 0;
  auto contig_in = [&](const void* p, int pitch, int* slot) {
    if (p == nullptr) return;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0 || ((rows * pitch) & 15) != 0 || n_in >= kGeomMaxRegions) { rc = LINKS_E_ALIGN; return; }
    GeomRegion& R = S.in[n_in++];
    R.g = const_cast<char*>(static_cast<const char*>(p));
    R.pitch = pitch; R.cpr = 0; R.soff = in_off; R.spitch = 0;
    *slot = in_off;
    in_off += (rows * pitch + 15) & ~15;
    chunks += rows * pitch / 16;
  };
  contig_in(A.u, 2 * kJ * 4, &S.u_off);
  contig_in(A.eps_x, 4, &S.eps_off);
  contig_in(A.u_y, 4, &S.uy_off);
  if (level >= 2)
    for (int h = 0; h < 2; ++h) contig_in(A.dflow[h], 8 * A.maps.n_joints[h], &S.dflow_off[h]);
  S.n_in = n_in; S.in_bytes = in_off; S.in_chunks = chunks;
  if (chunks > 32 * kGeomMaxCopyIters) return LINKS_E_RANGE;
  // outputs
  int out_off = 0, n_out = 0, ochunks = 0;
  if (level == 0) {
    for (int h = 0; h < 2; ++h) {
      if (A.qpart[h] == nullptr) continue;
      const int pitch = 8 * A.maps.n_joints[h];
      if ((reinterpret_cast<uintptr_t>(A.qpart[h]) & 15u) != 0) return LINKS_E_ALIGN;
      GeomRegion& R = S.out[n_out++];
      R.g = reinterpret_cast<char*>(A.qpart[h]); R.pitch = pitch; R.cpr = 0; R.soff = out_off; R.spitch = 0;
      S.qpart_off[h] = out_off;
      out_off += (rows * pitch + 15) & ~15;
      ochunks += rows * pitch / 16;
    }
  } else {
    for (int h = 0; h < 2; ++h) {
      __nv_bfloat16* g = level == 2 ? A.g1[h] : A.g2[h];
      if (g == nullptr) continue;
      if ((reinterpret_cast<uintptr_t>(g) & 15u) != 0) return LINKS_E_ALIGN;
      GeomRegion& R = S.out[n_out++];
      R.g = reinterpret_cast<char*>(g); R.pitch = 128; R.cpr = (2 * (maxcol[h] + 1) + 15) / 16; R.soff = out_off;
      R.spitch = kGeomSp;
      S.g_off[h] = out_off;
      out_off += rows * R.spitch;
      ochunks += rows * R.cpr;
    }
  }
  // destinations that do not exist (a net that never feeds a joint, a joint outside every part) land in a scratch slice
  // (strided form: the unused tail of the first gradient slice's row slots; contiguous form, pitch 0: one 16-byte slot)
  if (level >= 1 && n_out > 0) {
    S.trash_off = S.out[0].soff + 128;
  } else {
    S.trash_off = out_off;
    out_off += level == 0 ? 16 : rows * kGeomSp;
  }
  S.n_out = n_out; S.out_bytes = out_off; S.out_chunks = ochunks;
  if (ochunks > 32 * kGeomMaxOutIters) return LINKS_E_RANGE;
  return rc;
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

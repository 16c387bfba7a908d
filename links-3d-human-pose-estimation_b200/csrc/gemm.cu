// Grouped bf16 GEMM with fused epilogue for sm_100a -- persistent, warp-specialised:
//   TMA -> shared memory (128B swizzle, 4-stage ring) -> tcgen05.mma (M128 x N256 x K16, fp32 accumulators in TMEM,
//   two accumulator buffers) -> tcgen05.ld -> fused epilogue -> 16-byte global stores straight from registers
//   Each epilogue warp owns a [32 rows x 64 columns] block of the tile (two passes of 32 columns).  Its bf16 operands /
//   results cross between the row-per-thread register layout (what tcgen05.ld delivers) and global memory through
//   private swizzled shared-memory scratch blocks, so every global access of a warp covers whole sectors (8 rows x
//   64 B per instruction).  Row-per-thread global accesses touch 32 different lines per instruction and made the epilogue
//   2x slower than the main loop (L1TEX tag-stage bound, measured 8-10 us per tile against 4.2 us of MMA).
//
// Replaces every nn.Linear(+LeakyReLU, +residual) of reference utils/models_def.py (forward) and its autograd
// backward (dgrad, wgrad); see include/links_b200.h for the epilogue contract.
//
// One CTA per SM; the two CTAs of a cluster form a tcgen05 CTA PAIR (cta_group::2) that computes a 256x256 output
// tile: each CTA holds its own 128 rows of A, HALF of the shared B tile and its own 128x256 fp32 accumulator in TMEM;
// the leader CTA issues one M=256 MMA for both.  Per CTA and k-block that is 16 KB of A + 16 KB of B instead of
// 16 + 32 KB: 1.5x less TMA traffic from L2 and 1.5x less shared-memory operand traffic -- the two measured limiters
// of the single-CTA version (steady state ~9 TB/s of L2->SM traffic; isolated main loop at 75 % of the MMA floor).
//   warp 0      : TMA producer (one elected lane), runs ahead across tile boundaries
//   warp 1      : TMEM allocator + tcgen05.mma issuer (warp-uniform control flow, one elected lane)
//   warps 4..19 : epilogue (TMEM lane group = warp % 4, 64-column slab = (warp - 4) / 4; warps 2-3 idle, see setmaxnreg); the epilogue of tile i overlaps
//                 the main loop of tile i + 1 through the second accumulator buffer.
// Operands may be K-major (row-major with the contraction dimension contiguous) or MN-major (contraction dimension
// strided): dgrad reads W itself as an MN-major B operand and wgrad reads the row-major activations / gradients as
// MN-major A and B, so no transposed copies of weights or activations are ever written.
#include <cuda.h>
#include <algorithm>
#include <mutex>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"

// A/B switches of the fused-optimiser epilogue (see profiles/): streaming cache policy and L2 prefetch of p / m / v
#ifndef LINKS_ADAM_CS
#define LINKS_ADAM_CS 0
#endif
#ifndef LINKS_ADAM_PREFETCH
#define LINKS_ADAM_PREFETCH 1
#endif
// L2 policy of the weight-gradient operand loads (G and X, both MN-major).  Every 256-column slice of G / X is streamed by
// four tiles of the problem while the fused optimiser streams 26 B per parameter through the same L2: 0 = default policy,
// 1 = evict_last on the TMA loads of weight-gradient tiles (the re-read operands outlive the once-touched p / m / v lines).
#ifndef LINKS_WGRAD_EVICT_LAST
#define LINKS_WGRAD_EVICT_LAST 0
#endif

namespace links {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kEpiWarps = 16;     // 4 TMEM lane groups x 4 column quarters: 4 warps per scheduler hide the epilogue's latencies
constexpr int kChunk = 16;        // accumulator columns per TMEM load / scalar-path chunk
constexpr int kSlab = 64;         // columns of the [32 rows x 64 cols] block one epilogue warp owns
constexpr int kFirstEpiWarp = 4;  // warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps}; warpgroups 1-4 = epilogue
constexpr int kThreads = (kFirstEpiWarp + kEpiWarps) * 32;
// setmaxnreg budget: 128 threads x kRegsCtl + 512 threads x kRegsEpi <= 64 K registers, and per scheduler partition
// (one control warp + four epilogue warps) 32 x (kRegsCtl + 4 kRegsEpi) <= 16 K.  The kernel is compiled for 96.
constexpr int kRegsCtl = 32, kRegsEpi = 112;
constexpr int kAccCols = BN;      // fp32 accumulator columns per buffer
constexpr int kTmemCols = 2 * kAccCols;
constexpr uint32_t kStageBytesA = BM * BK * 2;   // 16 KB
constexpr uint32_t kStageBytesB = (BN / 2) * BK * 2;   // 16 KB: this CTA's half of the pair's B tile
constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
constexpr uint32_t kOffStage = 0;
constexpr uint32_t kScratchBytes = 2 * 32 * 64 + 256 + 64;     // per epilogue warp: two operand/output scratches ([32 rows][64 B] each) + 64 bias floats + next-tile operand descriptor
constexpr uint32_t kOffScratch = kStages * kStageBytes;
constexpr uint32_t kOffBar = kOffScratch + kEpiWarps * kScratchBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024 /*align slack*/;

// barriers: full[kStages], empty[kStages], acc_full[2], acc_empty[2]
// chain launches: tile_done[4] -- the 16 epilogue warps of a CTA arrive when their stores of a tile are issued; the
// signalling warp (warp 2) then publishes the tile's completion counter (see the kernel).  Four slots: an epilogue warp
// can run at most two tiles ahead of the slowest one (it needs that warp's accumulator release two tiles back).
enum { GB_FULL = 0, GB_EMPTY = kStages, GB_ACCFULL = 2 * kStages, GB_ACCEMPTY = 2 * kStages + 2, GB_TILEDONE = 2 * kStages + 4,
       GB_COUNT = 2 * kStages + 8 };

struct alignas(64) GemmProblemDev {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int M, N, K;
  int tile_begin, pairs_m, tiles_n;     // scheduling unit: a pair of M-adjacent tiles (one per CTA of the cluster)
  uint32_t flags;
  int ld_add0, ld_add1, ld_ymask, ld_bits, ld_sign, ld_f32, ld_out, ld_mid;
  const float* bias;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  const __nv_bfloat16* ymask;
  const uint32_t* bits;
  uint32_t* sign_out;
  float* out_f32;
  __nv_bfloat16* out;
  __nv_bfloat16* mid;
  int a_boxes, b_boxes; // TMA boxes per stage actually needed (MN-major: 64-wide slabs; K-major: 1)
  int b_rows;           // rows of the (K-major) B tile actually loaded per stage, split in two halves across the pair
  uint32_t stage_tx;    // bytes the boxes of one stage deliver (small M / N problems load smaller boxes)
  int vec_ok;           // every present epilogue operand is 16-byte aligned with a vector-friendly leading dimension
  int epi_mode;         // index into kEpiMask (0 = run-time flags)
  // ---- fused Adam (weight-gradient problems): the tile's gradient updates p / m / v and the bf16 shadow in place
  float* adam_p;
  float* adam_m;
  float* adam_v;
  __nv_bfloat16* adam_shadow;
  const float* adam_hyper;
  int ld_shadow;
  // ---- reduce-scatter by peer stores (weight-gradient problems under data parallelism)
  __nv_bfloat16* push[LINKS_MAX_PUSH_RANKS];
  int push_rows, ld_push;
  // ---- chain launches only (links_gemm_chain_*): tile-level dependencies through completion counters in global memory
  int cnt_base;         // first completion counter of this problem (one per PAIR of M tiles); -1: nobody waits on it
  int dep_base[3];      // [A operand, add0, add1]: first counter of the producing problem of the chain, -1: none
  int dep_blocks[3];    // 0: wait for the counter of the tile's own row block (base + pair index);
                        // n > 0: wait for all n counters of the producer (operand contracted over rows: wgrad)
  int dep_need[3];      // arrivals that complete a counter = 2 CTAs * kEpiWarps * tiles_n of the producer
};

// Kernel argument of a chain launch: everything lives in the caller's workspace (global memory).
struct ChainDev {
  const GemmProblemDev* probs;
  const uint32_t* sched;     // [n_cl][sched_ld] tiles of every cluster in execution order: pi << 22 | pair_m << 11 | tn
  const int* sched_cnt;      // [n_cl]
  int* counters;             // [n_counters] completion counters, zero between launches (the last cluster to exit resets them)
  int* exit_cnt;
  int n_counters, sched_ld;
};

struct GemmGroupDev {
  GemmProblemDev p[LINKS_MAX_GEMM_PROBLEMS];
  int n_problems;
  int total_tiles;      // number of pair-tiles
};

// MN-major operand tile (contraction dimension strided), 128-byte swizzle: each 64-element slab along M/N is
// [BK rows][128 B]; 8-row groups 1024 B apart (SBO), slabs BK*128 B apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((BK * 128) >> 4) << 16;        // LBO = 8192 B
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&o)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  o[0] = bf16_bits_to_f32(q.x & 0xFFFFu); o[1] = bf16_bits_to_f32(q.x >> 16);
  o[2] = bf16_bits_to_f32(q.y & 0xFFFFu); o[3] = bf16_bits_to_f32(q.y >> 16);
  o[4] = bf16_bits_to_f32(q.z & 0xFFFFu); o[5] = bf16_bits_to_f32(q.z >> 16);
  o[6] = bf16_bits_to_f32(q.w & 0xFFFFu); o[7] = bf16_bits_to_f32(q.w >> 16);
}
__device__ __forceinline__ void unpack8_bf16(const uint4& q, float (&o)[8]) {
  o[0] = bf16_bits_to_f32(q.x & 0xFFFFu); o[1] = bf16_bits_to_f32(q.x >> 16);
  o[2] = bf16_bits_to_f32(q.y & 0xFFFFu); o[3] = bf16_bits_to_f32(q.y >> 16);
  o[4] = bf16_bits_to_f32(q.z & 0xFFFFu); o[5] = bf16_bits_to_f32(q.z >> 16);
  o[6] = bf16_bits_to_f32(q.w & 0xFFFFu); o[7] = bf16_bits_to_f32(q.w >> 16);
}

// Register-resident copy of one problem's epilogue parameters.
struct EpiParams {
  int M, N;
  uint32_t flags;
  int vec_ok;
  int ld_add0, ld_add1, ld_ymask, ld_bits, ld_sign, ld_f32, ld_out, ld_mid;
  const float* bias;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  const __nv_bfloat16* ymask;
  const uint32_t* bits;
  uint32_t* sign_out;
  float* out_f32;
  __nv_bfloat16* out;
  __nv_bfloat16* mid;
  float* adam_p;
  float* adam_m;
  float* adam_v;
  __nv_bfloat16* adam_shadow;
  const float* adam_hyper;
  int ld_shadow;
  __nv_bfloat16* const* push;      // -> the problem descriptor's push[] (global / parameter memory)
  int push_rows, ld_push;
};

// Element-wise path for one 16-column chunk of row m: N tails and operands that are not 16-byte aligned (heads,
// upscale dgrad / wgrad).  Same arithmetic as the vector path.  Fully inlined on purpose: a noinline helper would pin
// the parameter block and the accumulators in local memory (measured: ~0.5 us per element).
__device__ __forceinline__ void epilogue_chunk_scalar(const EpiParams& E, const uint32_t (&acc)[kChunk], const float* sbias16,
                                                      int m, int n0) {
  const int sh = n0 & 16;
  const size_t mo = static_cast<size_t>(m);
  uint32_t bits_word = 0;
  if (E.bits) bits_word = __ldg(E.bits + mo * E.ld_bits + (n0 >> 5)) >> sh;
  uint32_t sign_word = 0;
  const bool accum = (E.flags & LINKS_EPI_ACCUM_F32) != 0;
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const int n = n0 + i;
    if (n < E.N) {
      float v = __uint_as_float(acc[i]) + sbias16[i];                 // staged bias (zero when the problem has none)
      sign_word |= (v > 0.f ? 0u : 1u) << i;
      if (E.flags & LINKS_EPI_LEAKY_PRE) v = links_leaky(v);
      if (E.flags & LINKS_EPI_RELU_PRE) v = fmaxf(v, 0.f);
      if (E.add0) v += __bfloat162float(E.add0[mo * E.ld_add0 + n]);
      if (E.add1) v += __bfloat162float(E.add1[mo * E.ld_add1 + n]);
      if (E.flags & LINKS_EPI_LEAKY_POST) v = links_leaky(v);
      if (E.ymask) v *= (__bfloat162float(E.ymask[mo * E.ld_ymask + n]) > 0.f ? 1.f : ((E.flags & LINKS_EPI_YMASK_ZERO) ? 0.f : 0.01f));
      if (E.mid) E.mid[mo * E.ld_mid + n] = __float2bfloat16_rn(v);
      if (E.bits) v *= ((bits_word >> i) & 1u) ? 0.01f : 1.f;
      if (E.out) E.out[mo * E.ld_out + n] = __float2bfloat16_rn(v);
      if (E.out_f32) {
        float* p = E.out_f32 + mo * E.ld_f32 + n;
        *p = accum ? *p + v : v;
      }
    }
  }
  if (E.sign_out)
    reinterpret_cast<unsigned short*>(E.sign_out + mo * E.ld_sign + (n0 >> 5))[sh >> 4] = static_cast<unsigned short>(sign_word);
}

// ----------------------------------------------------------------------------------------------
// Epilogue specialisations.  The per-problem combination of epilogue steps is classified on the host into one of a
// few modes; each mode is a compile-time feature mask so the hot loop is straight-line code (mask 0 = decide
// everything at run time).  The arithmetic is identical in every mode (see include/links_b200.h).
// ----------------------------------------------------------------------------------------------
enum : uint32_t { F_BIAS = 1, F_LPRE = 2, F_RPRE = 4, F_ADD0 = 8, F_ADD1 = 16, F_LPOST = 32, F_YMASK = 64, F_MID = 128,
                  F_BITS = 256, F_SIGN = 512, F_OUT = 1024, F_F32 = 2048, F_ACC = 4096, F_ADAM = 8192, F_PUSH = 16384 };
constexpr int kEpiModes = 13;
__device__ constexpr uint32_t kEpiMask[kEpiModes] = {
    0u,
    F_BIAS | F_OUT,                                             // 1  upscale forward
    F_BIAS | F_LPRE | F_OUT,                                    // 2  res-block l1 forward
    F_BIAS | F_LPRE | F_ADD0 | F_LPOST | F_SIGN | F_OUT,        // 3  res-block l2 forward (training)
    F_BIAS | F_LPRE | F_ADD0 | F_LPOST | F_OUT,                 // 4  res-block l2 forward (inference)
    F_YMASK | F_OUT,                                            // 5  l2 dgrad
    F_ADD0 | F_OUT,                                             // 6  l1 dgrad into a plain sum
    F_ADD0 | F_YMASK | F_MID | F_BITS | F_OUT,                  // 7  l1 dgrad + masks of the previous block
    F_ADD0 | F_ADD1 | F_YMASK | F_MID | F_BITS | F_OUT,         // 8  l1 dgrad merging two branches
    F_YMASK | F_MID | F_BITS | F_OUT,                           // 9  head dgrad
    F_F32,                                                      // 10 wgrad
    F_ADAM,                                                     // 11 wgrad with the optimiser step fused (no gradient is stored)
    F_PUSH,                                                     // 12 wgrad stored as bf16 into the OWNER rank's staging buffer (P2P)
};

// arrive on the barrier at the same shared offset in CTA `rank` of the cluster.  Relaxed: the TMEM reads are ordered by
// tcgen05.fence::before_thread_sync; a release would wait for all of the warp's outstanding global stores (ERRBAR, ~9 %).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// all but the most recently committed group have landed
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// scratch addressing: [32 rows][64 B]; 16-byte chunk j (0..3) of row r, XOR-swizzled so that both the row-per-thread
// and the coalesced (8 rows x 64 B per instruction) access patterns are bank-conflict free
__device__ __forceinline__ uint32_t scr_addr(uint32_t S, int r, int j) {
  return S + static_cast<uint32_t>(r * 64 + ((j ^ ((r >> 1) & 3)) << 4));
}

// Asynchronous coalesced fetch of a bf16 [32 x 32] block (rows m0.., columns n0..) into a scratch:
// instruction i moves rows 8i .. 8i+7, 64 contiguous bytes (two full sectors) each.
__device__ __forceinline__ void block_fetch_async(uint32_t S, const __nv_bfloat16* base, int ld, int m0, int n0, int M, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2), j = lane & 3;
    if (m0 + r < M) cp_async16(scr_addr(S, r, j), base + static_cast<size_t>(m0 + r) * ld + n0 + j * 8);
  }
}
// Coalesced store of a scratch (bf16 [32 x 32]) to global memory.
__device__ __forceinline__ void block_store(uint32_t S, __nv_bfloat16* base, int ld, int m0, int n0, int M, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2), j = lane & 3;
    const uint4 v = lds128(scr_addr(S, r, j));
    if (m0 + r < M) *reinterpret_cast<uint4*>(base + static_cast<size_t>(m0 + r) * ld + n0 + j * 8) = v;
  }
}

constexpr int kHalf = 32;   // columns processed per pass of a warp over its 64-column slab

// The operand block that is fetched AHEAD of its use: add0 when the problem has one, else ymask.  The descriptor of the
// warp's NEXT tile lives in the warp's scratch (32 bytes, written by lane 0 at the start of a tile) rather than in
// registers: it is warp-uniform, needed only at the two refill points, and keeping it live across a pass cost spills.
struct PrimaryOp {
  const __nv_bfloat16* p;
  int ld, M, m0, n0;
  int ok;             // present, 16-byte aligned and the warp's 64-column slab lies fully inside N (the vector path)
  int pad;
  int dep_tile;       // chain launches: 1 + index (in this cluster's schedule) of the tile the operand belongs to when it is
                      // produced inside the chain -- its fetch may only start once the scout warp has seen the tile's
                      // dependencies complete (deps_ok >= dep_tile); 0: no dependency
  int pad2, pad3, pad4;
};
static_assert(sizeof(PrimaryOp) == 48, "PrimaryOp slot");
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Returns true when the fetch was SKIPPED because the operand's producer tiles have not completed yet (chain launches
// only).  The prefetch runs a tile ahead of the producer warp's own dependency wait and must never block: the warp
// still owes the completion signal of its CURRENT tile, which the missing producer may (transitively) be waiting for.
// A skipped fetch is repeated at the top of the next tile, behind the accumulator barrier (all dependencies done).
__device__ __forceinline__ int ld_acquire_cta_shared(uint32_t saddr) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
template <bool kChain>
__device__ __forceinline__ bool prefetch_primary(const PrimaryOp* slot, uint32_t S, int h, int lane, uint32_t deps_ok_addr) {
  const uint4 a = *reinterpret_cast<const uint4*>(slot);              // p, ld, M
  const uint4 b = *(reinterpret_cast<const uint4*>(slot) + 1);        // m0, n0, ok
  bool missed = false;
  if (b.z != 0u) {
    if (kChain) {
      const int dep_tile = *reinterpret_cast<const int*>(reinterpret_cast<const uint4*>(slot) + 2);
      if (dep_tile > 0) missed = __any_sync(0xffffffffu, ld_acquire_cta_shared(deps_ok_addr) < dep_tile);
    }
    if (!missed) {
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(static_cast<uint64_t>(a.x) | (static_cast<uint64_t>(a.y) << 32));
      block_fetch_async(S, p, static_cast<int>(a.z), static_cast<int>(b.x), static_cast<int>(b.y) + h * 32, static_cast<int>(a.w), lane);
    }
  }
  cp_async_commit();                      // one group per pass, empty or not: keeps the wait_group arithmetic uniform
  return missed;
}

// One [32 rows x 64 columns] block of an accumulator tile per warp (this thread: row m0 + lane), in two passes of
// 32 columns.  Pass h owns scratch S0 + 2048 h: it holds the pass's primary operand (cp.async group committed TWO passes
// earlier, i.e. a full pass of work ahead), then any secondary operand, then stages the pass's outputs for the
// coalesced stores, and is finally refilled with the primary operand of the same pass of the warp's NEXT tile.
//   t_addr : TMEM address of (lane group, first column of the block);  sbias : bias of the block's 64 columns (smem)
template <uint32_t F, bool kChain>
__device__ __forceinline__ bool epilogue_block(const EpiParams& E, uint32_t t_addr, const float* sbias, uint32_t S0,
                                               const PrimaryOp* nxt, int m0, int n_blk, int lane, uint32_t acc_empty_bar,
                                               uint32_t deps_ok_addr) {
  constexpr bool kDyn = F == 0u;
  bool missed = false;
  const int m = m0 + lane;
  const bool row_ok = m < E.M;
  const bool has_bias = kDyn ? E.bias != nullptr : (F & F_BIAS) != 0;
  const bool lpre = kDyn ? (E.flags & LINKS_EPI_LEAKY_PRE) != 0 : (F & F_LPRE) != 0;
  const bool rpre = kDyn ? (E.flags & LINKS_EPI_RELU_PRE) != 0 : (F & F_RPRE) != 0;
  const bool has_add0 = kDyn ? E.add0 != nullptr : (F & F_ADD0) != 0;
  const bool has_add1 = kDyn ? E.add1 != nullptr : (F & F_ADD1) != 0;
  const bool lpost = kDyn ? (E.flags & LINKS_EPI_LEAKY_POST) != 0 : (F & F_LPOST) != 0;
  const bool has_y = kDyn ? E.ymask != nullptr : (F & F_YMASK) != 0;
  const bool has_mid = kDyn ? E.mid != nullptr : (F & F_MID) != 0;
  const bool has_bits = kDyn ? E.bits != nullptr : (F & F_BITS) != 0;
  const bool has_sign = kDyn ? E.sign_out != nullptr : (F & F_SIGN) != 0;
  const bool has_out = kDyn ? E.out != nullptr : (F & F_OUT) != 0;
  const bool has_f32 = kDyn ? E.out_f32 != nullptr : (F & F_F32) != 0;
  const bool has_adam = kDyn ? false : (F & F_ADAM) != 0;          // specialised mode only (host rejects other combinations)
  const bool has_push = kDyn ? false : (F & F_PUSH) != 0;
  const size_t mo = static_cast<size_t>(m);
  const float yneg = (E.flags & LINKS_EPI_YMASK_ZERO) ? 0.f : 0.01f;  // slope applied where the mask activation is <= 0
  const bool fast = E.vec_ok && (n_blk + kSlab <= E.N);               // warp-uniform

  if (!fast) {
    // tails / unaligned operands: element-wise path straight from TMEM, 16 columns at a time
#pragma unroll 1
    for (int c = 0; c < kSlab / kChunk; ++c) {
      uint32_t acc[kChunk];
      tmem_ld16(t_addr + static_cast<uint32_t>(c * kChunk), acc);
      if (row_ok && n_blk + c * kChunk < E.N) epilogue_chunk_scalar(E, acc, sbias + c * kChunk, m, n_blk + c * kChunk);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(acc_empty_bar, 0);
    missed |= prefetch_primary<kChain>(nxt, S0, 0, lane, deps_ok_addr);
    missed |= prefetch_primary<kChain>(nxt, S0 + 2048u, 1, lane, deps_ok_addr);
    return missed;
  }

  const bool y_primary = has_y && !has_add0;     // ymask is the operand fetched ahead when there is no add0
  uint32_t sign_prev = 0;
#pragma unroll 1
  for (int h = 0; h < kSlab / kHalf; ++h) {
    const int n0 = n_blk + h * kHalf;
    const uint32_t SA = S0 + static_cast<uint32_t>(h) * 2048u, SB = SA;
    // ---- accumulator row -> registers; after the second pass the TMEM buffer is free for the MMA warp
    float v[kHalf];
#pragma unroll
    for (int c = 0; c < kHalf / kChunk; ++c) {
      uint32_t acc[kChunk];
      tmem_ld16(t_addr + static_cast<uint32_t>(h * kHalf + c * kChunk), acc);
#pragma unroll
      for (int i = 0; i < kChunk; ++i) v[c * kChunk + i] = __uint_as_float(acc[i]);
    }
    if (h == kSlab / kHalf - 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_bar, 0);
    }
    uint32_t bits_word = 0;
    if (has_bits && row_ok) bits_word = __ldg(E.bits + mo * E.ld_bits + (n0 >> 5));
    if (has_bias) {
#pragma unroll
      for (int i = 0; i < kHalf; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sbias + h * kHalf + i);
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    }
    if (has_sign) {
      uint32_t sw = 0;
#pragma unroll
      for (int i = 0; i < kHalf; ++i) sw |= (v[i] > 0.f ? 0u : 1u) << i;
      // both passes' words of a row are adjacent: one 8-byte store per row instead of two 4-byte ones (each is a
      // separate 32-byte sector per lane -- the row-per-thread pattern the L1 tag stage serialises)
      if ((E.ld_sign & 1) == 0) {
        if (h == 1 && row_ok) *reinterpret_cast<uint2*>(E.sign_out + mo * E.ld_sign + (n_blk >> 5)) = make_uint2(sign_prev, sw);
        sign_prev = sw;
      } else if (row_ok) {
        E.sign_out[mo * E.ld_sign + (n0 >> 5)] = sw;
      }
    }
    if (lpre) {
#pragma unroll
      for (int i = 0; i < kHalf; ++i) v[i] = fmaxf(v[i], 0.01f * v[i]);      // == leaky for slope < 1
    }
    if (rpre) {
#pragma unroll
      for (int i = 0; i < kHalf; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    // bf16 operand blocks: fetched coalesced into scratch SA, read back row-per-thread
    if (has_add0) {
      cp_async_wait_but_one();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t[8]; unpack8_bf16(lds128(scr_addr(SA, lane, j)), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j * 8 + i] += t[i];
      }
      __syncwarp();
    }
    if (has_add1) {
      block_fetch_async(SA, E.add1, E.ld_add1, m0, n0, E.M, lane);
      cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t[8]; unpack8_bf16(lds128(scr_addr(SA, lane, j)), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j * 8 + i] += t[i];
      }
      __syncwarp();
    }
    if (has_y && !y_primary) block_fetch_async(SA, E.ymask, E.ld_ymask, m0, n0, E.M, lane);   // in flight during the leaky below
    if (lpost) {
#pragma unroll
      for (int i = 0; i < kHalf; ++i) v[i] = fmaxf(v[i], 0.01f * v[i]);
    }
    if (has_y) {
      if (y_primary) cp_async_wait_but_one(); else cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t[8]; unpack8_bf16(lds128(scr_addr(SA, lane, j)), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j * 8 + i] = t[i] > 0.f ? v[j * 8 + i] : yneg * v[j * 8 + i];
      }
      __syncwarp();
    }
    if (has_mid) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(scr_addr(SB, lane, j), make_uint4(pack_bf16x2(v[j * 8], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                                pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7])));
      __syncwarp();
      block_store(SB, E.mid, E.ld_mid, m0, n0, E.M, lane);
      __syncwarp();
    }
    if (has_bits) {
#pragma unroll
      for (int i = 0; i < kHalf; ++i) v[i] = ((bits_word >> i) & 1u) ? 0.01f * v[i] : v[i];
    }
    if (has_out) {
#ifdef LINKS_GEMM_TRACE
      if (!(E.flags & (1u << 30)))
#endif
      {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(scr_addr(SB, lane, j), make_uint4(pack_bf16x2(v[j * 8], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                                  pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7])));
        __syncwarp();
        block_store(SB, E.out, E.ld_out, m0, n0, E.M, lane);
        __syncwarp();
      }
    }
    if (has_push) {
      // Data-parallel reduce-scatter fused into the weight-gradient GEMM: rows [r * push_rows, (r+1) * push_rows) of dW
      // belong to rank r, whose staging buffer (peer-mapped over NVLink) receives this block as bf16 -- plain coalesced
      // stores, overlapped with the main loops of the following tiles; the owner sums the ranks' slots in its Adam kernel.
      const int owner = m0 / E.push_rows;                              // warp-uniform (push_rows is a multiple of 128)
      __nv_bfloat16* pbase = E.push[owner];
      const int lrow = m0 - owner * E.push_rows;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(scr_addr(SB, lane, j), make_uint4(pack_bf16x2(v[j * 8], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                                pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7])));
      __syncwarp();
      block_store(SB, pbase, E.ld_push, lrow, n0, E.push_rows, lane);
      __syncwarp();
    }
    if (has_f32) {
      // fp32 [32 x 32] block in two quarters of 16 columns (one scratch row = 16 floats); coalesced read-modify-write
      const bool accum = (E.flags & LINKS_EPI_ACCUM_F32) != 0;
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(scr_addr(SB, lane, j), make_uint4(__float_as_uint(v[qq * 16 + j * 4]), __float_as_uint(v[qq * 16 + j * 4 + 1]),
                                                  __float_as_uint(v[qq * 16 + j * 4 + 2]), __float_as_uint(v[qq * 16 + j * 4 + 3])));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + (lane >> 2), j = lane & 3;
          const uint4 q = lds128(scr_addr(SB, r, j));
          if (m0 + r < E.M) {
            float4* p = reinterpret_cast<float4*>(E.out_f32 + static_cast<size_t>(m0 + r) * E.ld_f32 + n0 + qq * 16 + j * 4);
            float4 o = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
            if (accum) { const float4 a = *p; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
            *p = o;
          }
        }
        __syncwarp();
      }
    }
    if (has_adam) {
      // torch.optim.Adam with coupled L2 decay on this [32 x 32] block of the weight matrix, in the coalesced domain:
      // g = acc * grad_scale + wd * p; m, v, p updated in place; the bf16 shadow (the GEMM operand of the next step) is
      // refreshed from the new p.  Arithmetic = csrc/elementwise.cuh::adam_kernel.  The block is staged through BOTH
      // scratches of the warp as [32 rows][128 B] (this mode has no operand prefetch in flight), so that every global
      // access of the warp covers 4 rows x one full 128-byte line of p / m / v (64-byte half lines at a 4 KB stride ran
      // the read-modify-write at a third of the HBM rate).
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(E.adam_hyper));        // lr / (1 - b1^t), sqrt(1 - b2^t), eps, b1
      const float4 h1 = __ldg(reinterpret_cast<const float4*>(E.adam_hyper) + 1);    // b2, weight decay, grad scale, -
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(S0 + static_cast<uint32_t>(lane * 128 + ((j ^ (lane & 7)) << 4)),
               make_uint4(__float_as_uint(v[j * 4]), __float_as_uint(v[j * 4 + 1]), __float_as_uint(v[j * 4 + 2]), __float_as_uint(v[j * 4 + 3])));
      __syncwarp();
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3), j = lane & 7;
        const uint4 q = lds128(S0 + static_cast<uint32_t>(r * 128 + ((j ^ (r & 7)) << 4)));
        if (m0 + r < E.M) {
          const size_t off = static_cast<size_t>(m0 + r) * E.ld_f32 + n0 + j * 4;
          float4* pp = reinterpret_cast<float4*>(E.adam_p + off);
          float4* pm = reinterpret_cast<float4*>(E.adam_m + off);
          float4* pv = reinterpret_cast<float4*>(E.adam_v + off);
#if LINKS_ADAM_CS
          float4 P4 = __ldcs(pp), M4 = __ldcs(pm), V4 = __ldcs(pv);
#else
          float4 P4 = *pp, M4 = *pm, V4 = *pv;
#endif
          float gq[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
          float pa[4] = {P4.x, P4.y, P4.z, P4.w}, ma[4] = {M4.x, M4.y, M4.z, M4.w}, va[4] = {V4.x, V4.y, V4.z, V4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float gi = gq[e] * h1.z;
            gi = gi + h1.y * pa[e];
            const float mi = ma[e] + (gi - ma[e]) * (1.f - h0.w);
            const float vi = va[e] * h1.x + (1.f - h1.x) * gi * gi;
            const float denom = sqrtf(vi) / h0.y + h0.z;
            pa[e] = pa[e] - h0.x * (mi / denom);
            ma[e] = mi;
            va[e] = vi;
          }
#if LINKS_ADAM_CS
          __stcs(pp, make_float4(pa[0], pa[1], pa[2], pa[3]));
          __stcs(pm, make_float4(ma[0], ma[1], ma[2], ma[3]));
          __stcs(pv, make_float4(va[0], va[1], va[2], va[3]));
#else
          *pp = make_float4(pa[0], pa[1], pa[2], pa[3]);
          *pm = make_float4(ma[0], ma[1], ma[2], ma[3]);
          *pv = make_float4(va[0], va[1], va[2], va[3]);
#endif
          *reinterpret_cast<uint2*>(E.adam_shadow + static_cast<size_t>(m0 + r) * E.ld_shadow + n0 + j * 4) =
              make_uint2(pack_bf16x2(pa[0], pa[1]), pack_bf16x2(pa[2], pa[3]));
        }
      }
      __syncwarp();
      // both scratches were used as staging: the next tile's operand prefetches start after the last pass
      if (h == kSlab / kHalf - 1) {
        missed |= prefetch_primary<kChain>(nxt, S0, 0, lane, deps_ok_addr);
        missed |= prefetch_primary<kChain>(nxt, S0 + 2048u, 1, lane, deps_ok_addr);
      }
      continue;
    }
    // this pass's scratch is free again: refill it with the primary operand of the same pass of the next tile
    missed |= prefetch_primary<kChain>(nxt, SA, h, lane, deps_ok_addr);
  }
  return missed;
}

// Run-time-flag variant (rare combinations, e.g. the flow-training GEMMs) kept OUT of line: inlined next to the
// specialised modes its register demand made the compiler spill state that is live across the mode switch in every
// mode.  The caller passes copies, so only those copies are pinned in local memory.
template <bool kChain>
__device__ __noinline__ bool epilogue_block_dyn(const EpiParams& E, uint32_t t_addr, const float* sbias, uint32_t S0,
                                                const PrimaryOp* nxt, int m0, int n_blk, int lane, uint32_t acc_empty_bar,
                                                uint32_t deps_ok_addr) {
  return epilogue_block<0u, kChain>(E, t_addr, sbias, S0, nxt, m0, n_blk, lane, acc_empty_bar, deps_ok_addr);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a CTA pair: the data lands in the executing CTA's shared memory, the transaction bytes are signalled on
// the LEADER CTA's barrier (peer bit of the shared-window address cleared), which is the one the MMA issuer waits on.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// same with an L2 cache-policy operand (createpolicy encodings as used by CUTLASS: full-fraction evict_last)
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// pair MMA: D[256 x N] (128 rows in each CTA's TMEM) += A[256 x 16] * B[N x 16]^T, operands split across the two CTAs
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// pair MMA completion -> arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

#ifdef LINKS_GEMM_TRACE
// debug build only (scratch/trace_gemm.py): per-CTA timestamps of the pipeline phases
__device__ unsigned long long g_gemm_trace[148 * 16];
__device__ __forceinline__ void trace_mark(int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (blockIdx.x < 148) g_gemm_trace[blockIdx.x * 16 + slot] = t;
}
#define TRACE(slot) trace_mark(slot)
// chain launches: per cluster (leader CTA) and tile, 8 timestamps:
//   0 producer: dependencies cleared   1 MMA: first operands landed   2 MMA: last commit issued
//   3 epilogue (first epilogue warp): accumulator full   4 epilogue: tile done   5 scout: dependencies seen complete
constexpr int kTraceTiles = 96;
__device__ unsigned long long g_chain_trace[74 * kTraceTiles * 8];
__device__ __forceinline__ void ctrace(int cl, int ti, int slot) {
  if (cl < 74 && ti < kTraceTiles) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_chain_trace[(cl * kTraceTiles + ti) * 8 + slot] = t;
  }
}
#define CTRACE(cl, ti, slot) ctrace(cl, ti, slot)
#else
#define TRACE(slot)
#define CTRACE(cl, ti, slot)
#endif

struct TileCoord { int pi, tm, tn; };
__device__ __forceinline__ TileCoord tile_coord(const GemmGroupDev& G, int tile) {
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < G.n_problems; ++i) if (tile >= G.p[i].tile_begin) pi = i;
  const int local = tile - G.p[pi].tile_begin;
  TileCoord t;
  t.pi = pi;
  t.tn = local / G.p[pi].pairs_m;          // consecutive pair-tiles share the B (weight) slice
  t.tm = local - t.tn * G.p[pi].pairs_m;   // pair index along M: the CTA's tile is 2 * tm + cluster rank
  return t;
}

// The two launch flavours share one kernel body:
//   grouped (links_gemm_grouped): <= 8 independent problems, descriptors in the kernel parameters, static round-robin
//     of the pair-tiles over the clusters;
//   chain (links_gemm_chain_run): a whole dependent sequence of layers in ONE launch -- descriptors, a host-built
//     per-cluster tile schedule and tile-completion counters live in global memory; a tile of layer l+1 starts as soon
//     as the tiles of layer l that produce its operand rows have signalled, so setup / TMEM allocation / pipeline fill
//     are paid once per chain instead of once per layer and the epilogue of layer l overlaps the main loop of layer l+1.
template <bool kChain> struct KernelArgs { typedef GemmGroupDev type; };
template <> struct KernelArgs<true> { typedef ChainDev type; };

template <bool kChain>
__device__ __forceinline__ const GemmProblemDev& problem_of(const typename KernelArgs<kChain>::type& A, int pi);
template <> __device__ __forceinline__ const GemmProblemDev& problem_of<false>(const GemmGroupDev& A, int pi) { return A.p[pi]; }
template <> __device__ __forceinline__ const GemmProblemDev& problem_of<true>(const ChainDev& A, int pi) { return A.probs[pi]; }

// i-th tile of cluster cl_id (false: no more tiles).  tc.tm is the PAIR index along M.
template <bool kChain>
__device__ __forceinline__ bool tile_at(const typename KernelArgs<kChain>::type& A, int cl_id, int n_cl, int n_mine, int i, TileCoord& tc);
template <> __device__ __forceinline__ bool tile_at<false>(const GemmGroupDev& A, int cl_id, int n_cl, int, int i, TileCoord& tc) {
  const int tile = cl_id + i * n_cl;
  if (tile >= A.total_tiles) return false;
  tc = tile_coord(A, tile);
  return true;
}
template <> __device__ __forceinline__ bool tile_at<true>(const ChainDev& A, int cl_id, int, int n_mine, int i, TileCoord& tc) {
  if (i >= n_mine) return false;
  const uint32_t e = __ldg(A.sched + static_cast<size_t>(cl_id) * A.sched_ld + i);
  tc.pi = static_cast<int>(e >> 22);
  tc.tm = static_cast<int>((e >> 11) & 0x7FFu);
  tc.tn = static_cast<int>(e & 0x7FFu);
  return true;
}

// Block until every dependency of a tile of problem P (pair index pair_m) has completed: spin on the producers'
// completion counters (bounded: a protocol bug traps instead of hanging the GPU), then order the TMA (async proxy)
// reads that follow behind the acquire.
__device__ __forceinline__ void chain_wait_deps(const GemmProblemDev& P, const int* counters, int pair_m) {
#pragma unroll 1
  for (int d = 0; d < 3; ++d) {
    const int base = P.dep_base[d];
    if (base < 0) continue;
    const int nb = P.dep_blocks[d], need = P.dep_need[d];
    const int* c = counters + base + (nb == 0 ? pair_m : 0);
    const int n = nb == 0 ? 1 : nb;
#pragma unroll 1
    for (int b = 0; b < n; ++b) {
      if (ld_acquire_gpu(c + b) >= need) continue;
      const long long t0 = clock64();
      while (ld_acquire_gpu(c + b) < need) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// Kernel
// ----------------------------------------------------------------------------------------------
template <bool kChain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ typename KernelArgs<kChain>::type G) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // 1024-B aligned (128B swizzle atoms)
  const uint32_t bars = base + kOffBar;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (bars + GB_COUNT * 8 - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int cl_id = blockIdx.x >> 1, n_cl = gridDim.x >> 1;
  if (threadIdx.x == 0) TRACE(0);
  // chain launches: number of tiles of this cluster's schedule whose dependencies the scout warp has seen complete
  const uint32_t deps_ok_addr = bars + GB_COUNT * 8 + 16;

  if (warp == 0 && lane == 0) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(deps_ok_addr), "r"(0) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bars + (GB_FULL + s) * 8, 1);
      mbar_init(bars + (GB_EMPTY + s) * 8, 1);     // released by the leader's pair-MMA commit (multicast to both CTAs)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + (GB_ACCFULL + s) * 8, 1);
      mbar_init(bars + (GB_ACCEMPTY + s) * 8, 2 * kEpiWarps);   // leader's copy: one arrival per epilogue warp of BOTH CTAs
    }
    for (int s = 0; s < 4; ++s) mbar_init(bars + (GB_TILEDONE + s) * 8, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair(bars + GB_COUNT * 8, kTmemCols);
  tc_fence_before();
  cluster_sync_all();                                     // both CTAs' barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) TRACE(1);
  // Programmatic dependent launch: everything above (barriers, TMEM, cluster handshake) overlapped the tail of the
  // previous kernel in the stream; from here on this grid reads what that kernel wrote.  Let our own successor be
  // scheduled as soon as SMs free up -- it will wait at the same point for this grid to finish.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int n_mine = 0;                                         // chain: number of tiles in this cluster's schedule
  if (kChain) n_mine = __ldg(reinterpret_cast<const ChainDev&>(G).sched_cnt + cl_id);

  if (warp < kFirstEpiWarp) {
    // the control warpgroup hands registers to the epilogue warpgroups (whose 32-column accumulator slice, operand
    // conversion and store staging did not fit 96 registers without spills).  The two register budgets must not meet
    // again before the end of the kernel: ptxas compiles code after a join for the smaller one.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;   // running k-block index over all tiles of this CTA
      TileCoord tc;
      for (int ti = 0; tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, tc); ++ti) {
        const GemmProblemDev& P = problem_of<kChain>(G, tc.pi);
        if (kChain) {
          // the scout warp has seen every producer tile of this tile complete (acquire at gpu scope, handed over at cta
          // scope); order the TMA (async proxy) reads behind it
          if (ld_acquire_cta_shared(deps_ok_addr) <= ti) {
            const long long t0 = clock64();
            while (ld_acquire_cta_shared(deps_ok_addr) <= ti) {
              if (clock64() - t0 > 4000000000LL) __trap();
            }
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");
          if (cta_rank == 0) CTRACE(cl_id, ti, 0);
        }
        tc.tm = 2 * tc.tm + cta_rank;
        const int num_kb = (P.K + BK - 1) / BK;
        const bool a_mn = (P.flags & LINKS_GEMM_A_MN) != 0, b_mn = (P.flags & LINKS_GEMM_B_MN) != 0;
        const int b_half = P.b_rows >> 1;                 // this CTA's rows (K-major) / columns (MN-major) of the pair's B tile
        const uint32_t stage_tx = P.stage_tx;
        const int a_boxes = P.a_boxes, b_boxes = P.b_boxes;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages, use = it / kStages;
          mbar_wait(bars + (GB_EMPTY + s) * 8, (use & 1u) ^ 1u);        // own slot free (leader's commit reaches both CTAs)
          const uint32_t full = bars + (GB_FULL + s) * 8;
          if (cta_rank == 0) mbar_expect_tx(full, 2u * stage_tx);       // the leader's barrier collects both CTAs' bytes
          const uint32_t sA = base + kOffStage + s * kStageBytes, sB = sA + kStageBytesA;
          if (!a_mn) {
            tma_load_2d_pair(sA, &P.tmA, full, kb * BK, tc.tm * BM);
          } else {
#pragma unroll
            for (int q = 0; q < BM / 64; ++q)
              if (q < a_boxes) {
                if (LINKS_WGRAD_EVICT_LAST && b_mn) tma_load_2d_pair_hint(sA + q * (BK * 128), &P.tmA, full, tc.tm * BM + q * 64, kb * BK, kL2EvictLast);
                else tma_load_2d_pair(sA + q * (BK * 128), &P.tmA, full, tc.tm * BM + q * 64, kb * BK);
              }
          }
          if (!b_mn) {
            tma_load_2d_pair(sB, &P.tmB, full, kb * BK, tc.tn * BN + cta_rank * b_half);
          } else {
#pragma unroll
            for (int q = 0; q < BN / 128; ++q)
              if (q < b_boxes) {
                if (LINKS_WGRAD_EVICT_LAST && a_mn) tma_load_2d_pair_hint(sB + q * (BK * 128), &P.tmB, full, tc.tn * BN + cta_rank * b_half + q * 64, kb * BK, kL2EvictLast);
                else tma_load_2d_pair(sB + q * (BK * 128), &P.tmB, full, tc.tn * BN + cta_rank * b_half + q * 64, kb * BK);
              }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool leader = elect_one() && cta_rank == 0;     // only the leader CTA issues the pair's MMAs
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t bb = sb + kOffBar;
    uint32_t it = 0, lt = 0;   // k-block counter, local tile counter
    TileCoord tc;
    for (int ti = 0; cta_rank == 0 && tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, tc); ++ti, ++lt) {   // leader CTA only
      const GemmProblemDev& P = problem_of<kChain>(G, tc.pi);
      const int num_kb = (P.K + BK - 1) / BK;
      const bool a_mn = (P.flags & LINKS_GEMM_A_MN) != 0, b_mn = (P.flags & LINKS_GEMM_B_MN) != 0;
      // N actually needed by this tile (multiple of 16): the MMA cost scales with it
      int n_eff = P.N - tc.tn * BN;
      n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
                             (static_cast<uint32_t>(n_eff >> 3) << 17) | (static_cast<uint32_t>((2 * BM) >> 4) << 24);
      const uint32_t slot = lt & 1u, acc_use = lt >> 1;
      mbar_wait(bb + (GB_ACCEMPTY + slot) * 8, (acc_use & 1u) ^ 1u);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_addr = tb + slot * kAccCols;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t s = it % kStages, use = it / kStages;
        mbar_wait(bb + (GB_FULL + s) * 8, use & 1u);
        tc_fence_after();
        if (leader && it == 0) TRACE(2);
        if (kChain && leader && kb == 0) CTRACE(cl_id, ti, 1);
        if (leader) {
          const uint32_t sA = sb + kOffStage + s * kStageBytes, sB = sA + kStageBytesA;
          const uint64_t adesc = a_mn ? make_smem_desc_mn128(sA) : make_smem_desc_k128(sA);
          const uint64_t bdesc = b_mn ? make_smem_desc_mn128(sB) : make_smem_desc_k128(sB);
          // per UMMA_K step: K-major advances 32 B inside the swizzle row, MN-major advances 16 rows of 128 B
          const uint32_t a_step = a_mn ? 128u : 2u, b_step = b_mn ? 128u : 2u;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_pair(d_addr, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(bb + (GB_EMPTY + s) * 8, 3);                  // frees the slot in BOTH CTAs when the MMAs retire
          if (kb == num_kb - 1) umma_commit_pair(bb + (GB_ACCFULL + slot) * 8, 3);   // both CTAs' epilogues
          if (kChain && kb == num_kb - 1) CTRACE(cl_id, ti, 2);
        }
        __syncwarp();
      }
    }
  } else if (kChain && warp == 3) {
    // ================= dependency scout (chain launches) =================
    // Walks this cluster's schedule ahead of the producer: spins on the completion counters of each tile's producer
    // tiles (gpu-scope acquire loads: an L2 round trip each) and publishes the number of cleared tiles in shared memory,
    // where the TMA producer and the epilogue's operand prefetch read it at shared-memory latency.
    if (lane == 0) {
      TileCoord tc;
      for (int ti = 0; tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, tc); ++ti) {
        chain_wait_deps(problem_of<kChain>(G, tc.pi), reinterpret_cast<const ChainDev&>(G).counters, tc.tm);
        asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(deps_ok_addr), "r"(ti + 1) : "memory");
        if (cta_rank == 0) CTRACE(cl_id, ti, 5);
      }
    }
  } else if (kChain && warp == 2) {
    // ================= completion signaller (chain launches) =================
    // Publishing a tile needs a gpu-scope fence behind the epilogue's stores; issued by the epilogue warps themselves it
    // sat on their critical path (the epilogue is the longer side of the per-tile pipeline).  They only arrive on a
    // CTA-local barrier (release.cta) and move on; this otherwise idle warp observes the barrier (acquire.cta), fences at
    // gpu scope -- cumulative over the stores it has synchronised with -- and bumps the tile's completion counter.
    if (lane == 0) {
      TileCoord tc;
      for (int ti = 0; tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, tc); ++ti) {
        mbar_wait(bars + (GB_TILEDONE + (ti & 3)) * 8, static_cast<uint32_t>(ti >> 2) & 1u);
        const int cb = problem_of<kChain>(G, tc.pi).cnt_base;
        if (cb >= 0) {
          int* c = reinterpret_cast<const ChainDev&>(G).counters + cb + tc.tm;
          asm volatile("fence.proxy.async.global;\n\tfence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.s32 [%0], %1;"
                       ::"l"(c), "r"(kEpiWarps) : "memory");
        }
      }
    }
  }
  } else {
    // ================= epilogue warps (512 threads) =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    const int lane_grp = warp & 3;                                    // TMEM lanes 32*lane_grp .. +31
    const int slab = (warp - kFirstEpiWarp) >> 2;                                 // which 64 columns of the tile
    const bool store_thread = threadIdx.x == kFirstEpiWarp * 32;
    const uint32_t S = base + kOffScratch + static_cast<uint32_t>(warp - kFirstEpiWarp) * kScratchBytes;
    uint32_t lt = 0;
    // bias of this warp's 64 columns: two floats per lane, loaded one tile ahead, published through the warp's own
    // scratch (no CTA-wide barrier in the epilogue: the 16 warps run fully decoupled)
    float* sbias = reinterpret_cast<float*>(smem_raw + (S + 4096u - raw));
    auto load_bias = [&](int ti, float& b0, float& b1) {
      b0 = 0.f; b1 = 0.f;
      TileCoord t0;
      if (tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, t0)) {
        const GemmProblemDev& P0 = problem_of<kChain>(G, t0.pi);
        const float* bp = P0.bias;
        const int nb = t0.tn * BN + slab * kSlab + lane;
        if (bp != nullptr) {
          const int N0 = P0.N;
          if (nb < N0) b0 = __ldg(bp + nb);
          if (nb + 32 < N0) b1 = __ldg(bp + nb + 32);
        }
      }
    };
    float bias_n0, bias_n1;
    load_bias(0, bias_n0, bias_n1);
    // primary operand (add0, else ymask) of a tile's two passes for this warp -> the warp's descriptor slot
    PrimaryOp* const nxt = reinterpret_cast<PrimaryOp*>(smem_raw + (S + 4096u + 256u - raw));
    auto publish_primary = [&](int ti) {
      if (lane == 0) {
        PrimaryOp o;
        o.p = nullptr; o.ld = 0; o.M = 0; o.m0 = 0; o.n0 = 0; o.ok = 0; o.pad = 0; o.dep_tile = 0; o.pad2 = 0; o.pad3 = 0; o.pad4 = 0;
        TileCoord t0;
        if (tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, t0)) {
          const GemmProblemDev& P = problem_of<kChain>(G, t0.pi);
          o.p = P.add0 != nullptr ? P.add0 : P.ymask;
          o.ld = P.add0 != nullptr ? P.ld_add0 : P.ld_ymask;
          o.M = P.M;
          o.m0 = (2 * t0.tm + cta_rank) * BM + lane_grp * 32;
          o.n0 = t0.tn * BN + slab * kSlab;
          o.ok = (o.p != nullptr && P.vec_ok && o.n0 + kSlab <= P.N) ? 1 : 0;
          if (kChain && P.add0 != nullptr && P.dep_base[1] >= 0) o.dep_tile = ti + 1;
        }
        *nxt = o;
      }
      __syncwarp();
    };
    publish_primary(0);                                      // in flight while the first main loop runs
    bool missed = prefetch_primary<kChain>(nxt, S, 0, lane, deps_ok_addr);
    missed |= prefetch_primary<kChain>(nxt, S + 2048u, 1, lane, deps_ok_addr);
    TileCoord tc;
    for (int ti = 0; tile_at<kChain>(G, cl_id, n_cl, n_mine, ti, tc); ++ti, ++lt) {
      const int pair_m = tc.tm;
      tc.tm = 2 * tc.tm + cta_rank;
      // register copy of the problem's epilogue parameters (an indexed constant-bank load per use otherwise)
      EpiParams E;
      int mode;
      {
        const GemmProblemDev& P = problem_of<kChain>(G, tc.pi);
        E.M = P.M; E.N = P.N; E.flags = P.flags; E.vec_ok = P.vec_ok;
        E.ld_add0 = P.ld_add0; E.ld_add1 = P.ld_add1; E.ld_ymask = P.ld_ymask; E.ld_bits = P.ld_bits; E.ld_sign = P.ld_sign;
        E.ld_f32 = P.ld_f32; E.ld_out = P.ld_out; E.ld_mid = P.ld_mid;
        E.bias = P.bias; E.add0 = P.add0; E.add1 = P.add1; E.ymask = P.ymask; E.bits = P.bits; E.sign_out = P.sign_out;
        E.out_f32 = P.out_f32; E.out = P.out; E.mid = P.mid;
        E.adam_p = P.adam_p; E.adam_m = P.adam_m; E.adam_v = P.adam_v; E.adam_shadow = P.adam_shadow;
        E.adam_hyper = P.adam_hyper; E.ld_shadow = P.ld_shadow;
        E.push = P.push; E.push_rows = P.push_rows; E.ld_push = P.ld_push;
        mode = P.epi_mode;
      }
      const uint32_t slot = lt & 1u, acc_use = lt >> 1;
      const int m0 = tc.tm * BM + lane_grp * 32, n0 = tc.tn * BN + slab * kSlab;
      __syncwarp();                                  // every lane is done with the previous tile's bias and descriptor
      publish_primary(ti + 1);
      sbias[lane] = bias_n0;
      sbias[lane + 32] = bias_n1;
      __syncwarp();
      load_bias(ti + 1, bias_n0, bias_n1);           // next tile's bias: in flight during this tile's epilogue
#if LINKS_ADAM_PREFETCH
      if (mode == 11 && n0 + kSlab <= E.N && m0 + lane < E.M) {
        // fused optimiser: pull this warp's [32 x 64] blocks of p, m, v into L2 while the tile's main loop runs
        const size_t off = static_cast<size_t>(m0 + lane) * E.ld_f32 + n0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(E.adam_p + off + q * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(E.adam_m + off + q * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(E.adam_v + off + q * 32));
        }
      }
#endif
      mbar_wait(bars + (GB_ACCFULL + slot) * 8, acc_use & 1u);
      tc_fence_after();
      if (kChain && missed) {
        // The operand prefetch of THIS tile was skipped (its producer had not signalled yet).  The accumulator barrier
        // implies the producer warp has seen every dependency of the tile complete: fetch both passes now.
        (void)ld_acquire_cta_shared(deps_ok_addr);
        const __nv_bfloat16* pp = E.add0 != nullptr ? E.add0 : E.ymask;
        const int pld = E.add0 != nullptr ? E.ld_add0 : E.ld_ymask;
        if (pp != nullptr && E.vec_ok && n0 + kSlab <= E.N) {
          block_fetch_async(S, pp, pld, m0, n0, E.M, lane);
          block_fetch_async(S + 2048u, pp, pld, m0, n0 + 32, E.M, lane);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
      }
      missed = false;
      if (store_thread && lt < 3) TRACE(3 + 3 * lt);
      if (kChain && store_thread && cta_rank == 0) CTRACE(cl_id, ti, 3);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + slot * kAccCols + slab * kSlab;
      const uint32_t ae = bars + (GB_ACCEMPTY + slot) * 8;
      if (n0 >= E.N) {
        // this warp's slab lies entirely beyond N: nothing to do but release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ae, 0);
        missed |= prefetch_primary<kChain>(nxt, S, 0, lane, deps_ok_addr);
        missed |= prefetch_primary<kChain>(nxt, S + 2048u, 1, lane, deps_ok_addr);
      } else {
        const float* sb = sbias;
        switch (mode) {
          case 1: missed = epilogue_block<kEpiMask[1], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 2: missed = epilogue_block<kEpiMask[2], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 3: missed = epilogue_block<kEpiMask[3], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 4: missed = epilogue_block<kEpiMask[4], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 5: missed = epilogue_block<kEpiMask[5], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 6: missed = epilogue_block<kEpiMask[6], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 7: missed = epilogue_block<kEpiMask[7], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 8: missed = epilogue_block<kEpiMask[8], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 9: missed = epilogue_block<kEpiMask[9], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 10: missed = epilogue_block<kEpiMask[10], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 11: missed = epilogue_block<kEpiMask[11], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          case 12: missed = epilogue_block<kEpiMask[12], kChain>(E, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr); break;
          default: {
            const EpiParams Ed = E;
            missed = epilogue_block_dyn<kChain>(Ed, t_addr, sb, S, nxt, m0, n0, lane, ae, deps_ok_addr);
            break;
          }
        }
      }
      if (kChain) {
        // this warp's stores of the tile are issued: tell the signalling warp (release.cta orders them before the arrive)
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (GB_TILEDONE + (lt & 3u)) * 8);
        if (store_thread && cta_rank == 0) CTRACE(cl_id, ti, 4);
      }
      if (store_thread && lt < 3) TRACE(4 + 3 * lt);
    }
    if (store_thread) TRACE(14);
  }
  tc_fence_before();
  cluster_sync_all();                                     // the peer may still multicast into / arrive on this CTA
  if (threadIdx.x == 0) TRACE(15);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
  if (kChain && cta_rank == 0) {
    // the last cluster to get here zeroes the completion counters for the next launch of this chain (every other
    // cluster has finished all of its waits and signals)
    __shared__ int s_last;
    const ChainDev& CH = reinterpret_cast<const ChainDev&>(G);
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(CH.exit_cnt, 1) == n_cl - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      for (int i = threadIdx.x; i < CH.n_counters; i += kThreads) CH.counters[i] = 0;
      if (threadIdx.x == 0) *CH.exit_cnt = 0;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major matrix [rows, cols] with leading dimension ld (elements); box = box_cols x box_rows, 128B swizzle.
static int encode_2d(EncodeTiledFn fn, CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_cols,
                     int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : LINKS_E_DRIVER;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Validate one problem and build its device descriptor (tensor maps included).
static int build_problem(EncodeTiledFn fn, const LinksGemmProblem& s, GemmProblemDev& d) {
  if (!s.A || !s.B || s.M < 1 || s.N < 1 || s.K < 1) return LINKS_E_ARG;
  const bool a_mn = (s.flags & LINKS_GEMM_A_MN) != 0, b_mn = (s.flags & LINKS_GEMM_B_MN) != 0;
  if (!aligned16(s.A) || !aligned16(s.B) || (s.lda & 7) || (s.ldb & 7)) return LINKS_E_ALIGN;
  if (s.lda < (a_mn ? s.M : s.K) || s.ldb < (b_mn ? s.N : s.K)) return LINKS_E_ALIGN;
  // small problems load smaller boxes: rows the MMA never needs are not fetched (or zero-filled) at all
  const int a_rows = s.M >= BM ? BM : ((s.M + 7) & ~7), b_rows = s.N >= BN ? BN : ((s.N + 15) & ~15);
  d.a_boxes = a_mn ? (s.M >= BM ? BM / 64 : (s.M + 63) / 64) : 1;
  d.b_boxes = b_mn ? (b_rows / 2 + 63) / 64 : 1;       // per CTA: 64-wide slabs covering its half of the B tile
  d.stage_tx = static_cast<uint32_t>((a_mn ? d.a_boxes * 64 : a_rows) * BK * 2 + (b_mn ? d.b_boxes * 64 : b_rows / 2) * BK * 2);
  int rc = a_mn ? encode_2d(fn, &d.tmA, s.A, s.K, s.M, s.lda, 64, BK) : encode_2d(fn, &d.tmA, s.A, s.M, s.K, s.lda, BK, a_rows);
  if (rc) return rc;
  d.b_rows = b_rows;
  rc = b_mn ? encode_2d(fn, &d.tmB, s.B, s.K, s.N, s.ldb, 64, BK) : encode_2d(fn, &d.tmB, s.B, s.N, s.K, s.ldb, BK, b_rows / 2);
  if (rc) return rc;
  d.M = s.M; d.N = s.N; d.K = s.K;
  d.pairs_m = ((s.M + BM - 1) / BM + 1) / 2;
  d.tiles_n = (s.N + BN - 1) / BN;
  d.flags = s.flags;
  if ((s.out && s.ld_out < s.N) || (s.mid && s.ld_mid < s.N)) return LINKS_E_ALIGN;
  d.out = static_cast<__nv_bfloat16*>(s.out); d.ld_out = s.ld_out;
  d.mid = static_cast<__nv_bfloat16*>(s.mid); d.ld_mid = s.ld_mid;
  bool vec = true;
  auto chk = [&](const void* p, int ld, int mult) {
    if (p && (!aligned16(p) || (ld % mult) != 0)) vec = false;
  };
  chk(s.bias, 4, 4); chk(s.add0, s.ld_add0, 8); chk(s.add1, s.ld_add1, 8); chk(s.ymask, s.ld_ymask, 8);
  chk(s.mid, s.ld_mid, 8); chk(s.out, s.ld_out, 8); chk(s.out_f32, s.ld_f32, 4);
  d.vec_ok = vec ? 1 : 0;
  {
    uint32_t f = 0;
    if (s.bias) f |= F_BIAS;
    if (s.flags & LINKS_EPI_LEAKY_PRE) f |= F_LPRE;
    if (s.flags & LINKS_EPI_RELU_PRE) f |= F_RPRE;
    if (s.add0) f |= F_ADD0;
    if (s.add1) f |= F_ADD1;
    if (s.flags & LINKS_EPI_LEAKY_POST) f |= F_LPOST;
    if (s.ymask) f |= F_YMASK;
    if (s.mid) f |= F_MID;
    if (s.bits) f |= F_BITS;
    if (s.sign_out) f |= F_SIGN;
    if (s.out) f |= F_OUT;
    if (s.out_f32) f |= F_F32;
    if (s.adam_p) f |= F_ADAM;
    if (s.push_rows > 0) f |= F_PUSH;
    static const uint32_t host_masks[kEpiModes] = {
        0u, F_BIAS | F_OUT, F_BIAS | F_LPRE | F_OUT, F_BIAS | F_LPRE | F_ADD0 | F_LPOST | F_SIGN | F_OUT,
        F_BIAS | F_LPRE | F_ADD0 | F_LPOST | F_OUT, F_YMASK | F_OUT, F_ADD0 | F_OUT,
        F_ADD0 | F_YMASK | F_MID | F_BITS | F_OUT, F_ADD0 | F_ADD1 | F_YMASK | F_MID | F_BITS | F_OUT,
        F_YMASK | F_MID | F_BITS | F_OUT, F_F32, F_ADAM, F_PUSH};
    d.epi_mode = 0;
    for (int k = 1; k < kEpiModes; ++k) if (host_masks[k] == f) d.epi_mode = k;
    if (s.adam_p) {
      // the fused optimiser exists in the vector path of its own epilogue mode only: a plain weight-gradient problem
      // whose tiles are full 64-column slabs of 16-byte aligned fp32 / bf16 rows
      if (f != F_ADAM || !s.adam_m || !s.adam_v || !s.adam_shadow || !s.adam_hyper) return LINKS_E_ARG;
      if ((s.N % kSlab) != 0 || (s.ld_f32 & 3) || (s.ld_shadow & 7) || s.ld_f32 < s.N || s.ld_shadow < s.N ||
          !aligned16(s.adam_p) || !aligned16(s.adam_m) || !aligned16(s.adam_v) || !aligned16(s.adam_shadow) ||
          !aligned16(s.adam_hyper))
        return LINKS_E_ALIGN;
    }
    if (s.push_rows > 0) {
      if (f != F_PUSH) return LINKS_E_ARG;
      if ((s.push_rows % BM) != 0 || (s.M % s.push_rows) != 0 || s.M / s.push_rows > LINKS_MAX_PUSH_RANKS) return LINKS_E_RANGE;
      if ((s.N % kSlab) != 0 || (s.ld_push & 7) || s.ld_push < s.N) return LINKS_E_ALIGN;
      for (int r = 0; r < s.M / s.push_rows; ++r)
        if (!s.push[r] || !aligned16(s.push[r])) return LINKS_E_ALIGN;
    }
  }
  for (int r = 0; r < LINKS_MAX_PUSH_RANKS; ++r) d.push[r] = static_cast<__nv_bfloat16*>(s.push[r]);
  d.push_rows = s.push_rows; d.ld_push = s.ld_push;
  d.adam_p = s.adam_p; d.adam_m = s.adam_m; d.adam_v = s.adam_v;
  d.adam_shadow = static_cast<__nv_bfloat16*>(s.adam_shadow);
  d.adam_hyper = s.adam_hyper; d.ld_shadow = s.ld_shadow;
  d.ld_add0 = s.ld_add0; d.ld_add1 = s.ld_add1; d.ld_ymask = s.ld_ymask; d.ld_bits = s.ld_bits;
  d.ld_sign = s.ld_sign; d.ld_f32 = s.ld_f32;
  d.bias = s.bias;
  d.add0 = static_cast<const __nv_bfloat16*>(s.add0);
  d.add1 = static_cast<const __nv_bfloat16*>(s.add1);
  d.ymask = static_cast<const __nv_bfloat16*>(s.ymask);
  d.bits = s.bits;
  d.sign_out = s.sign_out;
  d.out_f32 = s.out_f32;
  if (s.bits && s.ld_bits < (s.N + 31) / 32) return LINKS_E_RANGE;
  if (s.sign_out && s.ld_sign < (s.N + 31) / 32) return LINKS_E_RANGE;
  d.cnt_base = -1;
  for (int k = 0; k < 3; ++k) { d.dep_base[k] = -1; d.dep_blocks[k] = 0; d.dep_need[k] = 0; }
  return 0;
}

// Validate one group and build its device descriptor.
static int build_group(EncodeTiledFn fn, const LinksGemmProblem* problems, int n_problems, GemmGroupDev& G) {
  memset(&G, 0, sizeof(G));
  int tiles = 0;
  for (int i = 0; i < n_problems; ++i) {
    GemmProblemDev& d = G.p[i];
    const int rc = build_problem(fn, problems[i], d);
    if (rc) return rc;
    d.tile_begin = tiles;
    tiles += d.pairs_m * d.tiles_n;
  }
  G.n_problems = n_problems;
  G.total_tiles = tiles;
  return 0;
}

// Launch descriptors are pure functions of the problem array; cache them so that repeated (eager) launches of
// the same plan do not pay the cuTensorMapEncodeTiled calls.  Direct-mapped, full-key compare.
struct CacheEntry {
  int n;
  LinksGemmProblem key[LINKS_MAX_GEMM_PROBLEMS];
  GemmGroupDev G;
};
constexpr int kCacheSize = 1024;
static CacheEntry* g_cache = nullptr;
static std::mutex g_cache_mu;
static int g_gemm_max_ctas = 0;  // 0 = all SMs; data-parallel runs leave a few SMs to the NCCL kernels (links_gemm_set_max_ctas)
static bool g_gemm_pdl = true;   // programmatic dependent launch between consecutive GEMM launches (env LINKS_GEMM_PDL=0 disables)
// Per-device host state (one process may drive several devices): SM count + the kernels' shared-memory opt-in.
constexpr int kMaxDevices = 64;
static int g_dev_sms[kMaxDevices] = {0};

// Caller holds g_cache_mu.  Returns the SM count of the current device (> 0) or a negative / CUDA error via *err.
static int device_sms(int* err) {
  *err = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { *err = static_cast<int>(e); return 0; }
  if (dev < 0 || dev >= kMaxDevices) { *err = LINKS_E_RANGE; return 0; }
  if (g_dev_sms[dev] == 0) {
    if ((e = cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess) {
      *err = static_cast<int>(e);
      return 0;
    }
    int sms = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) { *err = static_cast<int>(e); return 0; }
    g_dev_sms[dev] = sms;
    const char* pdl = getenv("LINKS_GEMM_PDL");
    if (pdl != nullptr && pdl[0] == '0') g_gemm_pdl = false;
  }
  return g_dev_sms[dev];
}
static int max_clusters(int sms) { return (g_gemm_max_ctas > 0 && g_gemm_max_ctas < sms ? g_gemm_max_ctas : sms) / 2; }

static unsigned long long g_gemm_launches = 0;   // launches of either flavour since load (links_gemm_launch_count)

static uint64_t hash_bytes(const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
}  // namespace links

extern "C" __attribute__((visibility("default"))) int links_gemm_set_max_ctas(int n) {
  const int prev = links::g_gemm_max_ctas;
  links::g_gemm_max_ctas = n < 0 ? 0 : n;
  return prev;
}

#ifdef LINKS_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int links_debug_gemm_trace(unsigned long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, links::g_gemm_trace, sizeof(unsigned long long) * 148 * 16));
}
extern "C" __attribute__((visibility("default"))) int links_debug_chain_trace(unsigned long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, links::g_chain_trace, sizeof(links::g_chain_trace)));
}
#endif

extern "C" __attribute__((visibility("default"))) int links_gemm_grouped(const LinksGemmProblem* problems, int n_problems, void* stream) {
  using namespace links;
  if (problems == nullptr || n_problems < 1 || n_problems > LINKS_MAX_GEMM_PROBLEMS) return LINKS_E_ARG;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return LINKS_E_DRIVER;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  int derr = 0;
  const int sms = device_sms(&derr);
  if (derr) return derr;
  const size_t key_bytes = sizeof(LinksGemmProblem) * static_cast<size_t>(n_problems);
  const uint64_t h = hash_bytes(problems, key_bytes);
  if (g_cache == nullptr) g_cache = static_cast<CacheEntry*>(calloc(kCacheSize, sizeof(CacheEntry)));
  if (g_cache == nullptr) return LINKS_E_ARG;
  CacheEntry& ce = g_cache[h % kCacheSize];
  if (ce.n != n_problems || memcmp(ce.key, problems, key_bytes) != 0) {
    ce.n = 0;
    int rc = build_group(fn, problems, n_problems, ce.G);
    if (rc) return rc;
    memcpy(ce.key, problems, key_bytes);
    ce.n = n_problems;
  }
  const int max_cl = max_clusters(sms);
  const int grid = 2 * (ce.G.total_tiles < max_cl ? ce.G.total_tiles : max_cl);   // clusters of 2 CTAs
  g_gemm_launches++;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = links_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_gemm_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_kernel<false>, ce.G);
  if (le != cudaSuccess) return static_cast<int>(le);
  return links_launch_status();
}

// ----------------------------------------------------------------------------------------------
// Chain launches (host side): descriptors + dependency counters + a per-cluster tile schedule in the caller's workspace
// ----------------------------------------------------------------------------------------------
namespace links {

struct ChainLayout {
  size_t off_probs, off_sched, off_cnt, off_counters, total;
  int n_cl, sched_ld, n_counters, total_tiles;
};

static size_t rup256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct ChainTile { int pi, tm, tn; };

// Tiles in execution order: levels ascending; inside a level row block by row block, the problems of the level side
// by side, then the N tiles -- the producers of a row block finish together and a consumer of level l+1 finds the same
// distance (one level's worth of tiles) to its producers wherever it sits.
static void chain_tile_order(const LinksChainProblem* problems, const GemmProblemDev* d, int n, std::vector<ChainTile>& tiles) {
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return problems[a].level < problems[b].level; });
  size_t i0 = 0;
  while (i0 < order.size()) {
    size_t i1 = i0;
    int max_pairs = 0;
    while (i1 < order.size() && problems[order[i1]].level == problems[order[i0]].level) {
      max_pairs = std::max(max_pairs, d[order[i1]].pairs_m);
      ++i1;
    }
    bool all_wgrad = true;
    for (size_t j = i0; j < i1; ++j) all_wgrad = all_wgrad && (d[order[j]].flags & LINKS_GEMM_A_MN) != 0;
    if (all_wgrad) {
      // weight-gradient levels: problem by problem.  Every tile of a problem streams the SAME G and X rows along K, so
      // tiles that start together walk them in lockstep and L2 serves all but the first reader; spread over the level
      // (row-block-major) the X operand of a layer was re-read from DRAM by every row of tiles (measured: 3.9 GB of DRAM
      // reads for 2.2 GB of unique operands, L2 hit rate 46 %).
      for (size_t j = i0; j < i1; ++j) {
        const int pi = order[j];
        for (int tm = 0; tm < d[pi].pairs_m; ++tm)
          for (int tn = 0; tn < d[pi].tiles_n; ++tn) tiles.push_back({pi, tm, tn});
      }
    } else {
      for (int tm = 0; tm < max_pairs; ++tm)
        for (size_t j = i0; j < i1; ++j) {
          const int pi = order[j];
          if (tm >= d[pi].pairs_m) continue;
          for (int tn = 0; tn < d[pi].tiles_n; ++tn) tiles.push_back({pi, tm, tn});
        }
    }
    i0 = i1;
  }
}

static int chain_prepare(const LinksChainProblem* problems, int n, std::vector<GemmProblemDev>& d, bool encode) {
  if (problems == nullptr || n < 1 || n > LINKS_MAX_CHAIN_PROBLEMS) return LINKS_E_ARG;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return LINKS_E_DRIVER;
  d.assign(n, GemmProblemDev());
  for (int i = 0; i < n; ++i) {
    memset(&d[i], 0, sizeof(GemmProblemDev));
    if (encode) {
      const int rc = build_problem(fn, problems[i].g, d[i]);
      if (rc) return rc;
    } else {
      const LinksGemmProblem& g = problems[i].g;
      if (g.M < 1 || g.N < 1 || g.K < 1) return LINKS_E_ARG;
      d[i].pairs_m = ((g.M + BM - 1) / BM + 1) / 2;
      d[i].tiles_n = (g.N + BN - 1) / BN;
      d[i].K = g.K;
    }
    if (d[i].pairs_m > 2047 || d[i].tiles_n > 2047) return LINKS_E_RANGE;
    for (int k = 0; k < 3; ++k) {
      const int dp = problems[i].dep[k];
      if (dp >= i) return LINKS_E_RANGE;                       // producers come first: the chain is topologically ordered
      if (dp >= 0 && problems[dp].level >= problems[i].level) return LINKS_E_RANGE;
    }
  }
  return 0;
}

static int chain_layout(const LinksChainProblem* problems, const std::vector<GemmProblemDev>& d, int n, int sms, ChainLayout& L) {
  std::vector<char> is_dep(n, 0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) if (problems[i].dep[k] >= 0) is_dep[problems[i].dep[k]] = 1;
  int counters = 0, tiles = 0;
  for (int i = 0; i < n; ++i) {
    if (is_dep[i]) counters += d[i].pairs_m;
    tiles += d[i].pairs_m * d[i].tiles_n;
  }
  const int max_cl = max_clusters(sms);
  L.n_cl = tiles < max_cl ? tiles : max_cl;
  L.total_tiles = tiles;
  L.n_counters = counters + 1;                                  // + the exit counter (last word)
  L.sched_ld = tiles;                                           // upper bound; the real maximum is known after scheduling
  L.off_probs = 0;
  L.off_sched = rup256(sizeof(GemmProblemDev) * static_cast<size_t>(n));
  // worst-case schedule length per cluster is bounded below once the schedule exists; reserve 2x the mean + slack
  const size_t per = static_cast<size_t>((tiles + L.n_cl - 1) / L.n_cl) * 2 + 8;
  L.sched_ld = static_cast<int>(per);
  L.off_cnt = L.off_sched + rup256(per * L.n_cl * 4);
  L.off_counters = L.off_cnt + rup256(static_cast<size_t>(L.n_cl) * 4);
  L.total = L.off_counters + rup256(static_cast<size_t>(L.n_counters) * 4);
  return 0;
}

}  // namespace links

extern "C" __attribute__((visibility("default"))) size_t links_gemm_chain_ws_bytes(const LinksChainProblem* problems, int n_problems) {
  using namespace links;
  std::vector<GemmProblemDev> d;
  if (chain_prepare(problems, n_problems, d, false)) return 0;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  int derr = 0;
  const int sms = device_sms(&derr);
  if (derr) return 0;
  ChainLayout L;
  chain_layout(problems, d, n_problems, sms, L);
  return L.total;
}

extern "C" __attribute__((visibility("default"))) int links_gemm_chain_build(const LinksChainProblem* problems, int n_problems, void* ws_dev,
                                                                            size_t ws_bytes, LinksGemmChainPlan* plan, void* stream) {
  using namespace links;
  if (ws_dev == nullptr || plan == nullptr) return LINKS_E_ARG;
  if (reinterpret_cast<uintptr_t>(ws_dev) & 255u) return LINKS_E_ALIGN;
  std::vector<GemmProblemDev> d;
  int rc = chain_prepare(problems, n_problems, d, true);
  if (rc) return rc;
  const int n = n_problems;
  int sms;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    int derr = 0;
    sms = device_sms(&derr);
    if (derr) return derr;
  }
  ChainLayout L;
  chain_layout(problems, d, n, sms, L);
  if (ws_bytes < L.total) return LINKS_E_RANGE;
  // ---- completion counters and dependencies
  {
    std::vector<char> is_dep(n, 0);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) if (problems[i].dep[k] >= 0) is_dep[problems[i].dep[k]] = 1;
    int c = 0;
    for (int i = 0; i < n; ++i) if (is_dep[i]) { d[i].cnt_base = c; c += d[i].pairs_m; }
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const int dp = problems[i].dep[k];
        if (dp < 0) continue;
        const bool all_rows = ((problems[i].dep_all_rows >> k) & 1) != 0;
        if (!all_rows && d[dp].pairs_m != d[i].pairs_m) return LINKS_E_RANGE;     // row-block dependencies need equal M tiling
        d[i].dep_base[k] = d[dp].cnt_base;
        d[i].dep_blocks[k] = all_rows ? d[dp].pairs_m : 0;
        d[i].dep_need[k] = 2 * kEpiWarps * d[dp].tiles_n;
      }
  }
  // ---- tile order and the per-cluster schedule (list scheduling on a simple cost model; 1 unit = one 64-deep k-block
  //      of a 256x256 pair tile ~ 0.44 us): every cluster's list ascends in the global tile order, so the lowest
  //      unfinished tile is always at the head of some cluster's list and its producers (lower tiles) are done or running
  std::vector<ChainTile> tiles;
  chain_tile_order(problems, d.data(), n, tiles);
  const int T = static_cast<int>(tiles.size()), n_cl = L.n_cl;
  std::vector<std::vector<uint32_t>> lists(n_cl);
  double sim_makespan = 0.0;
  {
    const double kEpi = 11.0, kSignal = 3.0, kMinMain = 2.0;
    std::vector<double> mma_free(n_cl, 0.0), epi_free(n_cl, 0.0), epi_prev(n_cl, 0.0), epi_prev2(n_cl, 0.0);
    std::vector<std::vector<double>> done(n);                   // per problem, per pair row: completion time of the LAST N tile
    for (int i = 0; i < n; ++i) done[i].assign(d[i].pairs_m, 0.0);
    double makespan = 0.0;
    for (int t = 0; t < T; ++t) {
      const ChainTile& ct = tiles[t];
      double ready = 0.0;
      for (int k = 0; k < 3; ++k) {
        const int dp = problems[ct.pi].dep[k];
        if (dp < 0) continue;
        if ((problems[ct.pi].dep_all_rows >> k) & 1) { for (double v : done[dp]) ready = std::max(ready, v + kSignal); }
        else ready = std::max(ready, done[dp][ct.tm] + kSignal);
      }
      int best = 0;
      double best_start = 1e300;
      for (int c = 0; c < n_cl; ++c) {
        const double st = std::max(std::max(mma_free[c], epi_prev2[c]), ready);
        if (st < best_start - 1e-9) { best_start = st; best = c; }
      }
      const double main_units = std::max(kMinMain, static_cast<double>((d[ct.pi].K + BK - 1) / BK));
      const double main_end = best_start + main_units;
      const double epi_end = std::max(main_end, epi_free[best]) + kEpi;
      mma_free[best] = main_end;
      epi_free[best] = epi_end;
      epi_prev2[best] = epi_prev[best];                          // two accumulator buffers: tile i+2 waits for epilogue i
      epi_prev[best] = epi_end;
      done[ct.pi][ct.tm] = std::max(done[ct.pi][ct.tm], epi_end);
      makespan = std::max(makespan, epi_end);
      lists[best].push_back((static_cast<uint32_t>(ct.pi) << 22) | (static_cast<uint32_t>(ct.tm) << 11) | static_cast<uint32_t>(ct.tn));
    }
    sim_makespan = makespan;
  }
  int max_len = 0;
  for (int c = 0; c < n_cl; ++c) max_len = std::max(max_len, static_cast<int>(lists[c].size()));
  if (max_len > L.sched_ld) return LINKS_E_RANGE;
  // ---- workspace image
  std::vector<unsigned char> img(L.total, 0);
  memcpy(img.data() + L.off_probs, d.data(), sizeof(GemmProblemDev) * static_cast<size_t>(n));
  uint32_t* sched = reinterpret_cast<uint32_t*>(img.data() + L.off_sched);
  int* cnt = reinterpret_cast<int*>(img.data() + L.off_cnt);
  for (int c = 0; c < n_cl; ++c) {
    cnt[c] = static_cast<int>(lists[c].size());
    for (size_t i = 0; i < lists[c].size(); ++i) sched[static_cast<size_t>(c) * L.sched_ld + i] = lists[c][i];
  }
  cudaError_t e = cudaMemcpyAsync(ws_dev, img.data(), L.total, cudaMemcpyHostToDevice, links_stream(stream));
  if (e != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaStreamSynchronize(links_stream(stream))) != cudaSuccess) return static_cast<int>(e);   // img dies with this frame
  memset(plan, 0, sizeof(*plan));
  unsigned char* w = static_cast<unsigned char*>(ws_dev);
  plan->ws = ws_dev;
  plan->probs = w + L.off_probs;
  plan->sched = w + L.off_sched;
  plan->sched_cnt = w + L.off_cnt;
  plan->counters = w + L.off_counters;
  plan->grid = 2 * n_cl;
  plan->n_counters = L.n_counters - 1;
  plan->sched_ld = L.sched_ld;
  plan->n_problems = n;
  plan->total_tiles = T;
  plan->sim_units = static_cast<float>(sim_makespan);
  {
    double units = 0.0;
    for (int t = 0; t < T; ++t) units += std::max(2.0, static_cast<double>((d[tiles[t].pi].K + BK - 1) / BK));
    plan->ideal_units = static_cast<float>(units / n_cl);
  }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int links_gemm_chain_run(const LinksGemmChainPlan* plan, void* stream) {
  using namespace links;
  if (plan == nullptr || plan->ws == nullptr || plan->grid < 2) return LINKS_E_ARG;
  ChainDev C;
  C.probs = static_cast<const GemmProblemDev*>(plan->probs);
  C.sched = static_cast<const uint32_t*>(plan->sched);
  C.sched_cnt = static_cast<const int*>(plan->sched_cnt);
  C.counters = static_cast<int*>(plan->counters);
  C.exit_cnt = C.counters + plan->n_counters;
  C.n_counters = plan->n_counters;
  C.sched_ld = plan->sched_ld;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = links_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_gemm_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  g_gemm_launches++;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_kernel<true>, C);
  if (le != cudaSuccess) return static_cast<int>(le);
  return links_launch_status();
}

/* GEMM kernel launches (grouped + chain) issued through this library since it was loaded. */
extern "C" __attribute__((visibility("default"))) size_t links_gemm_launch_count(void) { return static_cast<size_t>(links::g_gemm_launches); }

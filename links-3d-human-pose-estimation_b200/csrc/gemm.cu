// Grouped bf16 GEMM with fused epilogue for sm_100a -- persistent, warp-specialised:
//   TMA -> shared memory (128B swizzle, 4-stage ring) -> tcgen05.mma (M128 x N256 x K16, fp32 accumulators in TMEM,
//   two accumulator buffers) -> tcgen05.ld -> fused epilogue -> bf16 slabs staged in shared memory -> TMA stores.
//
// Replaces every nn.Linear(+LeakyReLU, +residual) of reference utils/models_def.py (forward) and its autograd
// backward (dgrad, wgrad); see include/links_b200.h for the epilogue contract.
//
// One CTA per SM walks the 128x256 output tiles of all problems of the group (static round-robin).
//   warp 0      : TMA producer (one elected lane), runs ahead across tile boundaries
//   warp 1      : TMEM allocator + tcgen05.mma issuer (warp-uniform control flow, one elected lane)
//   warps 2..9  : epilogue (TMEM lane group = warp % 4, column half = (warp - 2) / 4); the epilogue of tile i overlaps
//                 the main loop of tile i + 1 through the second accumulator buffer.
// Operands may be K-major (row-major with the contraction dimension contiguous) or MN-major (contraction dimension
// strided): dgrad reads W itself as an MN-major B operand and wgrad reads the row-major activations / gradients as
// MN-major A and B, so no transposed copies of weights or activations are ever written.
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace links {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kAccCols = BN;      // fp32 accumulator columns per buffer
constexpr int kTmemCols = 2 * kAccCols;
constexpr uint32_t kStageBytesA = BM * BK * 2;   // 16 KB
constexpr uint32_t kStageBytesB = BN * BK * 2;   // 32 KB
constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
constexpr uint32_t kSlabBytes = BM * 64 * 2;     // one 128 x 64 bf16 output slab
constexpr uint32_t kOffStage = 0;
constexpr uint32_t kOffOut = kStages * kStageBytes;          // staging: out slab
constexpr uint32_t kOffMid = kOffOut + kSlabBytes;           // staging: mid slab
constexpr uint32_t kOffBar = kOffMid + kSlabBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024 /*align slack*/;

// barriers: full[kStages], empty[kStages], acc_full[2], acc_empty[2]
enum { GB_FULL = 0, GB_EMPTY = kStages, GB_ACCFULL = 2 * kStages, GB_ACCEMPTY = 2 * kStages + 2, GB_COUNT = 2 * kStages + 4 };

struct alignas(64) GemmProblemDev {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmOut;    // bf16 [M, N] row-major, box 64 x 128, 128B swizzle (TMA store)
  CUtensorMap tmMid;
  int M, N, K;
  int tile_begin, tiles_m, tiles_n;
  uint32_t flags;
  int ld_add0, ld_add1, ld_ymask, ld_bits, ld_sign, ld_f32;
  const float* bias;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  const __nv_bfloat16* ymask;
  const uint32_t* bits;
  uint32_t* sign_out;
  float* out_f32;
  int has_out, has_mid;
};

struct GemmGroupDev {
  GemmProblemDev p[LINKS_MAX_GEMM_PROBLEMS];
  int n_problems;
  int total_tiles;
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
// 8 bf16 values of row-major X[m, n .. n+8) with a column guard (n + 8 may exceed N)
__device__ __forceinline__ void load8_guard(const __nv_bfloat16* base, int ld, int m, int n, int N, bool vec, float (&o)[8]) {
  const __nv_bfloat16* p = base + static_cast<size_t>(m) * ld + n;
  if (vec && n + 8 <= N) { load8_bf16(p, o); return; }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = (n + i < N) ? __bfloat162float(p[i]) : 0.f;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct TileCoord { int pi, tm, tn; };
__device__ __forceinline__ TileCoord tile_coord(const GemmGroupDev& G, int tile) {
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < G.n_problems; ++i) if (tile >= G.p[i].tile_begin) pi = i;
  const int local = tile - G.p[pi].tile_begin;
  TileCoord t;
  t.pi = pi;
  t.tn = local / G.p[pi].tiles_m;          // consecutive tiles share the B (weight) slice
  t.tm = local - t.tn * G.p[pi].tiles_m;
  return t;
}

// ----------------------------------------------------------------------------------------------
// Kernel
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) gemm_grouped_kernel(const __grid_constant__ GemmGroupDev G) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // 1024-B aligned (128B swizzle atoms)
  const uint32_t bars = base + kOffBar;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (bars + GB_COUNT * 8 - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bars + (GB_FULL + s) * 8, 1);
      mbar_init(bars + (GB_EMPTY + s) * 8, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + (GB_ACCFULL + s) * 8, 1);
      mbar_init(bars + (GB_ACCEMPTY + s) * 8, kEpiWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(bars + GB_COUNT * 8, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;   // running k-block index over all tiles of this CTA
      for (int tile = blockIdx.x; tile < G.total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(G, tile);
        const GemmProblemDev& P = G.p[tc.pi];
        const int num_kb = (P.K + BK - 1) / BK;
        const bool a_mn = (P.flags & LINKS_GEMM_A_MN) != 0, b_mn = (P.flags & LINKS_GEMM_B_MN) != 0;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages, use = it / kStages;
          mbar_wait(bars + (GB_EMPTY + s) * 8, (use & 1u) ^ 1u);
          const uint32_t full = bars + (GB_FULL + s) * 8;
          mbar_expect_tx(full, kStageBytes);
          const uint32_t sA = base + kOffStage + s * kStageBytes, sB = sA + kStageBytesA;
          if (!a_mn) {
            tma_load_2d(sA, &P.tmA, full, kb * BK, tc.tm * BM);
          } else {
#pragma unroll
            for (int q = 0; q < BM / 64; ++q) tma_load_2d(sA + q * (BK * 128), &P.tmA, full, tc.tm * BM + q * 64, kb * BK);
          }
          if (!b_mn) {
            tma_load_2d(sB, &P.tmB, full, kb * BK, tc.tn * BN);
          } else {
#pragma unroll
            for (int q = 0; q < BN / 64; ++q) tma_load_2d(sB + q * (BK * 128), &P.tmB, full, tc.tn * BN + q * 64, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t bb = sb + kOffBar;
    uint32_t it = 0, lt = 0;   // k-block counter, local tile counter
    for (int tile = blockIdx.x; tile < G.total_tiles; tile += gridDim.x, ++lt) {
      const TileCoord tc = tile_coord(G, tile);
      const GemmProblemDev& P = G.p[tc.pi];
      const int num_kb = (P.K + BK - 1) / BK;
      const bool a_mn = (P.flags & LINKS_GEMM_A_MN) != 0, b_mn = (P.flags & LINKS_GEMM_B_MN) != 0;
      // N actually needed by this tile (multiple of 16): the MMA cost scales with it
      int n_eff = P.N - tc.tn * BN;
      n_eff = n_eff >= BN ? BN : ((n_eff + 15) & ~15);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
                             (static_cast<uint32_t>(n_eff >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
      const uint32_t slot = lt & 1u, acc_use = lt >> 1;
      mbar_wait(bb + (GB_ACCEMPTY + slot) * 8, (acc_use & 1u) ^ 1u);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_addr = tb + slot * kAccCols;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t s = it % kStages, use = it / kStages;
        mbar_wait(bb + (GB_FULL + s) * 8, use & 1u);
        tc_fence_after();
        if (leader) {
          const uint32_t sA = sb + kOffStage + s * kStageBytes, sB = sA + kStageBytesA;
          const uint64_t adesc = a_mn ? make_smem_desc_mn128(sA) : make_smem_desc_k128(sA);
          const uint64_t bdesc = b_mn ? make_smem_desc_mn128(sB) : make_smem_desc_k128(sB);
          // per UMMA_K step: K-major advances 32 B inside the swizzle row, MN-major advances 16 rows of 128 B
          const uint32_t a_step = a_mn ? 128u : 2u, b_step = b_mn ? 128u : 2u;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(d_addr, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(bb + (GB_EMPTY + s) * 8);                          // frees the smem slot when the MMAs retire
          if (kb == num_kb - 1) umma_commit(bb + (GB_ACCFULL + slot) * 8);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps (256 threads) =================
    const int lane_grp = warp & 3;                                    // TMEM lanes 32*lane_grp .. +31
    const int half = (warp - 2) >> 2;                                 // which 32 columns of each 64-column slab
    const int r = lane_grp * 32 + lane;
    const uint32_t sOut = base + kOffOut, sMid = base + kOffMid;
    const uint32_t row_off = static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128);
    const bool store_thread = threadIdx.x == 64;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < G.total_tiles; tile += gridDim.x, ++lt) {
      const TileCoord tc = tile_coord(G, tile);
      const GemmProblemDev& P = G.p[tc.pi];
      const int M = P.M, N = P.N;
      const uint32_t flags = P.flags;
      const float* bias = P.bias;
      const __nv_bfloat16* add0 = P.add0; const __nv_bfloat16* add1 = P.add1; const __nv_bfloat16* ymask = P.ymask;
      const uint32_t* bits = P.bits; uint32_t* sign_out = P.sign_out; float* out_f32 = P.out_f32;
      const bool has_out = P.has_out != 0, has_mid = P.has_mid != 0;
      const bool v0 = (P.ld_add0 & 7) == 0, v1 = (P.ld_add1 & 7) == 0, vy = (P.ld_ymask & 7) == 0;
      const uint32_t slot = lt & 1u, acc_use = lt >> 1;
      const int m = tc.tm * BM + r;
      const bool row_ok = m < M;
      mbar_wait(bars + (GB_ACCFULL + slot) * 8, acc_use & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + slot * kAccCols;
#pragma unroll 1
      for (int sl = 0; sl < BN / 64; ++sl) {
        const int n_slab = tc.tn * BN + sl * 64;
        if (n_slab >= N) break;                                        // uniform over the CTA
        const int n0 = n_slab + half * 32;
        uint32_t acc[32];
        tmem_ld32(t_addr + static_cast<uint32_t>(sl * 64 + half * 32), acc);
        uint32_t bits_word = 0;
        if (bits && row_ok && n0 < N) bits_word = bits[static_cast<size_t>(m) * P.ld_bits + (n0 >> 5)];
        uint32_t sign_word = 0;
        uint32_t po[16], pm[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + g * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]);
          if (bias) {
            if (n + 8 <= N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) if (n + i < N) v[i] += __ldg(bias + n + i);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) sign_word |= (v[i] > 0.f ? 0u : 1u) << (g * 8 + i);
          if (flags & LINKS_EPI_LEAKY_PRE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = links_leaky(v[i]);
          }
          if (flags & LINKS_EPI_RELU_PRE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (add0 && row_ok && n < N) {
            float t[8]; load8_guard(add0, P.ld_add0, m, n, N, v0, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
          }
          if (add1 && row_ok && n < N) {
            float t[8]; load8_guard(add1, P.ld_add1, m, n, N, v1, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
          }
          if (flags & LINKS_EPI_LEAKY_POST) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = links_leaky(v[i]);
          }
          if (ymask && row_ok && n < N) {
            float t[8]; load8_guard(ymask, P.ld_ymask, m, n, N, vy, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= (t[i] > 0.f ? 1.f : 0.01f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) pm[g * 4 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          if (bits) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= ((bits_word >> (g * 8 + i)) & 1u) ? 0.01f : 1.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) po[g * 4 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          if (out_f32 && row_ok && n < N) {
            float* p = out_f32 + static_cast<size_t>(m) * P.ld_f32 + n;
            const bool accum = (flags & LINKS_EPI_ACCUM_F32) != 0;
            if (n + 8 <= N && (P.ld_f32 & 3) == 0) {
              float4 o0 = make_float4(v[0], v[1], v[2], v[3]);
              float4 o1 = make_float4(v[4], v[5], v[6], v[7]);
              if (accum) {
                const float4 a0 = *reinterpret_cast<const float4*>(p);
                const float4 a1 = *reinterpret_cast<const float4*>(p + 4);
                o0.x += a0.x; o0.y += a0.y; o0.z += a0.z; o0.w += a0.w;
                o1.x += a1.x; o1.y += a1.y; o1.z += a1.z; o1.w += a1.w;
              }
              *reinterpret_cast<float4*>(p) = o0;
              *reinterpret_cast<float4*>(p + 4) = o1;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) if (n + i < N) p[i] = accum ? p[i] + v[i] : v[i];
            }
          }
        }
        if (sign_out && row_ok && n0 < N) sign_out[static_cast<size_t>(m) * P.ld_sign + (n0 >> 5)] = sign_word;
        if (has_out || has_mid) {
          // staging slabs are free once the previous slab's TMA stores have finished reading them
          if (store_thread) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          epi_bar_sync();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t c_addr = row_off + ((static_cast<uint32_t>(half * 4 + g) ^ static_cast<uint32_t>(r & 7)) << 4);
            if (has_out) sts128(sOut + c_addr, po[g * 4], po[g * 4 + 1], po[g * 4 + 2], po[g * 4 + 3]);
            if (has_mid) sts128(sMid + c_addr, pm[g * 4], pm[g * 4 + 1], pm[g * 4 + 2], pm[g * 4 + 3]);
          }
          fence_proxy_async_smem();
          epi_bar_sync();
          if (store_thread) {
            if (has_out) tma_store_2d(&P.tmOut, sOut, n_slab, tc.tm * BM);
            if (has_mid) tma_store_2d(&P.tmMid, sMid, n_slab, tc.tm * BM);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bars + (GB_ACCEMPTY + slot) * 8);
    }
    if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // global writes complete before exit
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

// Validate one group and build its device descriptor (tensor maps included).
static int build_group(EncodeTiledFn fn, const LinksGemmProblem* problems, int n_problems, GemmGroupDev& G) {
  memset(&G, 0, sizeof(G));
  int tiles = 0;
  for (int i = 0; i < n_problems; ++i) {
    const LinksGemmProblem& s = problems[i];
    GemmProblemDev& d = G.p[i];
    if (!s.A || !s.B || s.M < 1 || s.N < 1 || s.K < 1) return LINKS_E_ARG;
    const bool a_mn = (s.flags & LINKS_GEMM_A_MN) != 0, b_mn = (s.flags & LINKS_GEMM_B_MN) != 0;
    if (!aligned16(s.A) || !aligned16(s.B) || (s.lda & 7) || (s.ldb & 7)) return LINKS_E_ALIGN;
    if (s.lda < (a_mn ? s.M : s.K) || s.ldb < (b_mn ? s.N : s.K)) return LINKS_E_ALIGN;
    int rc = a_mn ? encode_2d(fn, &d.tmA, s.A, s.K, s.M, s.lda, 64, BK) : encode_2d(fn, &d.tmA, s.A, s.M, s.K, s.lda, BK, BM);
    if (rc) return rc;
    rc = b_mn ? encode_2d(fn, &d.tmB, s.B, s.K, s.N, s.ldb, 64, BK) : encode_2d(fn, &d.tmB, s.B, s.N, s.K, s.ldb, BK, BN);
    if (rc) return rc;
    d.M = s.M; d.N = s.N; d.K = s.K;
    d.tile_begin = tiles;
    d.tiles_m = (s.M + BM - 1) / BM;
    d.tiles_n = (s.N + BN - 1) / BN;
    tiles += d.tiles_m * d.tiles_n;
    d.flags = s.flags;
    if (s.out) {
      if (!aligned16(s.out) || (s.ld_out & 7) || s.ld_out < s.N) return LINKS_E_ALIGN;
      rc = encode_2d(fn, &d.tmOut, s.out, s.M, s.N, s.ld_out, 64, BM);
      if (rc) return rc;
      d.has_out = 1;
    }
    if (s.mid) {
      if (!aligned16(s.mid) || (s.ld_mid & 7) || s.ld_mid < s.N) return LINKS_E_ALIGN;
      rc = encode_2d(fn, &d.tmMid, s.mid, s.M, s.N, s.ld_mid, 64, BM);
      if (rc) return rc;
      d.has_mid = 1;
    }
    if (s.bias && (reinterpret_cast<uintptr_t>(s.bias) & 15u)) return LINKS_E_ALIGN;
    if (s.out_f32 && (reinterpret_cast<uintptr_t>(s.out_f32) & 15u)) return LINKS_E_ALIGN;
    if ((s.add0 && !aligned16(s.add0)) || (s.add1 && !aligned16(s.add1)) || (s.ymask && !aligned16(s.ymask))) return LINKS_E_ALIGN;
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
static int g_num_sms = 0;

static uint64_t hash_bytes(const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
}  // namespace links

extern "C" __attribute__((visibility("default"))) int links_gemm_grouped(const LinksGemmProblem* problems, int n_problems, void* stream) {
  using namespace links;
  if (problems == nullptr || n_problems < 1 || n_problems > LINKS_MAX_GEMM_PROBLEMS) return LINKS_E_ARG;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return LINKS_E_DRIVER;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  if (g_num_sms == 0) {
    cudaError_t e = cudaFuncSetAttribute(gemm_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return static_cast<int>(e);
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return static_cast<int>(e);
    g_num_sms = sms;
  }
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
  const int grid = ce.G.total_tiles < g_num_sms ? ce.G.total_tiles : g_num_sms;
  gemm_grouped_kernel<<<grid, kThreads, kSmemBytes, links_stream(stream)>>>(ce.G);
  return links_launch_status();
}

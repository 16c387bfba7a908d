// Grouped bf16 GEMM with fused epilogue for sm_100a: TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld -> epilogue.
//
// Replaces every nn.Linear(+LeakyReLU, +residual) of reference utils/models_def.py (forward) and
// its autograd backward (dgrad, wgrad); see include/links_b200.h for the epilogue contract.
//
// One CTA computes one 128x128 output tile of one problem of the group.
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM lane group = warp_idx % 4): tcgen05.ld -> fused math -> bf16 tiles staged in the
//                 (by then idle) operand ring in TMA swizzle layout -> TMA bulk tensor stores of the row-major
//                 and the transposed tile (hardware clips the M / N tails)
// kStages x (16 KB A + 16 KB B) shared-memory ring, mbarrier full/empty pipeline; with 3 stages two
// CTAs are resident per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace links {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kTmemCols = 128;    // fp32 accumulator columns (= BN)
constexpr uint32_t kStageBytesA = BM * BK * 2;
constexpr uint32_t kStageBytesB = BN * BK * 2;
constexpr uint32_t kSmemBytes = kStages * (kStageBytesA + kStageBytesB) + 1024 /*align slack*/ + 256 /*barriers*/;

struct alignas(64) GemmProblemDev {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmOut;    // bf16 [M, N] row-major, box 64 x 128, 128B swizzle (TMA store)
  CUtensorMap tmMid;
  CUtensorMap tmOutT;   // bf16 [N, outT_col0 + M], box 64 x 128
  int M, N, K;
  int tile_begin, tiles_n;
  uint32_t flags;
  int vec_ok;
  int ld_add0, ld_add1, ld_ymask, ld_bits, ld_sign, ld_mid, ld_out, ld_outT, outT_col0, ld_f32;
  const float* bias;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  const __nv_bfloat16* ymask;
  const uint32_t* bits;
  uint32_t* sign_out;
  __nv_bfloat16* mid;
  __nv_bfloat16* out;
  __nv_bfloat16* outT;
  float* out_f32;
};

struct GemmGroupDev {
  GemmProblemDev p[LINKS_MAX_GEMM_PROBLEMS];
  int n_problems;
  int total_tiles;
};

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);      // start address      [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (unused, =1)   [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO = 1024 B       [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B       [61,64)
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=128.
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
         (static_cast<uint32_t>(BM >> 4) << 24);
}

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ----------------------------------------------------------------------------------------------
// Epilogue
// ----------------------------------------------------------------------------------------------
// Register-resident copy of the per-problem epilogue parameters (reading them through the kernel-parameter
// reference inside the unrolled epilogue costs an indexed LDC per use).
struct EpiParams {
  int M, N;
  uint32_t flags;
  int ld_add0, ld_add1, ld_ymask, ld_bits, ld_sign, ld_mid, ld_out, ld_outT, outT_col0, ld_f32;
  const float* bias;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  const __nv_bfloat16* ymask;
  const uint32_t* bits;
  uint32_t* sign_out;
  __nv_bfloat16* mid;
  __nv_bfloat16* out;
  __nv_bfloat16* outT;
  float* out_f32;
};

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&o)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  o[0] = bf16_bits_to_f32(q.x & 0xFFFFu); o[1] = bf16_bits_to_f32(q.x >> 16);
  o[2] = bf16_bits_to_f32(q.y & 0xFFFFu); o[3] = bf16_bits_to_f32(q.y >> 16);
  o[4] = bf16_bits_to_f32(q.z & 0xFFFFu); o[5] = bf16_bits_to_f32(q.z >> 16);
  o[6] = bf16_bits_to_f32(q.w & 0xFFFFu); o[7] = bf16_bits_to_f32(q.w >> 16);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

// Vector path (N % 32 == 0, aligned operands): one 32-column chunk of row r of the tile.  bf16 outputs are staged
// in shared memory in the TMA 128B-swizzle layout (row-major tile `C`, `MID`; transposed tile `T`) and written by
// TMA bulk tensor stores afterwards; fp32 outputs and the sign words go straight to global memory.
__device__ __forceinline__ void epilogue_chunk_vec(const EpiParams& E, const uint32_t (&acc)[32], bool row_ok, int m, int n0,
                                                   int r, int c, uint32_t sC, uint32_t sMid, uint32_t sT) {
  uint32_t bits_word = 0;
  if (E.bits && row_ok) bits_word = E.bits[static_cast<size_t>(m) * E.ld_bits + (n0 >> 5)];
  uint32_t sign_word = 0;
  // staging addresses
  const uint32_t row_box = static_cast<uint32_t>(c >> 1) * 16384u + static_cast<uint32_t>(r) * 128u;
  const uint32_t t_lane = static_cast<uint32_t>(r >> 6) * 16384u + static_cast<uint32_t>((r & 7) * 2);
  const uint32_t t_chunk = static_cast<uint32_t>((r & 63) >> 3);
  const uint32_t t_row0 = sT + static_cast<uint32_t>(c) * 4096u + t_lane;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n0 + g * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]);
    if (E.bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(E.bias + n));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(E.bias + n + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sign_word |= (v[i] > 0.f ? 0u : 1u) << (g * 8 + i);
    if (E.flags & LINKS_EPI_LEAKY_PRE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = links_leaky(v[i]);
    }
    if (E.flags & LINKS_EPI_RELU_PRE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (E.add0 && row_ok) {
      float t[8]; load8_bf16(E.add0 + static_cast<size_t>(m) * E.ld_add0 + n, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (E.add1 && row_ok) {
      float t[8]; load8_bf16(E.add1 + static_cast<size_t>(m) * E.ld_add1 + n, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (E.flags & LINKS_EPI_LEAKY_POST) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = links_leaky(v[i]);
    }
    if (E.ymask && row_ok) {
      float t[8]; load8_bf16(E.ymask + static_cast<size_t>(m) * E.ld_ymask + n, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= (t[i] > 0.f ? 1.f : 0.01f);
    }
    const uint32_t c_addr = row_box + ((static_cast<uint32_t>((c & 1) * 4 + g) ^ static_cast<uint32_t>(r & 7)) << 4);
    if (E.mid) sts128(sMid + c_addr, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    if (E.bits) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= ((bits_word >> (g * 8 + i)) & 1u) ? 0.01f : 1.f;
    }
    const uint32_t p01 = pack_bf16x2(v[0], v[1]), p23 = pack_bf16x2(v[2], v[3]);
    const uint32_t p45 = pack_bf16x2(v[4], v[5]), p67 = pack_bf16x2(v[6], v[7]);
    if (E.out) sts128(sC + c_addr, p01, p23, p45, p67);
    if (E.outT) {
      // transposed tile: element (n_local = 32c + 8g + i, m_local = r); row n_local is 128 B, 16B chunks swizzled by n_local & 7 = i
      const uint32_t pk[4] = {p01, p23, p45, p67};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t a = t_row0 + static_cast<uint32_t>((g * 8 + i) * 128) + ((t_chunk ^ static_cast<uint32_t>(i)) << 4);
        sts16(a, static_cast<uint16_t>((i & 1) ? (pk[i >> 1] >> 16) : (pk[i >> 1] & 0xFFFFu)));
      }
    }
    if (E.out_f32 && row_ok) {
      float* p = E.out_f32 + static_cast<size_t>(m) * E.ld_f32 + n;
      float4 o0 = make_float4(v[0], v[1], v[2], v[3]);
      float4 o1 = make_float4(v[4], v[5], v[6], v[7]);
      if (E.flags & LINKS_EPI_ACCUM_F32) {
        const float4 a0 = *reinterpret_cast<const float4*>(p);
        const float4 a1 = *reinterpret_cast<const float4*>(p + 4);
        o0.x += a0.x; o0.y += a0.y; o0.z += a0.z; o0.w += a0.w;
        o1.x += a1.x; o1.y += a1.y; o1.z += a1.z; o1.w += a1.w;
      }
      *reinterpret_cast<float4*>(p) = o0;
      *reinterpret_cast<float4*>(p + 4) = o1;
    }
  }
  if (E.sign_out && row_ok) E.sign_out[static_cast<size_t>(m) * E.ld_sign + (n0 >> 5)] = sign_word;
}

// Scalar path for small / unaligned N (heads, upscale dgrad / wgrad): compact, direct global accesses.
__device__ __noinline__ void epilogue_chunk_scalar(const EpiParams& E, const uint32_t (&acc)[32], int m, int n0) {
  if (m >= E.M) return;
  uint32_t bits_word = 0;
  if (E.bits) bits_word = E.bits[static_cast<size_t>(m) * E.ld_bits + (n0 >> 5)];
  uint32_t sign_word = 0;
#pragma unroll 1
  for (int i = 0; i < 32; ++i) {
    const int n = n0 + i;
    if (n >= E.N) break;
    float v = __uint_as_float(acc[i]);
    if (E.bias) v += __ldg(E.bias + n);
    sign_word |= (v > 0.f ? 0u : 1u) << i;
    if (E.flags & LINKS_EPI_LEAKY_PRE) v = links_leaky(v);
    if (E.flags & LINKS_EPI_RELU_PRE) v = fmaxf(v, 0.f);
    if (E.add0) v += __bfloat162float(E.add0[static_cast<size_t>(m) * E.ld_add0 + n]);
    if (E.add1) v += __bfloat162float(E.add1[static_cast<size_t>(m) * E.ld_add1 + n]);
    if (E.flags & LINKS_EPI_LEAKY_POST) v = links_leaky(v);
    if (E.ymask) v *= (__bfloat162float(E.ymask[static_cast<size_t>(m) * E.ld_ymask + n]) > 0.f ? 1.f : 0.01f);
    if (E.mid) E.mid[static_cast<size_t>(m) * E.ld_mid + n] = __float2bfloat16_rn(v);
    if (E.bits) v *= ((bits_word >> i) & 1u) ? 0.01f : 1.f;
    if (E.out) E.out[static_cast<size_t>(m) * E.ld_out + n] = __float2bfloat16_rn(v);
    if (E.outT) E.outT[static_cast<size_t>(n) * E.ld_outT + E.outT_col0 + m] = __float2bfloat16_rn(v);
    if (E.out_f32) {
      float* p = E.out_f32 + static_cast<size_t>(m) * E.ld_f32 + n;
      *p = (E.flags & LINKS_EPI_ACCUM_F32) ? *p + v : v;
    }
  }
  if (E.sign_out && (n0 >> 5) < E.ld_sign) E.sign_out[static_cast<size_t>(m) * E.ld_sign + (n0 >> 5)] = sign_word;
}

// ----------------------------------------------------------------------------------------------
// Kernel
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) gemm_grouped_kernel(const __grid_constant__ GemmGroupDev G) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // 1024-B aligned tile area (128B swizzle atoms)
  const uint32_t sA = base;
  const uint32_t sB = base + kStages * kStageBytesA;
  const uint32_t bars = sB + kStages * kStageBytesB;     // full[kStages], empty[kStages], tmem_full
  const uint32_t tmem_slot = bars + (2 * kStages + 1) * 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // which problem / tile
  const int tile = blockIdx.x;
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < G.n_problems; ++i) if (tile >= G.p[i].tile_begin) pi = i;
  const GemmProblemDev& P = G.p[pi];
  const int local = tile - P.tile_begin;
  const int tm = local / P.tiles_n;
  const int tn = local - tm * P.tiles_n;
  const int num_kb = (P.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bars + s * 8, 1);
      mbar_init(bars + (kStages + s) * 8, 1);
    }
    mbar_init(bars + 2 * kStages * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&P.tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&P.tmB)) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(bars + (kStages + s) * 8, ph ^ 1u);                 // slot free
        mbar_expect_tx(bars + s * 8, kStageBytesA + kStageBytesB);
        tma_load_2d(sA + s * kStageBytesA, &P.tmA, bars + s * 8, kb * BK, tm * BM);
        tma_load_2d(sB + s * kStageBytesB, &P.tmB, bars + s * 8, kb * BK, tn * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(bars + s * 8, ph);                                  // TMA bytes landed
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(sA + s * kStageBytesA);
        const uint64_t bdesc = make_smem_desc(sB + s * kStageBytesB);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 32 B (16 bf16) inside the 128-B swizzle row: +2 in 16-byte units
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(bars + (kStages + s) * 8);                        // frees the smem slot when MMAs retire
      }
      umma_commit(bars + 2 * kStages * 8);                            // accumulator complete
    }
  } else {
    // ---- epilogue warps (128 threads): thread <-> accumulator row r of the tile
    EpiParams E;
    E.M = P.M; E.N = P.N; E.flags = P.flags;
    E.ld_add0 = P.ld_add0; E.ld_add1 = P.ld_add1; E.ld_ymask = P.ld_ymask; E.ld_bits = P.ld_bits; E.ld_sign = P.ld_sign;
    E.ld_mid = P.ld_mid; E.ld_out = P.ld_out; E.ld_outT = P.ld_outT; E.outT_col0 = P.outT_col0; E.ld_f32 = P.ld_f32;
    E.bias = P.bias; E.add0 = P.add0; E.add1 = P.add1; E.ymask = P.ymask; E.bits = P.bits; E.sign_out = P.sign_out;
    E.mid = P.mid; E.out = P.out; E.outT = P.outT; E.out_f32 = P.out_f32;
    const bool vec = P.vec_ok != 0;
    const int lane_grp = warp & 3;                                    // TMEM lanes 32*lane_grp .. +31
    const int r = lane_grp * 32 + lane;
    const int m = tm * BM + r;
    const bool row_ok = m < E.M;
    // after the accumulator barrier every MMA has retired, so the operand ring is free: reuse it as staging
    const uint32_t sC = base, sMid = base + 32768u, sT = base + 65536u;
    mbar_wait(bars + 2 * kStages * 8, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n0 = tn * BN + c * 32;
      if (n0 >= E.N) break;                                           // warp-uniform
      uint32_t acc[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
      if (vec) epilogue_chunk_vec(E, acc, row_ok, m, n0, r, c, sC, sMid, sT);
      else epilogue_chunk_scalar(E, acc, m, n0);
    }
    tc_fence_before();
    if (vec && (E.out != nullptr || E.outT != nullptr || E.mid != nullptr)) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy smem writes -> async proxy
      asm volatile("bar.sync 1, 128;" ::: "memory");                  // the four epilogue warps only
      if (threadIdx.x == 64) {
        if (E.out != nullptr) {
          tma_store_2d(&P.tmOut, sC, tn * BN, tm * BM);
          tma_store_2d(&P.tmOut, sC + 16384u, tn * BN + 64, tm * BM);
        }
        if (E.mid != nullptr) {
          tma_store_2d(&P.tmMid, sMid, tn * BN, tm * BM);
          tma_store_2d(&P.tmMid, sMid + 16384u, tn * BN + 64, tm * BM);
        }
        if (E.outT != nullptr) {
          tma_store_2d(&P.tmOutT, sT, E.outT_col0 + tm * BM, tn * BN);
          tma_store_2d(&P.tmOutT, sT + 16384u, E.outT_col0 + tm * BM + 64, tn * BN);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must stay valid until read
      }
    }
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

// bf16 matrix [rows, cols] with leading dimension ld (elements); box = 64 (K) x 128 (rows), 128B swizzle.
static int encode_operand(EncodeTiledFn fn, CUtensorMap* map, const void* ptr, int rows, int cols, int ld) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : LINKS_E_DRIVER;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace links

namespace links {
// Validate one group and build its device descriptor (tensor maps included).
static int build_group(EncodeTiledFn fn, const LinksGemmProblem* problems, int n_problems, GemmGroupDev& G) {
  memset(&G, 0, sizeof(G));
  int tiles = 0;
  for (int i = 0; i < n_problems; ++i) {
    const LinksGemmProblem& s = problems[i];
    GemmProblemDev& d = G.p[i];
    if (!s.A || !s.B || s.M < 1 || s.N < 1 || s.K < 1) return LINKS_E_ARG;
    if (!aligned16(s.A) || !aligned16(s.B) || (s.lda & 7) || (s.ldb & 7) || s.lda < s.K || s.ldb < s.K) return LINKS_E_ALIGN;
    int rc = encode_operand(fn, &d.tmA, s.A, s.M, s.K, s.lda);
    if (rc) return rc;
    rc = encode_operand(fn, &d.tmB, s.B, s.N, s.K, s.ldb);
    if (rc) return rc;
    d.M = s.M; d.N = s.N; d.K = s.K;
    d.tile_begin = tiles;
    d.tiles_n = (s.N + BN - 1) / BN;
    tiles += ((s.M + BM - 1) / BM) * d.tiles_n;
    d.flags = s.flags;
    bool vec = (s.N % 32) == 0;
    auto chk = [&](const void* p, int ld, int mult) {
      if (p && (!aligned16(p) || (ld % mult) != 0)) vec = false;
    };
    chk(s.bias, 4, 4); chk(s.add0, s.ld_add0, 8); chk(s.add1, s.ld_add1, 8); chk(s.ymask, s.ld_ymask, 8);
    chk(s.mid, s.ld_mid, 8); chk(s.out, s.ld_out, 8); chk(s.out_f32, s.ld_f32, 4);
    d.vec_ok = vec ? 1 : 0;
    if (vec) {
      if (s.out) { rc = encode_operand(fn, &d.tmOut, s.out, s.M, s.N, s.ld_out); if (rc) return rc; }
      if (s.mid) { rc = encode_operand(fn, &d.tmMid, s.mid, s.M, s.N, s.ld_mid); if (rc) return rc; }
      if (s.outT) {
        // TMA stores need a 16-byte aligned start: the column offset of the transposed tile must be a multiple of 8
        if (!aligned16(s.outT) || (s.ld_outT & 7) || (s.outT_col0 & 7) || s.ld_outT < s.outT_col0 + s.M) return LINKS_E_ALIGN;
        rc = encode_operand(fn, &d.tmOutT, s.outT, s.N, s.outT_col0 + s.M, s.ld_outT);
        if (rc) return rc;
      }
    }
    d.ld_add0 = s.ld_add0; d.ld_add1 = s.ld_add1; d.ld_ymask = s.ld_ymask; d.ld_bits = s.ld_bits;
    d.ld_sign = s.ld_sign; d.ld_mid = s.ld_mid; d.ld_out = s.ld_out; d.ld_outT = s.ld_outT;
    d.outT_col0 = s.outT_col0; d.ld_f32 = s.ld_f32;
    d.bias = s.bias;
    d.add0 = static_cast<const __nv_bfloat16*>(s.add0);
    d.add1 = static_cast<const __nv_bfloat16*>(s.add1);
    d.ymask = static_cast<const __nv_bfloat16*>(s.ymask);
    d.bits = s.bits;
    d.sign_out = s.sign_out;
    d.mid = static_cast<__nv_bfloat16*>(s.mid);
    d.out = static_cast<__nv_bfloat16*>(s.out);
    d.outT = static_cast<__nv_bfloat16*>(s.outT);
    d.out_f32 = s.out_f32;
    if (s.bits && s.ld_bits < (s.N + 31) / 32) return LINKS_E_RANGE;
    if (s.sign_out && s.ld_sign < (s.N + 31) / 32) return LINKS_E_RANGE;
  }
  G.n_problems = n_problems;
  G.total_tiles = tiles;
  return 0;
}

// Launch descriptors are pure functions of the problem array; cache them so that repeated (eager) launches of
// the same plan do not pay ~5 cuTensorMapEncodeTiled calls per problem.  Direct-mapped, full-key compare.
struct CacheEntry {
  int n;
  LinksGemmProblem key[LINKS_MAX_GEMM_PROBLEMS];
  GemmGroupDev G;
};
constexpr int kCacheSize = 1024;
static CacheEntry* g_cache = nullptr;
static std::mutex g_cache_mu;

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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const size_t key_bytes = sizeof(LinksGemmProblem) * static_cast<size_t>(n_problems);
  const uint64_t h = hash_bytes(problems, key_bytes);
  std::lock_guard<std::mutex> lock(g_cache_mu);
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
  gemm_grouped_kernel<<<ce.G.total_tiles, kThreads, kSmemBytes, links_stream(stream)>>>(ce.G);
  return links_launch_status();
}

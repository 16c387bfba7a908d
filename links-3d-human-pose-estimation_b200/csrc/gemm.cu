// Grouped bf16 GEMM with fused epilogue for sm_100a: TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld -> epilogue.
//
// Replaces every nn.Linear(+LeakyReLU, +residual) of reference utils/models_def.py (forward) and
// its autograd backward (dgrad, wgrad); see include/links_b200.h for the epilogue contract.
//
// One CTA computes one 128x128 output tile of one problem of the group.
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM lane group = warp_idx % 4)
// kStages x (16 KB A + 16 KB B) shared-memory ring, mbarrier full/empty pipeline; with 3 stages two
// CTAs are resident per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include "common.cuh"

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

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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
// Epilogue for one 32-column chunk held by one thread (row m, columns n0..n0+31)
// ----------------------------------------------------------------------------------------------
template <bool kVec>
__device__ __forceinline__ void epilogue_chunk(const GemmProblemDev& P, const uint32_t (&acc)[32], int m, int n0) {
  const bool row_ok = m < P.M;
  if (!row_ok) return;
  uint32_t bits_word = 0;
  if (P.bits) bits_word = P.bits[static_cast<size_t>(m) * P.ld_bits + (n0 >> 5)];
  uint32_t sign_word = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n0 + g * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]);
    if (!kVec && n >= P.N) break;
    if (P.bias) {
      if (kVec) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(P.bias + n));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + n + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (n + i < P.N) v[i] += __ldg(P.bias + n + i);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sign_word |= (v[i] > 0.f ? 0u : 1u) << (g * 8 + i);
    if (P.flags & LINKS_EPI_LEAKY_PRE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = links_leaky(v[i]);
    }
    if (P.flags & LINKS_EPI_RELU_PRE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    // additive bf16 terms and derivative mask share one loader
    auto load8 = [&](const __nv_bfloat16* base, int ld, float (&o)[8]) {
      const __nv_bfloat16* p = base + static_cast<size_t>(m) * ld + n;
      if (kVec) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
        o[0] = bf16_bits_to_f32(q.x & 0xFFFFu); o[1] = bf16_bits_to_f32(q.x >> 16);
        o[2] = bf16_bits_to_f32(q.y & 0xFFFFu); o[3] = bf16_bits_to_f32(q.y >> 16);
        o[4] = bf16_bits_to_f32(q.z & 0xFFFFu); o[5] = bf16_bits_to_f32(q.z >> 16);
        o[6] = bf16_bits_to_f32(q.w & 0xFFFFu); o[7] = bf16_bits_to_f32(q.w >> 16);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (n + i < P.N) ? __bfloat162float(p[i]) : 0.f;
      }
    };
    if (P.add0) {
      float t[8]; load8(P.add0, P.ld_add0, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (P.add1) {
      float t[8]; load8(P.add1, P.ld_add1, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (P.flags & LINKS_EPI_LEAKY_POST) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = links_leaky(v[i]);
    }
    if (P.ymask) {
      float t[8]; load8(P.ymask, P.ld_ymask, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= (t[i] > 0.f ? 1.f : 0.01f);
    }
    auto store8_bf16 = [&](__nv_bfloat16* base, int ld) {
      __nv_bfloat16* p = base + static_cast<size_t>(m) * ld + n;
      if (kVec) {
        uint4 q;
        q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
        q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = q;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (n + i < P.N) p[i] = __float2bfloat16_rn(v[i]);
      }
    };
    if (P.mid) store8_bf16(P.mid, P.ld_mid);
    if (P.bits) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= ((bits_word >> (g * 8 + i)) & 1u) ? 0.01f : 1.f;
    }
    if (P.out) store8_bf16(P.out, P.ld_out);
    if (P.outT) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (kVec || n + i < P.N)
          P.outT[static_cast<size_t>(n + i) * P.ld_outT + P.outT_col0 + m] = __float2bfloat16_rn(v[i]);
      }
    }
    if (P.out_f32) {
      float* p = P.out_f32 + static_cast<size_t>(m) * P.ld_f32 + n;
      if (kVec) {
        float4 o0 = make_float4(v[0], v[1], v[2], v[3]);
        float4 o1 = make_float4(v[4], v[5], v[6], v[7]);
        if (P.flags & LINKS_EPI_ACCUM_F32) {
          const float4 a0 = *reinterpret_cast<const float4*>(p);
          const float4 a1 = *reinterpret_cast<const float4*>(p + 4);
          o0.x += a0.x; o0.y += a0.y; o0.z += a0.z; o0.w += a0.w;
          o1.x += a1.x; o1.y += a1.y; o1.z += a1.z; o1.w += a1.w;
        }
        *reinterpret_cast<float4*>(p) = o0;
        *reinterpret_cast<float4*>(p + 4) = o1;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (n + i < P.N) p[i] = (P.flags & LINKS_EPI_ACCUM_F32) ? p[i] + v[i] : v[i];
        }
      }
    }
  }
  if (P.sign_out && (n0 >> 5) < P.ld_sign) P.sign_out[static_cast<size_t>(m) * P.ld_sign + (n0 >> 5)] = sign_word;
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
    const int lane_grp = warp & 3;                                    // TMEM lanes 32*lane_grp .. +31
    mbar_wait(bars + 2 * kStages * 8, 0);
    tc_fence_after();
    const int m = tm * BM + lane_grp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n0 = tn * BN + c * 32;
      if (n0 >= P.N) break;                                           // warp-uniform
      uint32_t acc[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
      if (P.vec_ok && n0 + 32 <= P.N) epilogue_chunk<true>(P, acc, m, n0);
      else epilogue_chunk<false>(P, acc, m, n0);
    }
    tc_fence_before();
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
  GemmGroupDev G;
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
  gemm_grouped_kernel<<<tiles, kThreads, kSmemBytes, links_stream(stream)>>>(G);
  return links_launch_status();
}

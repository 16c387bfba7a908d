// Normalising flow on the 5th-generation tensor cores (sm_100a): FrEIA SequenceINN of AllInOneBlock(subnet_fc,
// permute_soft=True), semantics restated in oracle/flow.py (reference call sites train_leg_torso_lifter.py:134-136,
// 207-214, train_left_right_lifter.py:131-133,334-340, train_full_pose_norm_flow.py:75-90).
//
// One CTA carries a tile of 128 rows through ALL coupling blocks.  Per block the subnet
//     a = W2 . relu(W1 . x1 + b1)            (c1 -> 1024 -> 2*c2)
// runs as a chain of two tcgen05 GEMMs per 64-wide hidden chunk: D1 = A1 . B1^T into TMEM, the compute warps turn
// D1 into the bf16 operand H of the second GEMM (ReLU fused), D2 += H . B2^T.  The [rows, 1024] hidden layer never
// leaves the SM.  Precision: every operand is split x = hi + lo into two bf16 values and each product is evaluated
// as hi*hi + lo*hi + hi*lo with fp32 accumulation (the classic bf16x3 scheme, ~2^-17 relative), because the flow's
// log-det / exp terms are the precision-critical part of the loss; the bias rides along as two extra K columns.
// The backward pass is reversible: block inputs are reconstructed with the inverse coupling (whose subnet
// evaluation the backward needs anyway), so no activations are stored; the input gradient of the subnet is a second
// chain dx1 = ((da . W2) * relu') . W1 with the ReLU mask recomputed in TMEM.
//
//   warp 0      : producer -- one cp.async.bulk per chunk from the pre-swizzled packed weights (3-stage ring)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one lane)
//   warps 2..5  : row threads (thread <-> row): hold the row state in registers, do the coupling / soft permutation
//                 / log-det math, build the A operands, and convert half of every hidden chunk
//   warps 6..9  : convert the other half of every hidden chunk
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#include "flow.cuh"

namespace links {

constexpr int kTcRows = 128;
constexpr int kTcHc = 64;
constexpr int kTcChunks = kFlowHidden / kTcHc;
constexpr int kTcStages = 3;
constexpr int kTcThreads = 320;
constexpr int kTcPermLd = 36;   // padded row length of the fp32 permutation tables (16-byte aligned rows)

__host__ __device__ constexpr int tc_rup(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ constexpr int tc_n2p(int C) { return tc_rup(2 * flow_c2(C), 16); }
__host__ __device__ constexpr int tc_n4p(int C) { return tc_rup(flow_c1(C), 16); }
__host__ __device__ constexpr int tc_k3slabs(int C) { return (3 * 2 * flow_c2(C) + 63) / 64; }
__host__ __device__ constexpr int tc_fw_bytes(int C) { return kTcHc * 128 + 2 * tc_n2p(C) * 128; }
__host__ __device__ constexpr int tc_bw_bytes(int C) {
  return kTcHc * 128 + tc_k3slabs(C) * kTcHc * 128 + 2 * tc_n4p(C) * 128;
}
// fp32 tail of a block: b2[2c2] g[C] off[C] logg[1] dgfac[C] pad -> WT[C][36] WiT[C][36] Wrow[C][36]
// (dgfac = d g / d global_scale = 0.1 * sigmoid(0.5 * global_scale), used when the flow itself is trained)
__host__ __device__ constexpr int tc_tail_hdr(int C) { return tc_rup(2 * flow_c2(C) + 3 * C + 1, 4); }
__host__ __device__ constexpr int tc_tail_floats(int C) { return tc_tail_hdr(C) + 3 * C * kTcPermLd; }
__host__ __device__ constexpr size_t tc_block_bytes(int C) {
  return static_cast<size_t>(kTcChunks) * (tc_fw_bytes(C) + tc_bw_bytes(C)) + static_cast<size_t>(tc_tail_floats(C)) * 4;
}
constexpr int kTcStageBytes = 32768;
static_assert(tc_bw_bytes(34) <= kTcStageBytes && tc_fw_bytes(34) <= kTcStageBytes, "stage too small");

// byte offset of element (r, k) of a K-major bf16 tile with 64-element (128 B) rows in the 128B-swizzle layout
__host__ __device__ constexpr uint32_t tc_sw128(int r, int k) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// ----------------------------------------------------------------------------------------------
// Packing: FrEIA tensors -> swizzled bf16 hi/lo operand images + fp32 tables.  grid = (n_blocks), block = 256
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// grid = (n_blocks, kFlowPackSplit): the element loops of one coupling block are spread over the CTAs of its grid row
// (a single CTA per block needed ~0.5 ms -- a third of the B = 256 flow training step, which re-packs every step);
// slice 0 also writes the fp32 tail.
__global__ void flow_tc_pack_kernel(const FlowPackArgs A, unsigned char* tc_packed) {
  const int k = blockIdx.x;
  const int e_first = blockIdx.y * blockDim.x + threadIdx.x, e_step = blockDim.x * gridDim.y;
  const int C = A.C, c1 = flow_c1(C), c2 = flow_c2(C);
  const int n2p = tc_n2p(C), n4p = tc_n4p(C), s3 = tc_k3slabs(C);
  const int fw = tc_fw_bytes(C), bw = tc_bw_bytes(C);
  unsigned char* P = tc_packed + static_cast<size_t>(k) * tc_block_bytes(C);
  const float* w0 = A.w0[k];   // [1024, c1]
  const float* b0 = A.b0[k];   // [1024]
  const float* w2 = A.w2[k];   // [2c2, 1024]
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  // ---- forward sets
  const int fw_elems = fw / 2;
  for (int e = e_first; e < kTcChunks * fw_elems; e += e_step) {
    const int c = e / fw_elems;
    int i = e - c * fw_elems;
    unsigned char* base = P + static_cast<size_t>(c) * fw;
    __nv_bfloat16 val = zero;
    uint32_t off;
    if (i < kTcHc * 64) {                           // B1: row = hidden unit, K = [W1hi, W1hi, W1lo, b1hi, b1lo, 0..]
      const int r = i >> 6, kk = i & 63;
      const int h = c * kTcHc + r;
      __nv_bfloat16 hi, lo;
      if (kk < 3 * c1) {
        tc_split(w0[h * c1 + (kk % c1)], hi, lo);
        val = (kk < 2 * c1) ? hi : lo;
      } else if (kk < 3 * c1 + 2) {
        tc_split(b0[h], hi, lo);
        val = (kk == 3 * c1) ? hi : lo;
      }
      off = tc_sw128(r, kk);
    } else {                                        // B2 hi / lo: row = output o, K = hidden within the chunk
      i -= kTcHc * 64;
      const int part = i / (n2p * 64);
      i -= part * n2p * 64;
      const int r = i >> 6, kk = i & 63;
      if (r < 2 * c2) {
        __nv_bfloat16 hi, lo;
        tc_split(w2[static_cast<size_t>(r) * kFlowHidden + c * kTcHc + kk], hi, lo);
        val = part == 0 ? hi : lo;
      }
      off = kTcHc * 128 + part * n2p * 128 + tc_sw128(r, kk);
    }
    *reinterpret_cast<__nv_bfloat16*>(base + off) = val;
  }
  // ---- backward sets
  unsigned char* PB = P + static_cast<size_t>(kTcChunks) * fw;
  const int bw_elems = bw / 2;
  for (int e = e_first; e < kTcChunks * bw_elems; e += e_step) {
    const int c = e / bw_elems;
    int i = e - c * bw_elems;
    unsigned char* base = PB + static_cast<size_t>(c) * bw;
    __nv_bfloat16 val = zero;
    uint32_t off;
    if (i < kTcHc * 64) {                           // B1 again (ReLU mask recompute)
      const int r = i >> 6, kk = i & 63;
      const int h = c * kTcHc + r;
      __nv_bfloat16 hi, lo;
      if (kk < 3 * c1) {
        tc_split(w0[h * c1 + (kk % c1)], hi, lo);
        val = (kk < 2 * c1) ? hi : lo;
      } else if (kk < 3 * c1 + 2) {
        tc_split(b0[h], hi, lo);
        val = (kk == 3 * c1) ? hi : lo;
      }
      off = tc_sw128(r, kk);
    } else if (i < kTcHc * 64 + s3 * kTcHc * 64) {  // B3: row = hidden unit, K = [W2hi[:,h], W2hi[:,h], W2lo[:,h], 0..]
      i -= kTcHc * 64;
      const int slab = i / (kTcHc * 64);
      i -= slab * kTcHc * 64;
      const int r = i >> 6, kk = slab * 64 + (i & 63);
      const int h = c * kTcHc + r;
      if (kk < 3 * 2 * c2) {
        __nv_bfloat16 hi, lo;
        tc_split(w2[static_cast<size_t>(kk % (2 * c2)) * kFlowHidden + h], hi, lo);
        val = (kk < 2 * 2 * c2) ? hi : lo;
      }
      off = kTcHc * 128 + slab * kTcHc * 128 + tc_sw128(r, i & 63);
    } else {                                        // B4 hi / lo: row = input i of W1, K = hidden within the chunk
      i -= kTcHc * 64 + s3 * kTcHc * 64;
      const int part = i / (n4p * 64);
      i -= part * n4p * 64;
      const int r = i >> 6, kk = i & 63;
      if (r < c1) {
        __nv_bfloat16 hi, lo;
        tc_split(w0[(c * kTcHc + kk) * c1 + r], hi, lo);
        val = part == 0 ? hi : lo;
      }
      off = kTcHc * 128 + s3 * kTcHc * 128 + part * n4p * 128 + tc_sw128(r, kk);
    }
    *reinterpret_cast<__nv_bfloat16*>(base + off) = val;
  }
  // ---- fp32 tail
  if (blockIdx.y != 0) return;
  float* T = reinterpret_cast<float*>(P + static_cast<size_t>(kTcChunks) * (fw + bw));
  float* b2 = T;
  float* g = b2 + 2 * c2;
  float* of = g + C;
  float* logg = of + C;
  float* dgfac = logg + 1;
  float* WT = T + tc_tail_hdr(C);
  float* WiT = WT + C * kTcPermLd;
  float* Wrow = WiT + C * kTcPermLd;
  for (int i = threadIdx.x; i < 2 * c2; i += blockDim.x) b2[i] = A.b2[k][i];
  __shared__ float s_logg[256];
  float lg = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float x = A.gs[k][i];
    const float sp = (0.5f * x > 20.f) ? x : 2.f * log1pf(expf(0.5f * x));   // softplus(beta = 0.5, threshold = 20)
    const float gg = 0.1f * sp;
    g[i] = gg;
    of[i] = A.go[k][i];
    lg += logf(gg);
    dgfac[i] = 0.1f / (1.f + expf(-0.5f * x));
  }
  for (int e = threadIdx.x; e < C * kTcPermLd; e += blockDim.x) {
    const int i = e / kTcPermLd, o = e - i * kTcPermLd;
    WT[e] = o < C ? A.wp[k][o * C + i] : 0.f;       // WT[i][o]  = w_perm[o][i]
    WiT[e] = o < C ? A.wpi[k][o * C + i] : 0.f;     // WiT[i][o] = w_perm_inv[o][i]
    Wrow[e] = o < C ? A.wp[k][i * C + o] : 0.f;     // Wrow[o'][i'] = w_perm[o'][i']
  }
  s_logg[threadIdx.x] = lg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x); ++i) t += s_logg[i];
    *logg = t;
  }
}

// ----------------------------------------------------------------------------------------------
// Kernel
// ----------------------------------------------------------------------------------------------
struct FlowTcArgs {
  const unsigned char* packed;
  const float* x;
  const float* noise;
  float* out;
  float* ld;
  float* nll_sum;
  float scale;
  const float* gz;
  const float* gld;
  // flow training (NLL mode, all optional): per-block exports for the parameter-gradient GEMMs and row-reduced
  // gradients of the global affine.  ex_x1 / ex_dsub: bf16 [n_blocks][M][64] (only the first c1 / 2*c2 columns are
  // written; the caller keeps the padding zero), d_gscale / d_goffset: fp32 [n_blocks][C], accumulated atomically.
  void* ex_x1;
  void* ex_dsub;
  float* d_gscale;
  float* d_goffset;
  // NLL mode, optional: [M][n_blocks][C + 2*c2] floats.  When given, the forward pass stores every block's input and its
  // scaled subnet output there and the backward pass reads them back instead of reconstructing them with the inverse
  // coupling: 2 GEMM chains per block instead of 3 (the kernel is bound by the length of its dependent chain sequence).
  float* stash;
  int M, n_blocks;
};

// shared memory map (bytes, from a 1024-aligned base)
constexpr uint32_t kTcOffA = 0;                                   // A operands: 3 slabs x 16 KB (slab 2: x1 of a bwd chain)
constexpr uint32_t kTcOffH = 49152;                               // H[2] x (hi 16 KB + lo 16 KB)
constexpr uint32_t kTcOffW = kTcOffH + 65536;                     // weight stages
constexpr uint32_t kTcOffY = kTcOffW + kTcStages * kTcStageBytes; // row scratch [128][35] fp32
constexpr uint32_t kTcOffBar = kTcOffY + kTcRows * 35 * 4;        // mbarriers
constexpr uint32_t kTcSmemBytes = kTcOffBar + 128 + 1024;         // + alignment slack

// barrier indices
enum { TCB_WFULL = 0, TCB_WEMPTY = kTcStages, TCB_AFULL = 2 * kTcStages, TCB_D1FULL, TCB_D1FULL1, TCB_HFULL,
       TCB_HFULL1, TCB_D2FULL, TCB_COUNT };

constexpr uint32_t kTcColD1 = 0, kTcColD3 = 128, kTcColD2 = 256, kTcTmemCols = 512;

// The GEMM chains of one kernel, in execution order: chain i works on block k, forward (a = subnet(x1)) or
// backward (dx1 = subnet_vjp) flavour.
template <int MODE>
__device__ __forceinline__ int tc_num_chains(int nb, bool stash) {
  return MODE == FLOW_FWD || MODE == FLOW_REV ? nb : (MODE == FLOW_SAMPLE || stash ? 2 * nb : 3 * nb);
}
template <int MODE>
__device__ __forceinline__ void tc_chain_info(int i, int nb, bool stash, int& k, bool& bwd) {
  bwd = false;
  if (MODE == FLOW_FWD) { k = i; }
  else if (MODE == FLOW_REV) { k = nb - 1 - i; }
  else if (MODE == FLOW_SAMPLE) { k = i < nb ? i : 2 * nb - 1 - i; }
  else {
    if (i < nb) { k = i; }
    else if (stash) { k = 2 * nb - 1 - i; bwd = true; }
    else { const int j = i - nb; k = nb - 1 - (j >> 1); bwd = (j & 1) != 0; }
  }
}

__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct TcTail {
  const float* b2; const float* g; const float* off; const float* dgfac; const float* WT; const float* WiT; const float* Wrow;
  float logg;
};
template <int C>
__device__ __forceinline__ TcTail tc_tail(const unsigned char* packed, int k) {
  const float* T = reinterpret_cast<const float*>(packed + static_cast<size_t>(k) * tc_block_bytes(C) +
                                                  static_cast<size_t>(kTcChunks) * (tc_fw_bytes(C) + tc_bw_bytes(C)));
  TcTail t;
  t.b2 = T; t.g = T + 2 * flow_c2(C); t.off = t.g + C; t.logg = __ldg(t.off + C); t.dgfac = t.off + C + 1;
  t.WT = T + tc_tail_hdr(C); t.WiT = t.WT + C * kTcPermLd; t.Wrow = t.WiT + C * kTcPermLd;
  return t;
}

__device__ __forceinline__ uint32_t tc_pack_hi_lo(float v0, float v1, uint32_t& lo_out) {
  const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
  const float2 hf = __bfloat1622float2(hi);
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
  lo_out = *reinterpret_cast<const uint32_t*>(&lo);
  return *reinterpret_cast<const uint32_t*>(&hi);
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

// Conversion side of one chain (all 256 compute threads): per chunk D_first (TMEM) -> H hi/lo (smem operand).
template <bool BWD>
__device__ __forceinline__ void tc_chain_convert(uint32_t sbase, uint32_t bars, uint32_t tmem_base, int lane_grp, int lane,
                                                 int half, uint32_t& n_d1a, uint32_t& n_d1b) {
  const int row = lane_grp * 32 + lane;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
  const uint32_t row_off = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll 1
  for (int c = 0; c < kTcChunks; ++c) {
    const int slot = c & 1;
    mbar_wait(bars + (TCB_D1FULL + slot) * 8, (slot ? n_d1b : n_d1a) & 1u);
    if (slot) n_d1b++; else n_d1a++;
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(lane_addr + kTcColD1 + slot * 64 + half * 32, v);
    if (BWD) {
      uint32_t g[32];
      tmem_ld32(lane_addr + kTcColD3 + slot * 64 + half * 32, g);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (__uint_as_float(v[i]) > 0.f) ? g[i] : 0u;
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), 0.f));
    }
    const uint32_t hbase = sbase + kTcOffH + slot * 32768u + row_off;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        hi[e] = tc_pack_hi_lo(__uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]), lo[e]);
      const uint32_t ch = static_cast<uint32_t>(((half * 4 + j) ^ (row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hbase + ch), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hbase + 16384u + ch), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
    tc_fence_before();
    fence_proxy_async_smem();
    mbar_arrive(bars + (TCB_HFULL + slot) * 8);
  }
}

// Build an A operand row: K = [hi(v), lo(v), hi(v), (1, 1 if BIAS), 0 ...] over NSLAB 64-wide slabs.
template <int NV, int NSLAB, bool BIAS>
__device__ __forceinline__ void tc_build_A(uint32_t sA, int row, const float (&v)[NV]) {
  const uint32_t row_off = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
  uint32_t hi[(NV + 1) / 2 * 2], lo[(NV + 1) / 2 * 2];   // per element 16-bit patterns kept in 32-bit regs
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v[i]);
    const __nv_bfloat16 l = __float2bfloat16_rn(v[i] - __bfloat162float(h));
    hi[i] = *reinterpret_cast<const unsigned short*>(&h);
    lo[i] = *reinterpret_cast<const unsigned short*>(&l);
  }
  auto elem = [&](int k) -> uint32_t {
    if (k < NV) return hi[k];
    if (k < 2 * NV) return lo[k - NV];
    if (k < 3 * NV) return hi[k - 2 * NV];
    if (BIAS && k < 3 * NV + 2) return 0x3F80u;   // bf16 1.0
    return 0u;
  };
#pragma unroll
  for (int s = 0; s < NSLAB; ++s) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = s * 64 + j * 8 + 2 * e;
        w[e] = elem(k) | (elem(k + 1) << 16);
      }
      const uint32_t addr = sA + s * 16384u + row_off + static_cast<uint32_t>((j ^ (row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
  }
}

template <int C, int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) flow_tc_kernel(const FlowTcArgs A) {
  constexpr int c1 = flow_c1(C), c2 = flow_c2(C);
  constexpr int n2p = tc_n2p(C), n4p = tc_n4p(C), s3 = tc_k3slabs(C);
  constexpr int fw = tc_fw_bytes(C), bw = tc_bw_bytes(C);
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* gbase = tc_smem_raw + (sbase - raw);
  const uint32_t bars = sbase + kTcOffBar;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + kTcOffBar + TCB_COUNT * 8);
  float* ys_all = reinterpret_cast<float*>(gbase + kTcOffY);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = A.n_blocks;
  const bool stash = MODE == FLOW_NLL_FWDBWD && A.stash != nullptr;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(bars + (TCB_WFULL + s) * 8, 1); mbar_init(bars + (TCB_WEMPTY + s) * 8, 1); }
    mbar_init(bars + TCB_AFULL * 8, 128);
    mbar_init(bars + TCB_D1FULL * 8, 1); mbar_init(bars + TCB_D1FULL1 * 8, 1);
    mbar_init(bars + TCB_HFULL * 8, 256); mbar_init(bars + TCB_HFULL1 * 8, 256);
    mbar_init(bars + TCB_D2FULL * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(bars + TCB_COUNT * 8, kTcTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= producer =================
    if (lane == 0) {
      uint32_t g = 0;   // running chunk index
      const int nch = tc_num_chains<MODE>(nb, stash);
#pragma unroll 1
      for (int ci = 0; ci < nch; ++ci) {
        int k; bool bwd;
        tc_chain_info<MODE>(ci, nb, stash, k, bwd);
        const unsigned char* src = A.packed + static_cast<size_t>(k) * tc_block_bytes(C) +
                                   (bwd ? static_cast<size_t>(kTcChunks) * fw : 0);
        const uint32_t bytes = bwd ? bw : fw;
#pragma unroll 1
        for (int c = 0; c < kTcChunks; ++c, ++g) {
          const uint32_t s = g % kTcStages, use = g / kTcStages;
          mbar_wait(bars + (TCB_WEMPTY + s) * 8, (use & 1u) ^ 1u);
          mbar_expect_tx(bars + (TCB_WFULL + s) * 8, bytes);
          tma_load_1d(sbase + kTcOffW + s * kTcStageBytes, src + static_cast<size_t>(c) * bytes, bytes,
                      bars + (TCB_WFULL + s) * 8);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers; one elected
    // lane issues the tcgen05 instructions.
    const bool leader = tc_elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0);
    const uint32_t bb = sb + kTcOffBar;
    constexpr uint32_t idesc64 = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc2 = make_idesc_bf16(128, 2 * n2p);   // [B2hi ; B2lo] stacked along N
    constexpr uint32_t idesc2h = make_idesc_bf16(128, n2p);
    constexpr uint32_t idesc4 = make_idesc_bf16(128, 2 * n4p);
    constexpr uint32_t idesc4h = make_idesc_bf16(128, n4p);
    uint32_t g0 = 0, n_a = 0, n_h0 = 0, n_h1 = 0;
    const int nch = tc_num_chains<MODE>(nb, stash);
    const uint64_t a_slab0 = make_smem_desc_k128(sb + kTcOffA);
    const uint64_t a_slab1 = make_smem_desc_k128(sb + kTcOffA + 16384u);
    const uint64_t a_slab2 = make_smem_desc_k128(sb + kTcOffA + 2 * 16384u);
#pragma unroll 1
    for (int ci = 0; ci < nch; ++ci) {
      int kblk; bool bwd;
      tc_chain_info<MODE>(ci, nb, stash, kblk, bwd);
      mbar_wait(bb + TCB_AFULL * 8, n_a & 1u);
      n_a++;
      tc_fence_after();
      // first GEMM(s) of chunk c: D1[slot] = x1-operand . B1^T  (+ D3[slot] = da-operand . B3^T in a bwd chain)
      auto issue_first = [&](int c) {
        const uint32_t gg = g0 + c, s = gg % kTcStages, use = gg / kTcStages;
        mbar_wait(bb + (TCB_WFULL + s) * 8, use & 1u);
        tc_fence_after();
        const uint32_t st = sb + kTcOffW + s * kTcStageBytes;
        const uint32_t slot = c & 1;
        if (leader) {
          const uint64_t b1 = make_smem_desc_k128(st);
          const uint64_t ax = bwd ? a_slab2 : a_slab0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tb + kTcColD1 + slot * 64, ax + 2 * kk, b1 + 2 * kk, idesc64, kk ? 1u : 0u);
          if (bwd) {
#pragma unroll
            for (int sl = 0; sl < s3; ++sl) {
              const uint64_t a3 = sl ? a_slab1 : a_slab0;
              const uint64_t b3 = make_smem_desc_k128(st + kTcHc * 128 + sl * kTcHc * 128);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tb + kTcColD3 + slot * 64, a3 + 2 * kk, b3 + 2 * kk, idesc64, (sl | kk) ? 1u : 0u);
            }
          }
          umma_commit(bb + (TCB_D1FULL + slot) * 8);
        }
        __syncwarp();
      };
      // second GEMM of chunk c: [D2a | D2b] += Hhi . [Bhi ; Blo]^T,  D2a += Hlo . Bhi^T   (a = D2a + D2b)
      auto issue_second = [&](int c) {
        const uint32_t gg = g0 + c, s = gg % kTcStages;
        const uint32_t st = sb + kTcOffW + s * kTcStageBytes;
        const uint32_t slot = c & 1;
        if (leader) {
          const uint32_t hb = sb + kTcOffH + slot * 32768u;
          const uint64_t hhi = make_smem_desc_k128(hb), hlo = make_smem_desc_k128(hb + 16384u);
          const uint64_t bhl = make_smem_desc_k128(st + (bwd ? (kTcHc * 128 + s3 * kTcHc * 128) : (kTcHc * 128)));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tb + kTcColD2, hhi + 2 * kk, bhl + 2 * kk, bwd ? idesc4 : idesc2, (c | kk) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tb + kTcColD2, hlo + 2 * kk, bhl + 2 * kk, bwd ? idesc4h : idesc2h, 1u);
          umma_commit(bb + (TCB_WEMPTY + s) * 8);
          if (c == kTcChunks - 1) umma_commit(bb + TCB_D2FULL * 8);
        }
        __syncwarp();
      };
      issue_first(0);
      issue_first(1);
#pragma unroll 1
      for (int c = 0; c < kTcChunks; ++c) {
        const uint32_t slot = c & 1;
        mbar_wait(bb + (TCB_HFULL + slot) * 8, (slot ? n_h1 : n_h0) & 1u);
        if (slot) n_h1++; else n_h0++;
        tc_fence_after();
        issue_second(c);
        if (c + 2 < kTcChunks) issue_first(c + 2);
      }
      g0 += kTcChunks;
    }
  } else {
    // ================= compute warps =================
    const int cw = warp - 2;                 // 0..7
    const int half = cw >> 2;
    const int lane_grp = warp & 3;           // TMEM lane group this warp may access
    const int row = lane_grp * 32 + lane;
    uint32_t n_d1a = 0, n_d1b = 0;
    if (half == 1) {
      const int nch = tc_num_chains<MODE>(nb, stash);
#pragma unroll 1
      for (int ci = 0; ci < nch; ++ci) {
        int kblk; bool bwd;
        tc_chain_info<MODE>(ci, nb, stash, kblk, bwd);
        if (bwd) tc_chain_convert<true>(sbase, bars, tmem_base, lane_grp, lane, 1, n_d1a, n_d1b);
        else tc_chain_convert<false>(sbase, bars, tmem_base, lane_grp, lane, 1, n_d1a, n_d1b);
      }
    } else {
      // ---- row threads
      const int grow = blockIdx.x * kTcRows + row;
      const bool ok = grow < A.M;
      float* ys = ys_all + row * 35;
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
      uint32_t n_d2 = 0;
      float x[C];
#pragma unroll
      for (int i = 0; i < C; ++i) x[i] = ok ? A.x[static_cast<size_t>(grow) * C + i] : 0.f;
      float ldp = 0.f;

      // dst[o] = sum_i tab[i][o] * ys[i]   (tab rows padded to kTcPermLd, 16-byte aligned)
      auto matvec = [&](const float* tab, float (&dst)[C]) {
#pragma unroll
        for (int o = 0; o < C; ++o) dst[o] = 0.f;
#pragma unroll 2
        for (int i = 0; i < C; ++i) {
          const float yi = ys[i];
          const float4* tr = reinterpret_cast<const float4*>(tab + i * kTcPermLd);
#pragma unroll
          for (int q = 0; q < (C + 3) / 4; ++q) {
            const float4 w = __ldg(tr + q);
            if (4 * q + 0 < C) dst[4 * q + 0] = fmaf(w.x, yi, dst[4 * q + 0]);
            if (4 * q + 1 < C) dst[4 * q + 1] = fmaf(w.y, yi, dst[4 * q + 1]);
            if (4 * q + 2 < C) dst[4 * q + 2] = fmaf(w.z, yi, dst[4 * q + 2]);
            if (4 * q + 3 < C) dst[4 * q + 3] = fmaf(w.w, yi, dst[4 * q + 3]);
          }
        }
      };
      // forward chain on v1 = first c1 entries of v: returns a = 0.1 * (subnet(v1)) in acc[0 .. 2c2)
      auto subnet_fwd = [&](const float (&v)[C], const TcTail& T, float (&acc)[2 * c2]) {
        float v1[c1];
#pragma unroll
        for (int i = 0; i < c1; ++i) v1[i] = v[i];
        tc_build_A<c1, 1, true>(sbase + kTcOffA, row, v1);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bars + TCB_AFULL * 8);
        tc_chain_convert<false>(sbase, bars, tmem_base, lane_grp, lane, 0, n_d1a, n_d1b);
        mbar_wait(bars + TCB_D2FULL * 8, n_d2 & 1u);
        n_d2++;
        tc_fence_after();
#pragma unroll
        for (int p = 0; p < n2p / 16; ++p) {
          uint32_t r[16], r2[16];
          tmem_ld16(lane_addr + kTcColD2 + p * 16, r);
          tmem_ld16(lane_addr + kTcColD2 + n2p + p * 16, r2);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (p * 16 + i < 2 * c2)
              acc[p * 16 + i] = 0.1f * ((__uint_as_float(r[i]) + __uint_as_float(r2[i])) + __ldg(T.b2 + p * 16 + i));
        }
        tc_fence_before();
      };

      constexpr int kStashRow = C + 2 * c2;                      // floats per (row, block): block input, then a
      static_assert(kStashRow % 4 == 0, "stash rows are moved as float4");
      float* stash_row = stash ? A.stash + static_cast<size_t>(ok ? grow : 0) * nb * kStashRow : nullptr;
      auto block_fwd = [&](int k) {
        const TcTail T = tc_tail<C>(A.packed, k);
        float a[2 * c2];
        subnet_fwd(x, T, a);
        if (MODE == FLOW_NLL_FWDBWD && stash && ok) {
          float buf[kStashRow];
#pragma unroll
          for (int i = 0; i < C; ++i) buf[i] = x[i];
#pragma unroll
          for (int i = 0; i < 2 * c2; ++i) buf[C + i] = a[i];
          float4* dst = reinterpret_cast<float4*>(stash_row + k * kStashRow);
#pragma unroll
          for (int q = 0; q < kStashRow / 4; ++q) dst[q] = make_float4(buf[4 * q], buf[4 * q + 1], buf[4 * q + 2], buf[4 * q + 3]);
        }
        ldp += T.logg;
#pragma unroll
        for (int c = 0; c < c2; ++c) {
          const float s = 2.f * tanhf(a[c]);
          x[c1 + c] = x[c1 + c] * expf(s) + a[c2 + c];
          ldp += s;
        }
#pragma unroll
        for (int i = 0; i < C; ++i) ys[i] = x[i] * __ldg(T.g + i) + __ldg(T.off + i);
        matvec(T.WT, x);
      };
      // inverse of block k applied to x (= block output): x <- block input; a (scaled subnet output) is returned
      auto block_rev = [&](int k, const TcTail& T, float (&a)[2 * c2]) {
#pragma unroll
        for (int i = 0; i < C; ++i) ys[i] = x[i];
        float y[C];
        matvec(T.WiT, y);
#pragma unroll
        for (int i = 0; i < C; ++i) y[i] = (y[i] - __ldg(T.off + i)) / __ldg(T.g + i);
        subnet_fwd(y, T, a);
#pragma unroll
        for (int i = 0; i < c1; ++i) x[i] = y[i];
#pragma unroll
        for (int c = 0; c < c2; ++c) {
          const float s = 2.f * tanhf(a[c]);
          x[c1 + c] = (y[c1 + c] - a[c2 + c]) * expf(-s);
          ldp -= s;
        }
        ldp -= T.logg;
        (void)k;
      };

      if (MODE == FLOW_FWD || MODE == FLOW_NLL_FWDBWD || MODE == FLOW_SAMPLE)
        for (int k = 0; k < nb; ++k) block_fwd(k);
      if (MODE == FLOW_SAMPLE) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float nz = ok ? A.noise[static_cast<size_t>(grow) * C + i] : 0.f;
          x[i] = x[i] + 0.2f * (nz * x[i]);
        }
      }
      if (MODE == FLOW_REV || MODE == FLOW_SAMPLE) {
        ldp = 0.f;
        for (int k = nb - 1; k >= 0; --k) {
          const TcTail T = tc_tail<C>(A.packed, k);
          float a[2 * c2];
          block_rev(k, T, a);
        }
      }
      if (MODE == FLOW_FWD || MODE == FLOW_REV) {
        if (ok) {
#pragma unroll
          for (int i = 0; i < C; ++i) A.out[static_cast<size_t>(grow) * C + i] = x[i];
          if (A.ld) A.ld[grow] = ldp;
        }
      }
      if (MODE == FLOW_SAMPLE) {
        if (ok) {
#pragma unroll
          for (int i = 0; i < C; ++i) {
            A.out[static_cast<size_t>(grow) * C + i] = A.x[static_cast<size_t>(grow) * C + i];
            A.out[static_cast<size_t>(A.M + grow) * C + i] = (i == 0 || i == C / 2) ? 0.f : x[i];
          }
        }
      }
      if (MODE == FLOW_NLL_FWDBWD) {
        float d[C];
        float zz = 0.f;
#pragma unroll
        for (int i = 0; i < C; ++i) {
          zz = fmaf(x[i], x[i], zz);
          d[i] = A.gz ? (ok ? A.gz[static_cast<size_t>(grow) * C + i] : 0.f) : A.scale * x[i];
        }
        const float glv = A.gz ? ((ok && A.gld) ? A.gld[grow] : 0.f) : -A.scale;
        float nll = ok ? 0.5f * zz - ldp : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nll += __shfl_xor_sync(0xffffffffu, nll, o);
        if (lane == 0 && A.nll_sum) atomicAdd(A.nll_sum, nll);
        for (int k = nb - 1; k >= 0; --k) {
          const TcTail T = tc_tail<C>(A.packed, k);
          // dyg = W^T d, dy = g * dyg   (d = gradient w.r.t. the block output)
#pragma unroll
          for (int i = 0; i < C; ++i) ys[i] = d[i];
          float dy[C];
          matvec(T.Wrow, dy);
          float dyg[C];
          const bool train = A.d_gscale != nullptr;
          if (train) {
#pragma unroll
            for (int i = 0; i < C; ++i) dyg[i] = ok ? dy[i] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < C; ++i) dy[i] *= __ldg(T.g + i);
          // the block input (x <- input) and the coupling coefficients: read back, or reconstructed with the inverse
          float a[2 * c2];
          if (stash) {
            const float4* src = reinterpret_cast<const float4*>(stash_row + k * kStashRow);
            float buf[kStashRow];
#pragma unroll
            for (int q = 0; q < kStashRow / 4; ++q) {
              const float4 t = ok ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
              buf[4 * q] = t.x; buf[4 * q + 1] = t.y; buf[4 * q + 2] = t.z; buf[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < C; ++i) x[i] = buf[i];
#pragma unroll
            for (int i = 0; i < 2 * c2; ++i) a[i] = buf[C + i];
          } else {
            block_rev(k, T, a);
          }
          float da[2 * c2];
#pragma unroll
          for (int c = 0; c < c2; ++c) {
            const float th = tanhf(a[c]);
            const float e = expf(2.f * th);
            const float dy2 = dy[c1 + c];
            d[c1 + c] = dy2 * e;
            const float ds = dy2 * x[c1 + c] * e + glv;
            da[c] = ds * 2.f * (1.f - th * th) * 0.1f;
            da[c2 + c] = dy2 * 0.1f;
          }
          if (train) {
            // global affine: out = (y * g + off) W^T, log-det += sum log g  ->  d off = sum_rows dyg,
            // d g = sum_rows (dyg * y + glv / g), d global_scale = d g * dgfac;  y = (x1, x2 * e^s + t)
            float* dgs = A.d_gscale + static_cast<size_t>(k) * C;
            float* dgo = A.d_goffset + static_cast<size_t>(k) * C;
            const float glv_ok = ok ? glv : 0.f;
#pragma unroll
            for (int i = 0; i < C; ++i) {
              float yv = x[i];
              if (i >= c1) {
                const float th = tanhf(a[i - c1]);
                yv = x[i] * expf(2.f * th) + a[c2 + (i - c1)];
              }
              float v_off = dyg[i];
              float v_g = (dyg[i] * yv + glv_ok / __ldg(T.g + i)) * __ldg(T.dgfac + i);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                v_off += __shfl_xor_sync(0xffffffffu, v_off, o);
                v_g += __shfl_xor_sync(0xffffffffu, v_g, o);
              }
              if (lane == 0) { atomicAdd(dgo + i, v_off); atomicAdd(dgs + i, v_g); }
            }
            if (ok) {
              __nv_bfloat16* ex1 = static_cast<__nv_bfloat16*>(A.ex_x1) + (static_cast<size_t>(k) * A.M + grow) * 64;
              __nv_bfloat16* exd = static_cast<__nv_bfloat16*>(A.ex_dsub) + (static_cast<size_t>(k) * A.M + grow) * 64;
#pragma unroll
              for (int i = 0; i < c1; ++i) ex1[i] = __float2bfloat16_rn(x[i]);
#pragma unroll
              for (int i = 0; i < 2 * c2; ++i) exd[i] = __float2bfloat16_rn(da[i]);
            }
          }
          // backward chain: dx1 = ((da . W2) * relu'(W1 x1 + b1)) . W1
          {
            float v1[c1];
#pragma unroll
            for (int i = 0; i < c1; ++i) v1[i] = x[i];
            tc_build_A<c1, 1, true>(sbase + kTcOffA + 2 * 16384u, row, v1);   // x1 operand (mask recompute), slab 2
            tc_build_A<2 * c2, s3, false>(sbase + kTcOffA, row, da);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bars + TCB_AFULL * 8);
            tc_chain_convert<true>(sbase, bars, tmem_base, lane_grp, lane, 0, n_d1a, n_d1b);
            mbar_wait(bars + TCB_D2FULL * 8, n_d2 & 1u);
            n_d2++;
            tc_fence_after();
#pragma unroll
            for (int p = 0; p < n4p / 16; ++p) {
              uint32_t r[16], r2[16];
              tmem_ld16(lane_addr + kTcColD2 + p * 16, r);
              tmem_ld16(lane_addr + kTcColD2 + n4p + p * 16, r2);
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (p * 16 + i < c1) d[p * 16 + i] = dy[p * 16 + i] + (__uint_as_float(r[i]) + __uint_as_float(r2[i]));
            }
            tc_fence_before();
          }
        }
        if (ok && A.out) {
#pragma unroll
          for (int i = 0; i < C; ++i) A.out[static_cast<size_t>(grow) * C + i] = d[i];
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTcTmemCols);
  }
}

}  // namespace links

// Normalising flow: FrEIA SequenceINN of n_blocks AllInOneBlock(subnet_fc, permute_soft=True) on C-dim rows.
// Semantics restated in oracle/flow.py (FrEIA is third-party; reference call sites
// train_leg_torso_lifter.py:134-136,207-214, train_left_right_lifter.py:131-133,334-340,
// train_full_pose_norm_flow.py:75-90).
//
// One CTA (8 warps) carries 32 rows through ALL coupling blocks with the state in shared memory, so the
// [M,1024] hidden activations never exist in HBM.  lane = row, warp = 128-wide slice of the hidden layer:
// per hidden unit h a thread does c1 FMAs (layer 0), ReLU, 2*c2 FMAs (layer 2) against warp-uniform weight
// rows (one packed 16B-aligned record per h).  The soft clamp 2*tanh(0.1 a), exp scale / shift, ActNorm-style
// global affine, C x C soft permutation and the per-row log-det are fused behind it.  fp32 throughout: the
// flow is < 1 % of the step's FLOPs and its log-det/exp terms are the precision-critical part of the loss.
#pragma once
#include "devdefs.cuh"

namespace links {

constexpr int kFlowHidden = 1024;
constexpr int kFlowRows = 32;
constexpr int kFlowWarps = 8;
constexpr int kFlowMaxBlocks = 8;
constexpr int kFlowSlice = kFlowHidden / kFlowWarps;   // 128 hidden units per warp

enum FlowMode { FLOW_FWD = 0, FLOW_REV = 1, FLOW_NLL_FWDBWD = 2, FLOW_SAMPLE = 3 };

__host__ __device__ constexpr int flow_c1(int C) { return C - C / 2; }
__host__ __device__ constexpr int flow_c2(int C) { return C / 2; }
// floats per hidden-unit record: [b1, W1 row (c1), W2 column (2*c2)] padded to a float4 multiple
__host__ __device__ constexpr int flow_sh(int C) { return (1 + flow_c1(C) + 2 * flow_c2(C) + 3) & ~3; }
// packed block layout (floats): [hid records 1024*Sh][b2 2c2][g C][off C][wperm C*C][wperm_inv C*C][logg 1] -> pad 4
__host__ __device__ constexpr int flow_block_floats(int C) {
  return (kFlowHidden * flow_sh(C) + 2 * flow_c2(C) + 2 * C + 2 * C * C + 1 + 3) & ~3;
}

struct FlowPackArgs {
  const float* w0[kFlowMaxBlocks];
  const float* b0[kFlowMaxBlocks];
  const float* w2[kFlowMaxBlocks];
  const float* b2[kFlowMaxBlocks];
  const float* gs[kFlowMaxBlocks];
  const float* go[kFlowMaxBlocks];
  const float* wp[kFlowMaxBlocks];
  const float* wpi[kFlowMaxBlocks];
  float* packed;
  int C, n_blocks;
};

// grid = (n_blocks, kFlowPackSplit), block = 256: the flow training step re-packs after every Adam update, so the record
// loop of a coupling block is spread over kFlowPackSplit CTAs (one CTA per block took ~80 us); slice 0 also writes the tail
constexpr int kFlowPackSplit = 16;
__global__ void flow_pack_kernel(const FlowPackArgs A) {
  const int k = blockIdx.x;
  const int C = A.C, c1 = flow_c1(C), c2 = flow_c2(C), Sh = flow_sh(C);
  float* P = A.packed + static_cast<size_t>(k) * flow_block_floats(C);
  const int e_step = blockDim.x * gridDim.y;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < kFlowHidden * Sh; e += e_step) {
    const int h = e / Sh, i = e - h * Sh;
    float v = 0.f;
    if (i == 0) v = A.b0[k][h];
    else if (i <= c1) v = A.w0[k][h * c1 + (i - 1)];
    else if (i <= c1 + 2 * c2) v = A.w2[k][static_cast<size_t>(i - 1 - c1) * kFlowHidden + h];
    P[e] = v;
  }
  if (blockIdx.y != 0) return;
  float* T = P + kFlowHidden * Sh;
  for (int i = threadIdx.x; i < 2 * c2; i += blockDim.x) T[i] = A.b2[k][i];
  float* G = T + 2 * c2;
  float* O = G + C;
  float* W = O + C;
  float* WI = W + C * C;
  __shared__ float s_logg[256];
  float lg = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float x = A.gs[k][i];
    // 0.1 * softplus(beta = 0.5, threshold = 20)
    const float sp = (0.5f * x > 20.f) ? x : 2.f * log1pf(expf(0.5f * x));
    const float g = 0.1f * sp;
    G[i] = g;
    O[i] = A.go[k][i];
    lg += logf(g);
  }
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) { W[i] = A.wp[k][i]; WI[i] = A.wpi[k][i]; }
  s_logg[threadIdx.x] = lg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x); ++i) t += s_logg[i];
    WI[C * C] = t;
  }
}

struct FlowArgs {
  const float* packed;
  const float* x;       // [M, C]
  const float* noise;   // SAMPLE: [M, C]
  float* out;           // FWD/REV: [M, C];  NLL_FWDBWD: dx [M, C];  SAMPLE: [2M, C]
  float* ld;            // FWD/REV: [M]
  float* nll_sum;       // NLL_FWDBWD
  float scale;          // NLL_FWDBWD: dx = scale * d nll / dx
  const float* gz;      // NLL_FWDBWD, optional: general VJP seed d/dz [M, C] (replaces scale * z)
  const float* gld;     //                       and d/d(log_jac_det) [M] (replaces -scale)
  void* ex_x1;          // flow training exports (tensor-core kernel only, see flow_tc.cuh)
  void* ex_dsub;
  float* d_gscale;
  float* d_goffset;
  float* stash;         // NLL_FWDBWD, tensor-core kernel only, optional: activation stash (see flow_tc.cuh)
  int M, n_blocks;
};

template <int C>
struct FlowSmem {
  static constexpr int c1 = flow_c1(C), c2 = flow_c2(C);
  static constexpr int XS = C + 1;          // padded row strides (odd -> conflict free for lane = row)
  static constexpr int AS = 2 * c2 + 1;
  static constexpr int DS = c1 + 1;
  static constexpr int n_state = kFlowMaxBlocks + 1;
  static constexpr int off_xs = 0;                                          // [n_state][32][XS]
  static constexpr int off_red = off_xs + n_state * kFlowRows * XS;         // [8][32][AS]  (aliased by dx1 partials)
  static constexpr int off_a = off_red + kFlowWarps * kFlowRows * AS;       // [32][AS]
  static constexpr int off_y = off_a + kFlowRows * AS;                      // [32][XS]
  static constexpr int off_d0 = off_y + kFlowRows * XS;                     // [32][XS]
  static constexpr int off_d1 = off_d0 + kFlowRows * XS;                    // [32][XS]
  static constexpr int off_draw = off_d1 + kFlowRows * XS;                  // [32][AS]
  static constexpr int off_ld = off_draw + kFlowRows * AS;                  // [8][32]
  static constexpr int total = off_ld + kFlowWarps * kFlowRows;
  static constexpr size_t bytes = static_cast<size_t>(total) * sizeof(float);
};

template <int C>
struct FlowBlockPtrs {
  const float* hid; const float* b2; const float* g; const float* off; const float* w; const float* wi; float logg;
  __device__ __forceinline__ FlowBlockPtrs(const float* packed, int k) {
    const float* P = packed + static_cast<size_t>(k) * flow_block_floats(C);
    hid = P;
    b2 = P + kFlowHidden * flow_sh(C);
    g = b2 + 2 * flow_c2(C);
    off = g + C;
    w = off + C;
    wi = w + C * C;
    logg = wi[C * C];
  }
};

// Subnet partial pass for this warp's hidden slice: acc[o] = sum_h relu(b1 + W1 x1)[h] * W2[o][h].
// relu_bits (4 words) records hid > 0 for the slice.
template <int C>
__device__ __forceinline__ void subnet_partial(const float* __restrict__ hid, int warp, const float (&x1)[flow_c1(C)],
                                               float (&acc)[2 * flow_c2(C) > 0 ? 2 * flow_c2(C) : 1], uint32_t (&relu_bits)[4]) {
  constexpr int c1 = flow_c1(C), c2 = flow_c2(C), Sh = flow_sh(C);
#pragma unroll
  for (int o = 0; o < 2 * c2; ++o) acc[o] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) relu_bits[q] = 0u;
  const int h0 = warp * kFlowSlice;
#pragma unroll 2
  for (int hh = 0; hh < kFlowSlice; ++hh) {
    const float4* rec = reinterpret_cast<const float4*>(hid + static_cast<size_t>(h0 + hh) * Sh);
    float w[Sh];
#pragma unroll
    for (int q = 0; q < Sh / 4; ++q) {
      const float4 t = __ldg(rec + q);
      w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
    }
    float hv = w[0];
#pragma unroll
    for (int i = 0; i < c1; ++i) hv = fmaf(x1[i], w[1 + i], hv);
    if (hv > 0.f) relu_bits[hh >> 5] |= 1u << (hh & 31);
    hv = fmaxf(hv, 0.f);
#pragma unroll
    for (int o = 0; o < 2 * c2; ++o) acc[o] = fmaf(hv, w[1 + c1 + o], acc[o]);
  }
}

// a_s[r][o] = 0.1 * (b2[o] + sum_w red[w][r][o])      (FrEIA: a *= 0.1)
template <int C>
__device__ __forceinline__ void subnet_reduce(float* sm, const FlowBlockPtrs<C>& B, int warp, int lane) {
  using S = FlowSmem<C>;
  for (int o = warp; o < 2 * S::c2; o += kFlowWarps) {
    float a = __ldg(B.b2 + o);
#pragma unroll
    for (int w = 0; w < kFlowWarps; ++w) a += sm[S::off_red + (w * kFlowRows + lane) * S::AS + o];
    sm[S::off_a + lane * S::AS + o] = 0.1f * a;
  }
}

template <int C>
__device__ __forceinline__ void run_subnet(float* sm, const FlowBlockPtrs<C>& B, const float* xin /*[32][XS]*/, int warp,
                                           int lane, uint32_t (&relu_bits)[4]) {
  using S = FlowSmem<C>;
  float x1[S::c1];
#pragma unroll
  for (int i = 0; i < S::c1; ++i) x1[i] = xin[lane * S::XS + i];
  float acc[2 * S::c2 > 0 ? 2 * S::c2 : 1];
  subnet_partial<C>(B.hid, warp, x1, acc, relu_bits);
#pragma unroll
  for (int o = 0; o < 2 * S::c2; ++o) sm[S::off_red + (warp * kFlowRows + lane) * S::AS + o] = acc[o];
  __syncthreads();
  subnet_reduce<C>(sm, B, warp, lane);
  __syncthreads();
}

// forward through block k: xin -> xout; returns this thread's share of the row log-det
template <int C>
__device__ __forceinline__ float block_forward(float* sm, const float* packed, int k, const float* xin, float* xout,
                                               int warp, int lane) {
  using S = FlowSmem<C>;
  const FlowBlockPtrs<C> B(packed, k);
  uint32_t bits[4];
  run_subnet<C>(sm, B, xin, warp, lane, bits);
  float ldp = (warp == 0) ? B.logg : 0.f;
  float* yg = sm + S::off_y;
  for (int i = warp; i < C; i += kFlowWarps) {
    float y = xin[lane * S::XS + i];
    if (i >= S::c1) {
      const int c = i - S::c1;
      const float s = 2.f * tanhf(sm[S::off_a + lane * S::AS + c]);
      const float t = sm[S::off_a + lane * S::AS + S::c2 + c];
      y = y * expf(s) + t;
      ldp += s;
    }
    yg[lane * S::XS + i] = y * __ldg(B.g + i) + __ldg(B.off + i);
  }
  __syncthreads();
  for (int o = warp; o < C; o += kFlowWarps) {
    float acc = 0.f;
#pragma unroll 2
    for (int i = 0; i < C; ++i) acc = fmaf(__ldg(B.w + o * C + i), yg[lane * S::XS + i], acc);
    xout[lane * S::XS + o] = acc;
  }
  __syncthreads();
  return ldp;
}

// reverse through block k
template <int C>
__device__ __forceinline__ float block_reverse(float* sm, const float* packed, int k, const float* xin, float* xout,
                                               int warp, int lane) {
  using S = FlowSmem<C>;
  const FlowBlockPtrs<C> B(packed, k);
  float* ys = sm + S::off_y;
  for (int i = warp; i < C; i += kFlowWarps) {
    float acc = 0.f;
#pragma unroll 2
    for (int j = 0; j < C; ++j) acc = fmaf(__ldg(B.wi + i * C + j), xin[lane * S::XS + j], acc);
    ys[lane * S::XS + i] = (acc - __ldg(B.off + i)) / __ldg(B.g + i);
  }
  __syncthreads();
  uint32_t bits[4];
  run_subnet<C>(sm, B, ys, warp, lane, bits);
  float ldp = (warp == 0) ? -B.logg : 0.f;
  for (int i = warp; i < C; i += kFlowWarps) {
    float y = ys[lane * S::XS + i];
    if (i >= S::c1) {
      const int c = i - S::c1;
      const float s = 2.f * tanhf(sm[S::off_a + lane * S::AS + c]);
      const float t = sm[S::off_a + lane * S::AS + S::c2 + c];
      y = (y - t) * expf(-s);
      ldp -= s;
    }
    xout[lane * S::XS + i] = y;
  }
  __syncthreads();
  return ldp;
}

// backward through (forward) block k: din = d/d(block output), glv = dL/d(log-det) -> dout = d/d(block input)
template <int C>
__device__ __forceinline__ void block_backward(float* sm, const float* packed, int k, const float* xin, const float* din,
                                               float* dout, float glv, int warp, int lane) {
  using S = FlowSmem<C>;
  const FlowBlockPtrs<C> B(packed, k);
  float* dy = sm + S::off_y;
  // B0: dy = g * (W^T din)
  for (int i = warp; i < C; i += kFlowWarps) {
    float acc = 0.f;
#pragma unroll 2
    for (int o = 0; o < C; ++o) acc = fmaf(__ldg(B.w + o * C + i), din[lane * S::XS + o], acc);
    dy[lane * S::XS + i] = acc * __ldg(B.g + i);
  }
  // B1: recompute the subnet (a, relu mask) -- includes the barriers that also publish dy
  uint32_t bits[4];
  run_subnet<C>(sm, B, xin, warp, lane, bits);
  // B2: coupling
  float* draw = sm + S::off_draw;
  for (int c = warp; c < S::c2; c += kFlowWarps) {
    const float th = tanhf(sm[S::off_a + lane * S::AS + c]);
    const float s = 2.f * th;
    const float e = expf(s);
    const float x2 = xin[lane * S::XS + S::c1 + c];
    const float dy2 = dy[lane * S::XS + S::c1 + c];
    dout[lane * S::XS + S::c1 + c] = dy2 * e;
    const float ds = dy2 * x2 * e + glv;
    draw[lane * S::AS + c] = ds * 2.f * (1.f - th * th) * 0.1f;
    draw[lane * S::AS + S::c2 + c] = dy2 * 0.1f;
  }
  __syncthreads();
  // B3: dx1 partials over this warp's hidden slice
  {
    constexpr int c1 = S::c1, c2 = S::c2, Sh = flow_sh(C);
    float dr[2 * c2 > 0 ? 2 * c2 : 1];
#pragma unroll
    for (int o = 0; o < 2 * c2; ++o) dr[o] = draw[lane * S::AS + o];
    float dx1[c1];
#pragma unroll
    for (int i = 0; i < c1; ++i) dx1[i] = 0.f;
    const int h0 = warp * kFlowSlice;
#pragma unroll 2
    for (int hh = 0; hh < kFlowSlice; ++hh) {
      const float4* rec = reinterpret_cast<const float4*>(B.hid + static_cast<size_t>(h0 + hh) * Sh);
      float w[Sh];
#pragma unroll
      for (int q = 0; q < Sh / 4; ++q) {
        const float4 t = __ldg(rec + q);
        w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
      }
      float dh = 0.f;
#pragma unroll
      for (int o = 0; o < 2 * c2; ++o) dh = fmaf(w[1 + c1 + o], dr[o], dh);
      dh = ((bits[hh >> 5] >> (hh & 31)) & 1u) ? dh : 0.f;
#pragma unroll
      for (int i = 0; i < c1; ++i) dx1[i] = fmaf(w[1 + i], dh, dx1[i]);
    }
    float* part = sm + S::off_red;   // alias: [8][32][DS]
#pragma unroll
    for (int i = 0; i < c1; ++i) part[(warp * kFlowRows + lane) * S::DS + i] = dx1[i];
  }
  __syncthreads();
  for (int i = warp; i < S::c1; i += kFlowWarps) {
    float acc = dy[lane * S::XS + i];
#pragma unroll
    for (int w = 0; w < kFlowWarps; ++w) acc += sm[S::off_red + (w * kFlowRows + lane) * S::DS + i];
    dout[lane * S::XS + i] = acc;
  }
  __syncthreads();
}

template <int C, int MODE>
__global__ void __launch_bounds__(kFlowWarps * 32) flow_kernel(const FlowArgs A) {
  using S = FlowSmem<C>;
  LINKS_DYN_SMEM(float, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * kFlowRows + lane;
  const bool ok = row < A.M;
  float* xs = sm + S::off_xs;
  auto state = [&](int k) { return xs + k * kFlowRows * S::XS; };
  for (int i = warp; i < C; i += kFlowWarps) state(0)[lane * S::XS + i] = ok ? A.x[static_cast<size_t>(row) * C + i] : 0.f;
  __syncthreads();
  float ldp = 0.f;
  const int nb = A.n_blocks;
  float* result = nullptr;
  if (MODE == FLOW_FWD || MODE == FLOW_NLL_FWDBWD || MODE == FLOW_SAMPLE) {
    for (int k = 0; k < nb; ++k) {
      // FWD / SAMPLE only need two ping-pong states; NLL_FWDBWD keeps every block input for the backward
      float* xin = (MODE == FLOW_NLL_FWDBWD) ? state(k) : state(k & 1);
      float* xout = (MODE == FLOW_NLL_FWDBWD) ? state(k + 1) : state((k + 1) & 1);
      ldp += block_forward<C>(sm, A.packed, k, xin, xout, warp, lane);
      result = xout;
    }
  }
  if (MODE == FLOW_SAMPLE) {
    // z' = z + 0.2 * noise * z  (utils/helpers.py:298-308), then the reverse pass
    float* z = result;
    for (int i = warp; i < C; i += kFlowWarps) {
      const float zz = z[lane * S::XS + i];
      const float nz = ok ? A.noise[static_cast<size_t>(row) * C + i] : 0.f;
      z[lane * S::XS + i] = zz + 0.2f * (nz * zz);
    }
    __syncthreads();
  }
  if (MODE == FLOW_REV || MODE == FLOW_SAMPLE) {
    float* cur = (MODE == FLOW_SAMPLE) ? result : state(0);
    float* other = (cur == state(0)) ? state(1) : state(0);
    ldp = 0.f;
    for (int k = nb - 1; k >= 0; --k) {
      ldp += block_reverse<C>(sm, A.packed, k, cur, other, warp, lane);
      float* t = cur; cur = other; other = t;
    }
    result = cur;
  }
  if (MODE == FLOW_FWD || MODE == FLOW_REV) {
    float* lds = sm + S::off_ld;
    lds[warp * kFlowRows + lane] = ldp;
    __syncthreads();
    if (ok) {
      for (int i = warp; i < C; i += kFlowWarps) A.out[static_cast<size_t>(row) * C + i] = result[lane * S::XS + i];
      if (warp == 0 && A.ld) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kFlowWarps; ++w) t += lds[w * kFlowRows + lane];
        A.ld[row] = t;
      }
    }
  }
  if (MODE == FLOW_SAMPLE) {
    // out = [x ; s] with the root joint of s zeroed (train_leg_torso_lifter.py:137-142); C == 34
    if (ok) {
      for (int i = warp; i < C; i += kFlowWarps) {
        A.out[static_cast<size_t>(row) * C + i] = A.x[static_cast<size_t>(row) * C + i];
        const float v = (i == 0 || i == C / 2) ? 0.f : result[lane * S::XS + i];
        A.out[static_cast<size_t>(A.M + row) * C + i] = v;
      }
    }
  }
  if (MODE == FLOW_NLL_FWDBWD) {
    float* lds = sm + S::off_ld;
    lds[warp * kFlowRows + lane] = ldp;
    float* d0 = sm + S::off_d0;
    float* d1 = sm + S::off_d1;
    const float* z = state(nb);
    for (int i = warp; i < C; i += kFlowWarps)
      d0[lane * S::XS + i] = A.gz ? (ok ? A.gz[static_cast<size_t>(row) * C + i] : 0.f) : A.scale * z[lane * S::XS + i];
    const float glv = A.gz ? ((ok && A.gld) ? A.gld[row] : 0.f) : -A.scale;
    __syncthreads();
    if (warp == 0) {
      float ldt = 0.f;
#pragma unroll
      for (int w = 0; w < kFlowWarps; ++w) ldt += lds[w * kFlowRows + lane];
      float zz = 0.f;
      for (int i = 0; i < C; ++i) zz += z[lane * S::XS + i] * z[lane * S::XS + i];
      float nll = ok ? 0.5f * zz - ldt : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nll += __shfl_xor_sync(LINKS_FULL_MASK, nll, o);
      if (lane == 0 && A.nll_sum) atomicAdd(A.nll_sum, nll);
    }
    float* din = d0;
    float* dout = d1;
    for (int k = nb - 1; k >= 0; --k) {
      block_backward<C>(sm, A.packed, k, state(k), din, dout, glv, warp, lane);
      float* t = din; din = dout; dout = t;
    }
    if (ok && A.out) {
      for (int i = warp; i < C; i += kFlowWarps) A.out[static_cast<size_t>(row) * C + i] = din[lane * S::XS + i];
    }
  }
}

}  // namespace links

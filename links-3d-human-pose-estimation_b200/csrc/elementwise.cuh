// Operand packing, bias-gradient column sums, weight shadow casts and the fused Adam step.
// Device code only (also compiled by tests/hostsim with a CUDA shim).
#pragma once
#include "devdefs.cuh"

namespace links {

// dst[m, c] = bf16(src[m*ld_src + idx[c]]) (zero padded to 64 columns), dstT[c, colT0 + m] likewise.
// One thread per (row, 64-col slot); reference index maps: utils/helpers.py:55-65,
// train_leg_torso_lifter.py:147-148, train_occlusion_models.py:185-191.
// `period` > 1 supports gathers that mix `period` consecutive rows (split_data_left_right_3d,
// utils/helpers.py:81-91, period 2): idx then holds period*n_idx offsets relative to the row group.
__global__ void pack_rows_kernel(const float* __restrict__ src, int ld_src, int M, const int* __restrict__ idx,
                                 int n_idx, int period, __nv_bfloat16* __restrict__ dst,
                                 __nv_bfloat16* __restrict__ dstT, int ldT, int colT0) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = t >> 6;
  const int c = t & 63;
  if (m >= M) return;
  float v = 0.f;
  if (c < n_idx) {
    const int grp = m / period, sub = m - grp * period;
    v = src[static_cast<size_t>(grp) * period * ld_src + idx[sub * n_idx + c]];
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  dst[static_cast<size_t>(m) * 64 + c] = h;
  if (dstT != nullptr && c < n_idx) dstT[static_cast<size_t>(c) * ldT + colT0 + m] = h;
}

// out[n] (+)= sum_m G[m, n].  Block = 32 x 8 threads: 32 columns, 8 row phases; grid.x over column
// groups, grid.y over row slabs (atomics combine slabs).
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ G, int ldg, int M, int N,
                                   float* __restrict__ out, int rows_per_block) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m0 = blockIdx.y * rows_per_block;
  int m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float acc = 0.f;
  if (n < N) {
    for (int m = m0 + threadIdx.y; m < m1; m += 8) acc += __bfloat162float(G[static_cast<size_t>(m) * ldg + n]);
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}

// Batched variant: one launch for all bias gradients of a step.  grid = (col groups, items, row slabs).
struct ColsumBatch {
  LinksColsumItem it[LINKS_MAX_COLSUM_ITEMS];
  int n;
};
__global__ void colsum_batched_zero_kernel(const ColsumBatch B) {
  const LinksColsumItem& I = B.it[blockIdx.x];
  if (I.accumulate) return;
  for (int n = threadIdx.x; n < I.N; n += blockDim.x) I.out[n] = 0.f;
}
__global__ void colsum_batched_kernel(const ColsumBatch B, int rows_per_block) {
  __shared__ float red[8][33];
  const LinksColsumItem& I = B.it[blockIdx.y];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m0 = blockIdx.z * rows_per_block;
  if (blockIdx.x * 32 >= I.N || m0 >= I.M) return;     // block-uniform
  int m1 = m0 + rows_per_block;
  if (m1 > I.M) m1 = I.M;
  const __nv_bfloat16* G = static_cast<const __nv_bfloat16*>(I.G);
  float acc = 0.f;
  if (n < I.N) {
    for (int m = m0 + threadIdx.y; m < m1; m += 8) acc += __bfloat162float(G[static_cast<size_t>(m) * I.ldg + n]);
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < I.N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(I.out + n, s);
  }
}

// fp32 W[N,K] -> bf16 Wb[N, ldw] (cols >= K zero) and bf16 WT[K, ldwt] (cols >= N zero), 32x32 smem tiles.
__global__ void cast_weight_kernel(const float* __restrict__ W, int N, int K, __nv_bfloat16* __restrict__ Wb, int ldw,
                                   __nv_bfloat16* __restrict__ WT, int ldwt) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    float v = 0.f;
    if (n < N && k < K) v = W[static_cast<size_t>(n) * K + k];
    tile[r][threadIdx.x] = v;
    if (Wb != nullptr && n < N && k < ldw) Wb[static_cast<size_t>(n) * ldw + k] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (WT != nullptr) {
    for (int r = threadIdx.y; r < 32; r += 8) {
      const int k = k0 + r, n = n0 + threadIdx.x;
      if (k < K && n < ldwt) WT[static_cast<size_t>(k) * ldwt + n] = __float2bfloat16_rn(tile[threadIdx.x][r]);
    }
  }
}

// Batched shadow refresh: all layers of a network set in one launch.  grid = (row-slabs, items); a warp converts rows
// of fp32 W[N,K] to bf16 Wb[N, ldw] (columns >= K zero) with 8-element vectors when the layout allows.
struct alignas(16) CastBf8 { __nv_bfloat16 v[8]; };
struct CastBatch {
  LinksCastItem it[LINKS_MAX_CAST_ITEMS];
  int n;
};
__global__ void cast_weight_batched_kernel(const CastBatch B) {
  const LinksCastItem& I = B.it[blockIdx.y];
  const int N = I.N, K = I.K, ldw = I.ldw;
  const float* __restrict__ W = I.W;
  __nv_bfloat16* __restrict__ Wb = static_cast<__nv_bfloat16*>(I.Wb);
  const bool vec = (K % 8 == 0) && (ldw % 8 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Wb) & 15u) == 0);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  if (vec) {
    const int kv = ldw / 8;
    const size_t total = static_cast<size_t>(N) * kv;
    for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
      const int n = static_cast<int>(e / kv), k = static_cast<int>(e - static_cast<size_t>(n) * kv) * 8;
      CastBf8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = __float2bfloat16_rn(0.f);
      if (k < K) {
        const float4 a = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * K + k);
        const float4 b = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * K + k + 4);
        o.v[0] = __float2bfloat16_rn(a.x); o.v[1] = __float2bfloat16_rn(a.y);
        o.v[2] = __float2bfloat16_rn(a.z); o.v[3] = __float2bfloat16_rn(a.w);
        o.v[4] = __float2bfloat16_rn(b.x); o.v[5] = __float2bfloat16_rn(b.y);
        o.v[6] = __float2bfloat16_rn(b.z); o.v[7] = __float2bfloat16_rn(b.w);
      }
      *reinterpret_cast<CastBf8*>(Wb + static_cast<size_t>(n) * ldw + k) = o;
    }
  } else {
    const size_t total = static_cast<size_t>(N) * ldw;
    for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
      const int n = static_cast<int>(e / ldw), k = static_cast<int>(e - static_cast<size_t>(n) * ldw);
      Wb[e] = __float2bfloat16_rn(k < K ? W[static_cast<size_t>(n) * K + k] : 0.f);
    }
  }
}

// torch.optim.Adam (coupled L2 weight decay), same operation order as torch's single-tensor path:
//   g += wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g;
//   denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= (lr/(1-b1^t)) * m/denom
__device__ __forceinline__ float links_grad_load(const float* g, size_t i) { return g[i]; }
__device__ __forceinline__ float links_grad_load(const __nv_bfloat16* g, size_t i) { return __bfloat162float(g[i]); }

// fp32 gradients -> bf16 (gradient compression for the data-parallel all-reduce: half the NVLink bytes)
__global__ void grad_compress_bf16_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ out, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16_rn(g[i]);
}

template <typename GradT>
__global__ void adam_kernel(float* __restrict__ p, const GradT* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                            const int* __restrict__ step_dev, int step_host, float grad_scale,
                            const float* __restrict__ lr_dev) {
  if (lr_dev != nullptr) lr = *lr_dev;      // device-side learning rate: graph replays follow the scheduler
  // step number t: host value, or (completed steps on device) + 1 so that a captured CUDA graph replays correctly
  const int t = step_dev != nullptr ? (*step_dev + 1) : step_host;
  const float bc1 = 1.f - powf(b1, static_cast<float>(t));
  const float bc2_sqrt = sqrtf(1.f - powf(b2, static_cast<float>(t)));
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float pi = p[i];
    float gi = links_grad_load(g, i) * grad_scale;
    gi = gi + wd * pi;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // torch: exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;         // torch: mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float step = lr / bc1;
    p[i] = pi - step * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// out[i] = sum_j mat[i * n_in + j] * in[j]  (tiny: the step's loss summary -- normalisation, weighting and totals of the
// device-side loss sums in one launch instead of a handful of framework element-wise kernels)
__global__ void small_matvec_kernel(const float* __restrict__ mat, const float* __restrict__ in, int n_in, int n_out,
                                    float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  float acc = 0.f;
  for (int j = 0; j < n_in; ++j) acc = fmaf(mat[i * n_in + j], in[j], acc);
  out[i] = acc;
}

__global__ void adam_prepare_kernel(const int* __restrict__ step_dev, const float* __restrict__ lr_dev, float lr, float b1,
                                    float b2, float eps, float wd, float grad_scale, float* __restrict__ hyper) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int t = *step_dev + 1;
  if (lr_dev != nullptr) lr = *lr_dev;
  const float bc1 = 1.f - powf(b1, static_cast<float>(t));
  const float bc2_sqrt = sqrtf(1.f - powf(b2, static_cast<float>(t)));
  hyper[0] = lr / bc1; hyper[1] = bc2_sqrt; hyper[2] = eps; hyper[3] = b1;
  hyper[4] = b2; hyper[5] = wd; hyper[6] = grad_scale; hyper[7] = static_cast<float>(t);
}

// Sharded Adam behind the push reduce-scatter (include/links_b200.h::links_adam_zero).  grid = (blocks, n_layers); a thread
// owns 8 consecutive elements of the layer's owned block (16 bytes of bf16 per staging slot / shadow).
__global__ void adam_zero_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                 const __nv_bfloat16* __restrict__ stage, size_t slot_elems,
                                 const LinksAdamZeroLayer* __restrict__ layers, int rows_per_owner, int cols, int world, int rank,
                                 const float* __restrict__ hyper) {
  const LinksAdamZeroLayer L = layers[blockIdx.y];
  const size_t n = static_cast<size_t>(rows_per_owner) * cols;         // owned elements of this layer
  const float step_size = hyper[0], bc2_sqrt = hyper[1], eps = hyper[2], b1 = hyper[3], b2 = hyper[4], wd = hyper[5],
              gs = hyper[6];
  const size_t own0 = static_cast<size_t>(rank) * n;                   // first owned element inside the layer
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 8;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < world; ++s) {
      const uint4 q = *reinterpret_cast<const uint4*>(stage + static_cast<size_t>(s) * slot_elems + L.stage_off + i);
      const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[2 * e] += __uint_as_float(w[e] << 16);
        g[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
      }
    }
    const size_t mo = L.master_off + own0 + i;
    float pa[8], ma[8], va[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 P4 = *reinterpret_cast<const float4*>(p + mo + 4 * h), M4 = *reinterpret_cast<const float4*>(m + mo + 4 * h),
                   V4 = *reinterpret_cast<const float4*>(v + mo + 4 * h);
      pa[4 * h] = P4.x; pa[4 * h + 1] = P4.y; pa[4 * h + 2] = P4.z; pa[4 * h + 3] = P4.w;
      ma[4 * h] = M4.x; ma[4 * h + 1] = M4.y; ma[4 * h + 2] = M4.z; ma[4 * h + 3] = M4.w;
      va[4 * h] = V4.x; va[4 * h + 1] = V4.y; va[4 * h + 2] = V4.z; va[4 * h + 3] = V4.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float gi = g[e] * gs;
      gi = gi + wd * pa[e];
      const float mi = ma[e] + (gi - ma[e]) * (1.f - b1);
      const float vi = va[e] * b2 + (1.f - b2) * gi * gi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pa[e] = pa[e] - step_size * (mi / denom);
      ma[e] = mi;
      va[e] = vi;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(p + mo + 4 * h) = make_float4(pa[4 * h], pa[4 * h + 1], pa[4 * h + 2], pa[4 * h + 3]);
      *reinterpret_cast<float4*>(m + mo + 4 * h) = make_float4(ma[4 * h], ma[4 * h + 1], ma[4 * h + 2], ma[4 * h + 3]);
      *reinterpret_cast<float4*>(v + mo + 4 * h) = make_float4(va[4 * h], va[4 * h + 1], va[4 * h + 2], va[4 * h + 3]);
    }
    uint4 sh;
    {
      unsigned int hw[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const __nv_bfloat16 b = __float2bfloat16_rn(pa[e]);
        hw[e] = *reinterpret_cast<const unsigned short*>(&b);
      }
      sh.x = hw[0] | (hw[1] << 16); sh.y = hw[2] | (hw[3] << 16); sh.z = hw[4] | (hw[5] << 16); sh.w = hw[6] | (hw[7] << 16);
    }
    for (int r = 0; r < world; ++r)
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(L.shadow[r]) + own0 + i) = sh;
  }
}

__global__ void adam_incr_kernel(int* step_dev) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;
}

}  // namespace links

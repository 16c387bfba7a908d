// Occlusion-model training step pieces (reference train_occlusion_models.py:164-217).
#pragma once
#include "devdefs.cuh"

namespace links {

// pose[m] = [x*d, y*d, d] - root, d = head + depth, head col 0 forced to 0, NO clamp (:164-174).
// head_leg / head_torso are fp32 head outputs (ld = LINKS_HEAD_LD): joints 0..6 and 7..16.
__global__ void occ_lift_kernel(const float* __restrict__ x, const float* __restrict__ head_leg,
                                const float* __restrict__ head_torso, int M, float depth, float* __restrict__ pose) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = t / 17, j = t - m * 17;
  if (m >= M) return;
  const size_t mm = m;
  float dj = (j == 0) ? 0.f : (j < 7 ? head_leg[mm * LINKS_HEAD_LD + j] : head_torso[mm * LINKS_HEAD_LD + (j - 7)]);
  dj += depth;
  const float d0 = depth;   // root: head forced to 0
  const float x0 = x[mm * 34], y0 = x[mm * 34 + 17];
  pose[mm * 51 + j] = x[mm * 34 + j] * dj - x0 * d0;
  pose[mm * 51 + 17 + j] = x[mm * 34 + 17 + j] * dj - y0 * d0;
  pose[mm * 51 + 34 + j] = dj - d0;
}

// pose_out = Ry((u - 0.5) * 1.99 * pi) @ pose   (:213-217; utils/rotation_conversions.py:30-31)
__global__ void occ_rotate_y_kernel(const float* __restrict__ pose, const float* __restrict__ u, int M,
                                    float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = t / 17, j = t - m * 17;
  if (m >= M) return;
  const size_t mm = m;
  float s, c;
  sincosf((u[m] - 0.5f) * (1.99f * 3.14159265358979323846f), &s, &c);
  const float X = pose[mm * 51 + j], Y = pose[mm * 51 + 17 + j], Z = pose[mm * 51 + 34 + j];
  out[mm * 51 + j] = c * X + s * Z;
  out[mm * 51 + 17 + j] = Y;
  out[mm * 51 + 34 + j] = -s * X + c * Z;
}

// loss_sum += sum_m sum_c (pred - target)^2 ; g = scale * 2 (pred - target) as bf16 [M,64] (+ transposed).
// target[m][c] = pose[m*51 + tidx[c]]  (:176-183, :203-210).  One warp per row, lane strides the columns.
__global__ void __launch_bounds__(128) occ_mse_kernel(const float* __restrict__ pred, int ld_pred,
                                                      const float* __restrict__ pose, const int* __restrict__ tidx,
                                                      int n_out, int M, float scale, float* loss_sum,
                                                      __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ gT,
                                                      int ldT, int colT0) {
  __shared__ float s_part[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 4 + warp;
  float acc = 0.f;
  if (m < M) {
    const size_t mm = m;
    for (int c = lane; c < n_out; c += 32) {
      const float d = pred[mm * ld_pred + c] - pose[mm * 51 + tidx[c]];
      acc += d * d;
      const __nv_bfloat16 h = __float2bfloat16_rn(scale * 2.f * d);
      g[mm * 64 + c] = h;
      if (gT) gT[static_cast<size_t>(c) * ldT + colT0 + m] = h;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(LINKS_FULL_MASK, acc, o);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss_sum, s_part[0] + s_part[1] + s_part[2] + s_part[3]);
}

}  // namespace links

// Input pipeline on the device: the dataset classes' transpose-flatten of the raw [n, 17, 2] key-points
// (reference utils/h36m_dataset_class.py:25-27) fused with normalize_head / normalize_head_test
// (reference utils/helpers.py:198-207, 222-230):
//     p = raw.transpose(0, 2, 1).reshape(-1, 34);  p -= p[root];  scale_i = |p_i[joint 0] - p_i[joint 10]|;
//     out = p / mean_i(scale_i) * 0.1                      (normalize_head: the mean is over the WHOLE array)
//     out = p / fixed_scale * 0.1                          (normalize_head_test*)
// Two passes over HBM for the data-dependent scale (the mean must be known before anything can be scaled), one for the fixed
// one.  One thread per pose; rows are 136 B in, 136 B out.
#pragma once
#include "devdefs.cuh"

namespace links {

// pass 1: centre + transpose into out (unscaled), accumulate sum of head distances (double, one atomic per block)
__global__ void normalize_head_center_kernel(const float* __restrict__ raw, int n, int root, int transposed_input,
                                             float* __restrict__ out, double* __restrict__ dist_sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  if (i < n) {
    float x[17], y[17];
    const float* r = raw + static_cast<size_t>(i) * 34;
    if (transposed_input) {         // already [x0..x16, y0..y16]
#pragma unroll
      for (int j = 0; j < 17; ++j) { x[j] = r[j]; y[j] = r[17 + j]; }
    } else {                        // raw [17, 2]
#pragma unroll
      for (int j = 0; j < 17; ++j) { x[j] = r[2 * j]; y[j] = r[2 * j + 1]; }
    }
    float rx = 0.f, ry = 0.f;
#pragma unroll
    for (int j = 0; j < 17; ++j) if (j == root) { rx = x[j]; ry = y[j]; }
    float* o = out + static_cast<size_t>(i) * 34;
#pragma unroll
    for (int j = 0; j < 17; ++j) { x[j] -= rx; y[j] -= ry; o[j] = x[j]; o[17 + j] = y[j]; }
    const float dx = x[0] - x[10], dy = y[0] - y[10];
    d = static_cast<double>(sqrtf(dx * dx + dy * dy));
  }
  __shared__ double s[256];
  s[threadIdx.x] = d;
  __syncthreads();
  for (int k = blockDim.x / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0 && dist_sum != nullptr) atomicAdd(dist_sum, s[0]);
}

// pass 2: out *= 0.1 / scale, scale = fixed_scale > 0 ? fixed_scale : dist_sum / n
__global__ void normalize_head_scale_kernel(float* __restrict__ out, size_t total, int n, float fixed_scale,
                                            const double* __restrict__ dist_sum) {
  const double scale = fixed_scale > 0.f ? static_cast<double>(fixed_scale) : (*dist_sum / static_cast<double>(n));
  const float f = static_cast<float>(0.1 / scale);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) out[i] *= f;
}

}  // namespace links

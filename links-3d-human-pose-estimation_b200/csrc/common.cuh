// Shared helpers for the links_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/links_b200.h"

#define LINKS_CHECK_PTR(p) do { if ((p) == nullptr) return LINKS_E_ARG; } while (0)
#define LINKS_CHECK_ALIGN16(p) do { if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) return LINKS_E_ALIGN; } while (0)

// kernels launched through the C ABI since the library was loaded (links_launch_count; bench.py's `gpu_launches`)
extern unsigned long long g_links_kernel_launches;

static inline int links_launch_status(int n_kernels = 1) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) g_links_kernel_launches += static_cast<unsigned long long>(n_kernels);
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

static inline cudaStream_t links_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ __forceinline__ float links_leaky(float v) { return v > 0.f ? v : 0.01f * v; }

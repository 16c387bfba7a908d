// Device-side includes.  The non-tensor-core kernels (*.cuh) are written in plain CUDA C++ and are
// also compiled by tests/hostsim (g++, a thread-per-CUDA-thread shim) so their math can be checked
// against the oracle without a GPU.  The shim is test infrastructure; the product always runs the
// nvcc-compiled sm_100a code.
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifdef LINKS_HOSTSIM
#include "cuda_shim.h"
#define LINKS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(hostsim::dyn_smem())
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define LINKS_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#endif
#include "../../include/links_b200.h"
#define LINKS_FULL_MASK 0xffffffffu

// Chunk staging for the streaming (thread-per-pose) kernels: a block owns `kPosesPerBlock` consecutive rows of a few
// row-major tensors, i.e. one CONTIGUOUS range of HBM per tensor.  Full chunks are moved by the bulk-copy engine
// (cp.async.bulk, completion on an mbarrier: no load/store instructions, tens of KB in flight per SM); ragged tail
// chunks are copied cooperatively.  A small software pipeline overlaps the copy of chunk i+1 with the math of chunk i.
#pragma once
#include "devdefs.cuh"
#ifndef LINKS_HOSTSIM
#include "tc_ptx.cuh"
#endif

namespace links {

// Cooperative LINEAR copy of `count` floats starting at g (16-byte aligned) into shared memory.
__device__ __forceinline__ void stage_rows(const float* __restrict__ g, size_t count, float* s) {
  const size_t n4 = count >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* s4 = reinterpret_cast<float4*>(s);
  for (size_t i = threadIdx.x; i < n4; i += blockDim.x) s4[i] = g4[i];
  for (size_t ee = (n4 << 2) + threadIdx.x; ee < count; ee += blockDim.x) s[ee] = g[ee];
}

struct StageReq {
  float* s;
  const float* g;
  uint32_t count;   // floats
};

__device__ __forceinline__ void stage_bars_init(uint64_t* bars, int n) {
#ifndef LINKS_HOSTSIM
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#else
  (void)bars; (void)n;
#endif
  __syncthreads();
}

// bulk == true: every request is a multiple of 16 bytes from a 16-byte aligned address (block-uniform).
template <int NT>
__device__ __forceinline__ void stage_issue(const StageReq (&rq)[NT], uint64_t* bar, bool bulk) {
#ifndef LINKS_HOSTSIM
  if (bulk) {
    if (threadIdx.x == 0) {
      uint32_t total = 0;
#pragma unroll
      for (int t = 0; t < NT; ++t) total += rq[t].count * 4u;
      const uint32_t b = smem_u32(bar);
      mbar_expect_tx(b, total);
#pragma unroll
      for (int t = 0; t < NT; ++t) tma_load_1d(smem_u32(rq[t].s), rq[t].g, rq[t].count * 4u, b);
    }
    return;
  }
#else
  (void)bar; (void)bulk;
#endif
#pragma unroll
  for (int t = 0; t < NT; ++t) stage_rows(rq[t].g, rq[t].count, rq[t].s);
}

// Wait for the bulk copies of one stage (parity = number of earlier bulk uses of this barrier, mod 2), then make the
// staged data (either path) visible to the whole block.
__device__ __forceinline__ void stage_wait(uint64_t* bar, uint32_t parity, bool bulk) {
#ifndef LINKS_HOSTSIM
  if (bulk) mbar_wait(smem_u32(bar), parity);
#else
  (void)bar; (void)parity; (void)bulk;
#endif
  __syncthreads();
}

// Software pipeline over the chunks blockIdx.x, blockIdx.x + gridDim.x, ...:
//   issue(chunk, stage) -> bool bulk   starts the copies of a chunk into stage buffers
//   compute(chunk, stage)              runs after the data of that stage has landed
template <int kStages, class Issue, class Compute>
__device__ __forceinline__ void chunk_pipeline(int nchunks, uint64_t* bars, Issue issue, Compute compute) {
  uint32_t phase = 0;        // bit s: parity of stage s's barrier
  uint32_t bulk_mask = 0;    // bit s: the chunk in flight in stage s went through the bulk engine
  const int step = static_cast<int>(gridDim.x);
  int chunk = static_cast<int>(blockIdx.x);
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    const int c = chunk + s * step;
    if (c < nchunks) bulk_mask = (bulk_mask & ~(1u << s)) | (issue(c, s) ? (1u << s) : 0u);
  }
  int it = 0;
  for (; chunk < nchunks; chunk += step, ++it) {
    const int st = it % kStages;
    const int ahead = chunk + (kStages - 1) * step;
    if (ahead < nchunks) {
      const int sa = (it + kStages - 1) % kStages;     // last read kStages-1 iterations ago... released by the
      bulk_mask = (bulk_mask & ~(1u << sa)) | (issue(ahead, sa) ? (1u << sa) : 0u);   // trailing __syncthreads
    }
    const bool bulk = (bulk_mask >> st) & 1u;
    stage_wait(bars + st, (phase >> st) & 1u, bulk);
    if (bulk) phase ^= 1u << st;
    compute(chunk, st);
    __syncthreads();
  }
}

}  // namespace links

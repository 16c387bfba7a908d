// Host simulation shim (TEST INFRASTRUCTURE): runs CUDA kernels written in plain CUDA C++ on the CPU,
// one std::thread per CUDA thread, one block at a time.  Supports threadIdx/blockIdx/blockDim/gridDim,
// __syncthreads, warp shuffles, atomicAdd, static and dynamic __shared__, bf16 conversions.
// Only tests/ may use this; it is never part of the product path.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(x)
#define __restrict__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

namespace hostsim {
struct Block {
  int nthreads;
  std::unique_ptr<std::barrier<>> cta;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<uint64_t> xchg;  // [nwarps][32]
  std::vector<unsigned char> dyn;
};
inline Block*& cur_block() { static Block* b = nullptr; return b; }
inline thread_local int t_linear = 0;
inline unsigned char* dyn_smem() { return cur_block()->dyn.data(); }
inline std::mutex& atomic_mutex() { static std::mutex m; return m; }
}  // namespace hostsim

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void __syncthreads() { hostsim::cur_block()->cta->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { hostsim::cur_block()->warp[hostsim::t_linear >> 5]->arrive_and_wait(); }

template <class T>
inline T hostsim_shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  hostsim::Block* b = hostsim::cur_block();
  const int w = hostsim::t_linear >> 5, l = hostsim::t_linear & 31;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  b->xchg[w * 32 + l] = bits;
  b->warp[w]->arrive_and_wait();
  uint64_t r = b->xchg[w * 32 + (src & 31)];
  b->warp[w]->arrive_and_wait();
  T out;
  std::memcpy(&out, &r, sizeof(T));
  return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return hostsim_shfl(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return hostsim_shfl(v, (hostsim::t_linear & 31) ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) {
  const int l = hostsim::t_linear & 31;
  return hostsim_shfl(v, l + d < 32 ? l + d : l);
}

template <class T>
inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> g(hostsim::atomic_mutex());
  T old = *p;
  *p = old + v;
  return old;
}
template <class T> inline T __ldg(const T* p) { return *p; }

// ---- bf16 ----
struct __nv_bfloat16 { uint16_t x; };
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  __nv_bfloat16 h;
  if ((u & 0x7fffffffu) > 0x7f800000u) { h.x = 0x7fff; return h; }
  u += 0x7fffu + ((u >> 16) & 1u);
  h.x = static_cast<uint16_t>(u >> 16);
  return h;
}
inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = static_cast<uint32_t>(h.x) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// ---- math ----
inline void sincosf(float a, float* s, float* c) { *s = std::sin(a); *c = std::cos(a); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
using std::fabs; using std::fmax; using std::fmin; using std::sqrt;

namespace hostsim {
// Run `body()` for every thread of every block (blocks sequential, threads concurrent).
inline void launch(dim3 grid, dim3 block, size_t dyn_bytes, const std::function<void()>& body) {
  gridDim = grid;
  blockDim = block;
  const int nthreads = block.x * block.y * block.z;
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk;
        blk.nthreads = nthreads;
        blk.cta = std::make_unique<std::barrier<>>(nthreads);
        for (int w = 0; w < nwarps; ++w) {
          int cnt = std::min(32, nthreads - w * 32);
          blk.warp.push_back(std::make_unique<std::barrier<>>(cnt));
        }
        blk.xchg.assign(static_cast<size_t>(nwarps) * 32, 0);
        blk.dyn.assign(dyn_bytes + 16, 0);
        cur_block() = &blk;
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
          th.emplace_back([&, t] {
            t_linear = t;
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            blockIdx = dim3(bx, by, bz);
            body();
            blk.warp[t >> 5]->arrive_and_drop();
            blk.cta->arrive_and_drop();
          });
        }
        for (auto& x : th) x.join();
        cur_block() = nullptr;
      }
}
}  // namespace hostsim

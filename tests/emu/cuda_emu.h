// cuda_emu.h — TEST INFRASTRUCTURE: the small CUDA vocabulary the non-tensor-core kernels of ssd_b200/csrc use, mapped
// onto host threads so that a kernel's *source* (not a restatement of it) can be executed on a machine without a GPU:
// one OS thread per CUDA thread, __syncthreads = a per-CTA barrier, warp shuffles = a per-warp exchange + barrier,
// global atomics = host atomics.  Host memory is coherent and sequentially consistent enough for these tests, so this
// checks indexing, control flow, barrier placement (a shuffle or barrier that not all threads reach deadlocks here just
// as it is undefined on the device) and arithmetic — not the device memory model, cache behaviour or performance.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

// ---- bf16 (round-to-nearest-even, like cuda_bf16.h) ----
struct __nv_bfloat16 { uint16_t v; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return {(uint16_t)((u >> 16) | 0x40)};  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return {(uint16_t)(u >> 16)};
}
inline __nv_bfloat16 __float2bfloat16(float f) { return __float2bfloat16_rn(f); }
inline float __bfloat162float(__nv_bfloat16 b) {
  uint32_t u = (uint32_t)b.v << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline __nv_bfloat16 __ushort_as_bfloat16(unsigned short s) { return {s}; }
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return {__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
inline float2 __bfloat1622float2(__nv_bfloat162 v) { return {__bfloat162float(v.x), __bfloat162float(v.y)}; }

// ---- math ----
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __expf(float x) { return std::exp(x); }
inline float __logf(float x) { return std::log(x); }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
using std::max;
using std::min;

namespace emu {
struct Warp {
  uint32_t slot[32];
  std::barrier<> bar{32};
};
struct Cta {
  explicit Cta(int threads, size_t dyn) : bar(threads), dyn_smem(dyn + 64), warps((threads + 31) / 32) {
    for (auto& w : warps) w = std::make_unique<Warp>();
  }
  std::barrier<> bar;
  std::vector<unsigned char> dyn_smem;
  std::vector<std::unique_ptr<Warp>> warps;
  std::mutex mu;
  std::map<int, std::vector<unsigned char>> statics;
  std::map<int, std::unique_ptr<std::barrier<>>> named;  // bar.sync id, count
};
struct ThreadCtx {
  Cta* cta = nullptr;
};
inline thread_local ThreadCtx tctx;
inline void* dyn_smem() {
  auto p = reinterpret_cast<uintptr_t>(tctx.cta->dyn_smem.data());
  return reinterpret_cast<void*>((p + 63) & ~uintptr_t(63));
}
template <typename T>
T* static_smem(size_t n, int key) {
  std::lock_guard<std::mutex> g(tctx.cta->mu);
  auto& v = tctx.cta->statics[key];
  if (v.empty()) v.resize(n * sizeof(T) + 64);
  return reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(v.data()) + 63) & ~uintptr_t(63));
}
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::tctx.cta->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::tctx.cta->warps[threadIdx.x >> 5]->bar.arrive_and_wait(); }
namespace emu {
// bar.sync id, count: a barrier among `count` threads of the CTA (the same `count` every time for a given id)
inline void named_barrier(int id, int count) {
  Cta* c = tctx.cta;
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> g(c->mu);
    auto& slot = c->named[id];
    if (!slot) slot = std::make_unique<std::barrier<>>(count);
    b = slot.get();
  }
  b->arrive_and_wait();
}
}  // namespace emu
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __trap() {
  std::fprintf(stderr, "[cuda_emu] __trap()\n");
  std::abort();
}
inline long long clock64() {
  // ~2 "cycles" per ns keeps the kernels' cycle-count time-outs in the seconds range
  return 2 * std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  emu::Warp& w = *emu::tctx.cta->warps[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  std::memcpy(&w.slot[lane], &v, 4);
  w.bar.arrive_and_wait();
  T r;
  std::memcpy(&r, &w.slot[lane ^ lane_mask], 4);
  w.bar.arrive_and_wait();
  return r;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
inline T __ldcg(const T* p) { return *p; }

namespace emu {
// run kernel(params) on grid.x * grid.y CTAs of `block` host threads each.  `wave` > 0 runs the CTAs in batches of that
// size (only valid for kernels whose CTAs never wait for one another: ticket / last-block patterns are fine, device-wide
// barriers are not); wave == 0 makes every CTA co-resident.
template <typename Kernel, typename Params>
void launch(Kernel kernel, const Params& params, dim3 grid, int block, size_t dyn_smem_bytes, int wave = 0) {
  if (block % 32) {
    std::fprintf(stderr, "[cuda_emu] block size must be a multiple of 32\n");
    std::abort();
  }
  const int total = (int)(grid.x * grid.y);
  if (wave <= 0) wave = total;
  for (int first = 0; first < total; first += wave) {
    const int n = std::min(wave, total - first);
    std::vector<std::unique_ptr<Cta>> ctas;
    for (int b = 0; b < n; ++b) ctas.push_back(std::make_unique<Cta>(block, dyn_smem_bytes));
    std::vector<std::thread> ts;
    ts.reserve((size_t)n * block);
    for (int b = 0; b < n; ++b)
      for (int t = 0; t < block; ++t)
        ts.emplace_back([&, b, t] {
          tctx.cta = ctas[b].get();
          threadIdx.x = t;
          blockIdx.x = (unsigned)(first + b) % grid.x;
          blockIdx.y = (unsigned)(first + b) / grid.x;
          blockDim.x = block;
          gridDim = grid;
          kernel(params);
        });
    for (auto& t : ts) t.join();
  }
}
template <typename Kernel, typename Params>
void launch(Kernel kernel, const Params& params, int grid, int block, size_t dyn_smem_bytes) {
  dim3 g;
  g.x = (unsigned)grid;
  launch(kernel, params, g, block, dyn_smem_bytes, 0);
}
}  // namespace emu

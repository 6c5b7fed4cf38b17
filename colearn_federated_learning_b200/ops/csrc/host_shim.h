// SIMT-on-CPU shim: compiles the warp-synchronous CUDA kernels of this directory (mlp_persistent.cu + mlp_v2.inc) with
// g++ and runs them with one OS thread per CUDA thread, so that the CPU test-suite executes the *same source* the GPU
// runs — the flagship persistent-MLP kernel included (tests only; never used to train).
//
//   thread block      one block at a time, blockDim.x std::threads
//   __syncthreads()   std::barrier over the live threads of the block (a thread that returns from the kernel drops out,
//                     like an exited CUDA thread)
//   __shfl_*_sync()   per-warp exchange slots between two warp barriers (all 32 lanes of a warp must take part, which
//                     is also what the kernels assume on the device)
//   __shared__        function-local static (blocks run one after the other; contents are as undefined between blocks
//                     as real shared memory); dynamic shared memory: COLEARN_DYN_SMEM -> a per-block heap buffer
//   ld.acquire / st.release .sys   __atomic builtins
//   kernel<<<g, b, s, stream>>>(args)   COLEARN_LAUNCH(kernel, g, b, s, stream, args) -> colearn_shim::launch(...)
//
// Include this header BEFORE the kernel source and define COLEARN_HOST_SHIM.
#pragma once
#include <cuda_runtime.h>   // under g++ the CUDA qualifiers (__device__, __global__, __forceinline__, ...) are harmless

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

namespace colearn_shim {

struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
};

struct Block {
  explicit Block(int n, size_t smem_bytes) : nthreads(n), bar(n), dyn((smem_bytes + 63) / 4 + 16, 0.f) {
    for (int w = 0; w < (n + 31) / 32; ++w) {
      const int lanes = (w * 32 + 32 <= n) ? 32 : n - w * 32;
      warp_bar.emplace_back(std::make_unique<std::barrier<>>(lanes));
    }
    slots.assign((size_t)((n + 31) / 32) * 32, 0ull);
  }
  int nthreads;
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<unsigned long long> slots;   // one 8-byte exchange slot per lane
  std::vector<float> dyn;                  // dynamic shared memory (16-byte aligned start, see dyn_smem())
};

inline thread_local Dim3 t_threadIdx;
inline thread_local Block* t_block = nullptr;
inline Dim3 g_blockIdx, g_blockDim, g_gridDim;

inline float* dyn_smem() {
  uintptr_t p = reinterpret_cast<uintptr_t>(t_block->dyn.data());
  return reinterpret_cast<float*>((p + 15) & ~(uintptr_t)15);
}

inline void syncthreads() { t_block->bar.arrive_and_wait(); }
inline void syncwarp() { t_block->warp_bar[t_threadIdx.x / 32]->arrive_and_wait(); }

template <class T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  Block* b = t_block;
  const int w = t_threadIdx.x / 32, l = t_threadIdx.x % 32;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  b->slots[(size_t)w * 32 + l] = bits;
  b->warp_bar[w]->arrive_and_wait();
  const unsigned long long got = b->slots[(size_t)w * 32 + (src_lane & 31)];
  b->warp_bar[w]->arrive_and_wait();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}

// Runs `body()` once per CUDA thread, block after block.
template <class F>
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, F&& body) {
  g_gridDim = Dim3{grid, 1, 1};
  g_blockDim = Dim3{block, 1, 1};
  for (unsigned b = 0; b < grid; ++b) {
    g_blockIdx = Dim3{b, 1, 1};
    Block blk((int)block, smem_bytes);
    std::vector<std::thread> threads;
    threads.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      threads.emplace_back([&, t] {
        t_threadIdx = Dim3{t, 1, 1};
        t_block = &blk;
        body();
        // an exited thread no longer takes part in barriers (CUDA semantics)
        blk.bar.arrive_and_drop();
        blk.warp_bar[t / 32]->arrive_and_drop();
      });
    for (auto& th : threads) th.join();
  }
}

}  // namespace colearn_shim

#define threadIdx (::colearn_shim::t_threadIdx)
#define blockIdx (::colearn_shim::g_blockIdx)
#define blockDim (::colearn_shim::g_blockDim)
#define gridDim (::colearn_shim::g_gridDim)

inline void __syncthreads() { ::colearn_shim::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::colearn_shim::syncwarp(); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int = 32) {
  return ::colearn_shim::shfl_from(v, (int)(::colearn_shim::t_threadIdx.x % 32) ^ lane_mask);
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane, int = 32) {
  return ::colearn_shim::shfl_from(v, src_lane);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32) {
  const int l = (int)(::colearn_shim::t_threadIdx.x % 32);
  return ::colearn_shim::shfl_from(v, l + (int)delta < 32 ? l + (int)delta : l);
}
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float __fdividef(float a, float b) { return a / b; }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// runtime calls the launchers make: nothing to configure on the host
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)
#define cudaGetLastError() cudaSuccess

#define COLEARN_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::colearn_shim::dyn_smem())
#define COLEARN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::colearn_shim::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })
#define COLEARN_KERNEL_NAME(...) __VA_ARGS__

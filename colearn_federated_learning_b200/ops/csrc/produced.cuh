// Producer-side half of the fused  wgrad GEMM -> FedAvg reduce  (docs/KERNELS.md §3b; the consumer-side half, fused
// broadcast -> GEMM, is GemmEpilogue::ready_flags).
//
// During the LAST local SGD step of a round every kernel that writes final parameters into the fp32 work arena reports
// what it finished: the fused-SGD epilogue of the wgrad GEMM per (32 rows x BN/2 columns) warp block, everything else
// (biases, padded edge layers) through produced_mark_kernel.  Reports are element counts per arena chunk; whoever
// brings a chunk to its full length publishes the round's epoch in the chunk OWNER's table (rank c % world, the rank
// that reduces chunk c in twoshot_fedavg_kernel), which is polling it on a few CTAs next to the GEMMs: the reduce of the
// last layers' parameters crosses NVLink while the backward pass of the earlier layers is still computing.
//
// Memory model: the reporting thread has synchronised with the threads that stored the data (__syncwarp / kernel
// boundary), then fence.sys -> atomicAdd (relaxed).  The thread whose add completes the chunk observed every earlier
// add, fences again (acquire side of the fence-fence pattern, cumulative) and stores the flag with st.release.sys; the
// owner reads it with ld.acquire.sys before its P2P loads.
#pragma once
#include <stdint.h>

#include "colearn_kernels.h"   // struct ProducedSignal

namespace colearn {


#ifdef COLEARN_HOST_SHIM
inline void produced_st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline uint32_t produced_ld_relaxed(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
#define COLEARN_PRODUCED_FN inline
#else
__device__ __forceinline__ void produced_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t produced_ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#define COLEARN_PRODUCED_FN __device__ __forceinline__
#endif

// `len` more elements of chunk c are final.  Caller: fence.sys after having synchronised with the storing threads.
COLEARN_PRODUCED_FN void produced_add(const ProducedSignal* s, int64_t c, uint32_t len) {
  const int64_t lo = c << s->chunk_shift;
  int64_t full = (int64_t)1 << s->chunk_shift;
  if (s->n - lo < full) full = s->n - lo;
  const uint32_t old = atomicAdd(s->count + c, len);
  if ((int64_t)old + len == full) {
    __threadfence_system();
    atomicExch(s->count + c, 0u);          // next contributions come from the next round's last step (stream order)
    produced_st_release_sys(s->flags[c % s->world] + (int64_t)s->rank * s->n_chunks + c, produced_ld_relaxed(s->epoch_ptr) + s->epoch_add);
  }
}

// rows [row0, row0 + nrows) x columns [col0, col0 + width) of a row-major [*, ld] matrix that starts at arena element
// `off`: run-length over the chunks the block touches (rows of one block map to non-decreasing chunk indices)
COLEARN_PRODUCED_FN void produced_block(const ProducedSignal* s, int64_t off, int row0, int nrows, int ld, int col0, int width) {
  int64_t cur = -1;
  uint32_t acc = 0;
  for (int r = 0; r < nrows; ++r) {
    int64_t b = off + (int64_t)(row0 + r) * ld + col0;
    const int64_t e = b + width;
    while (b < e) {
      const int64_t c = b >> s->chunk_shift;
      const int64_t c_end = (c + 1) << s->chunk_shift;
      const int64_t seg = (e < c_end ? e : c_end) - b;
      if (c != cur) {
        if (acc) produced_add(s, cur, acc);
        cur = c;
        acc = 0;
      }
      acc += (uint32_t)seg;
      b += seg;
    }
  }
  if (acc) produced_add(s, cur, acc);
}

}  // namespace colearn

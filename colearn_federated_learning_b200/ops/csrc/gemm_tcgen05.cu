// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] . B[N,K]^T   (bf16 in, fp32 accumulate)
//
// Serves the Linear layers of the wide models: forward (bias+ReLU epilogue),
// dgrad (ReLU-mask epilogue) and wgrad (fused bias-grad column sums + fused SGD update of the fp32
// master weights and refresh of the bf16 shadows W / W^T).  All three are expressed as the same
// "both operands K-major" product by keeping transposed bf16 copies that the producing epilogue
// writes for free (out_bf16_t: lanes of a warp hold 32 consecutive rows, so the transposed store is
// naturally coalesced).
//
// Fused broadcast -> GEMM: when ep.ready_flags is set, the TMA producer warp polls
// (ld.acquire.sys) the per-chunk ready flag of the B rows it is about to fetch — flags that the
// two-shot FedAvg kernel on a PEER GPU raises after writing the new weights over NVLink — then
// issues fence.proxy.async before the cp.async.bulk.tensor, so the first layer's GEMM starts on
// the tiles that have landed while later chunks are still in flight.
//
// Structure (persistent, 1 CTA / SM, 192 threads):
//   warp 0   : TMA producer  (1 elected lane)     smem ring of kStages x {A 128x64, B 128x64} bf16, SW128
//   warp 1   : TMEM alloc + MMA issuer (1 lane)   tcgen05.mma.cta_group::1.kind::f16, UMMA 128x128x16
//   warps 2-9: epilogue (2 per TMEM lane quarter)  tcgen05.ld 32x32b.x32 -> regs -> fused epilogue -> global
//   TMEM     : 2 accumulator stages x 128 columns (epilogue of tile i overlaps mainloop of tile i+1)
#include "colearn_kernels.h"
#include "produced.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>
#include <string>
#include <unordered_map>

namespace colearn {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int UMMA_K = 16;
constexpr int kAccStages = 2;
constexpr int kEpilogueWarps = 8;                 // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kThreads = 64 + 32 * kEpilogueWarps;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kMaxStages = 6;
// Tile config.  BN = 256 halves the B-operand smem traffic per FLOP (128x128x16 UMMAs sit exactly at the
// 128 B/clk smem limit: 8 KB of operands per 64-cycle instruction; 128x256x16 needs 12 KB per 128 cycles).
template <int BN_>
struct Cfg {
  static constexpr int BN = BN_;
  static constexpr int kStages = BN_ == 256 ? 4 : 6;
  static constexpr int kTmemCols = kAccStages * BN_;  // 256 or 512 (all of TMEM)
  static constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN_ * BK * 2;
  static constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

std::string g_last_error;

// Opt-in limit of dynamic shared memory: operand stages + barriers
inline cudaError_t configure_smem(const void* kernel, int base_bytes, int dev) {
#ifdef COLEARN_HOST_SHIM
  (void)kernel; (void)base_bytes; (void)dev;
  return cudaSuccess;
#else
  (void)dev;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, base_bytes);
#endif
}

#ifdef COLEARN_HOST_SHIM
#include "tcgen05_host_model.h"   // functional CPU model of mbarrier / TMA / tcgen05 / TMEM (tests)
#else
// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
  return done;
}
// Watchdog: a wait that lasts longer than this can only be a protocol bug (e.g. a TMA box whose byte count does not match
// the expect_tx of its stage) — trap, so that the launch fails with an error instead of hanging the GPU.
constexpr unsigned long long kMbarTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if (mbar_try_wait(addr, parity)) return;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > kMbarTimeoutNs) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  // 4-D tiled box (implicit-GEMM convolution: (channel, w, h, image) of an NHWC activation); coordinates may be
  // negative / past the edge, those elements are zero-filled
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar, uint16_t cta_mask) {
  // the tile lands at the same CTA-relative smem offset in every CTA of cta_mask and completes tx bytes on the
  // mbarrier at the same offset in each of them
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}


// ---- cta_group::2 (two SMs on one 256-row tile) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint32_t mbar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// the three fences / hints the kernels issue inline (macros, so that the statements stay exactly where they were)
#define tmap_prefetch(t) asm volatile("prefetch.tensormap [%0];" ::"l"(t) : "memory")
#define fence_mbarrier_init() asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory")
#define fence_proxy_async() asm volatile("fence.proxy.async.global;" ::: "memory")

#endif  // COLEARN_HOST_SHIM

// K-major, 128B-swizzled smem operand descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, ignored for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major, 128B-swizzled operand (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): the
// operand sits in shared memory the way a row-major [K rows, MN columns] global matrix lands through TMA boxes of
// [64 rows x 64 columns]: inside a box, reduction row r is the 128-byte line r (8-line groups = one 1024-byte swizzle
// atom, SBO = 1024 between groups); consecutive 64-column blocks are consecutive boxes (LBO = box size = 8192).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t kIdescMnMajorA = 1u << 15, kIdescMnMajorB = 1u << 16;   // a_major / b_major = MN
constexpr uint32_t kMnBoxBytes = 64 * 64 * 2;                   // one [64 x 64] bf16 TMA box

// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b BF16 (1<<7, 1<<10), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Work decomposition (tile rasterisation, split-K slices): host+device functions in conv_ops.cuh (GemmSched), shared
// with the CPU test-suite
using convops::split_kb;
using convops::work_to_tile;

struct SharedBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[kAccStages];
  uint64_t tmem_empty[kAccStages];
  uint32_t tmem_base;
};

static_assert(sizeof(SharedBarriers) <= 256, "the barriers own 256 bytes behind the operand stages");

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Fused epilogue for one 32-column chunk of one accumulator row per thread (row = m0 + 32*q + lane).
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const GemmEpilogue& ep, int row, int col, int lane,
                                               int q, int m0, int M, int N) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (ep.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += __ldg(ep.bias + col + j);
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (ep.relu_mask != nullptr) {
    const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ep.relu_mask) + (size_t)row * N + col);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 mv = __ldg(mp + g);
      const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        // bf16 > 0  <=>  sign bit clear and magnitude non-zero
        const uint32_t lo = w[h] & 0xFFFFu, hi = w[h] >> 16;
        if (!(lo != 0 && !(lo & 0x8000u))) f[g * 8 + h * 2] = 0.f;
        if (!(hi != 0 && !(hi & 0x8000u))) f[g * 8 + h * 2 + 1] = 0.f;
      }
    }
  }
  if (ep.addend != nullptr) {
    const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ep.addend) + (size_t)row * N + col);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 av = __ldg(ap + g);
      const uint32_t w[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        f[g * 8 + h * 2] += __uint_as_float(w[h] << 16);              // bf16 -> fp32: the bits are the high half
        f[g * 8 + h * 2 + 1] += __uint_as_float(w[h] & 0xFFFF0000u);
      }
    }
  }
  if (ep.colsum != nullptr) {
    // bias gradient: per-(32-row block) partial column sums, plain coalesced stores (no atomics);
    // the consumer (bias_sgd_from_partials) adds the M/32 partial rows
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float s = f[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == j) mine = s;
    }
    ep.colsum[(size_t)((m0 >> 5) + q) * N + col + lane] = mine;
  }
  if (ep.sgd_master != nullptr) {
    float4* mp = reinterpret_cast<float4*>(ep.sgd_master + (size_t)row * N + col);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float4 w = mp[g];
      w.x = fmaf(-ep.sgd_lr, f[g * 4 + 0], w.x); w.y = fmaf(-ep.sgd_lr, f[g * 4 + 1], w.y);
      w.z = fmaf(-ep.sgd_lr, f[g * 4 + 2], w.z); w.w = fmaf(-ep.sgd_lr, f[g * 4 + 3], w.w);
      mp[g] = w;
      f[g * 4 + 0] = w.x; f[g * 4 + 1] = w.y; f[g * 4 + 2] = w.z; f[g * 4 + 3] = w.w;  // f now holds the new weights
    }
    if (ep.sgd_shadow != nullptr) {
      uint4* sp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.sgd_shadow) + (size_t)row * N + col);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        sp[g] = make_uint4(pack2(f[g * 8], f[g * 8 + 1]), pack2(f[g * 8 + 2], f[g * 8 + 3]),
                           pack2(f[g * 8 + 4], f[g * 8 + 5]), pack2(f[g * 8 + 6], f[g * 8 + 7]));
    }
    if (ep.sgd_shadow_t != nullptr) {
      __nv_bfloat16* tp = reinterpret_cast<__nv_bfloat16*>(ep.sgd_shadow_t);
#pragma unroll
      for (int j = 0; j < 32; ++j) tp[(size_t)(col + j) * M + row] = __float2bfloat16(f[j]);
    }
  } else {
    if (ep.out_f32 != nullptr) {
      float4* op = reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * N + col);
#pragma unroll
      for (int g = 0; g < 8; ++g) op[g] = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
    }
    if (ep.out_bf16 != nullptr) {
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_bf16) + (size_t)row * N + col);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        op[g] = make_uint4(pack2(f[g * 8], f[g * 8 + 1]), pack2(f[g * 8 + 2], f[g * 8 + 3]),
                           pack2(f[g * 8 + 4], f[g * 8 + 5]), pack2(f[g * 8 + 6], f[g * 8 + 7]));
    }
    if (ep.out_bf16_t != nullptr) {
      __nv_bfloat16* tp = reinterpret_cast<__nv_bfloat16*>(ep.out_bf16_t);
#pragma unroll
      for (int j = 0; j < 32; ++j) tp[(size_t)(col + j) * M + row] = __float2bfloat16(f[j]);
    }
  }
}


// CL = 2: thread-block cluster of two CTAs working on vertically adjacent tiles (same n-block).  Each CTA fetches
// its own A tile and HALF of the shared B tile, multicasting that half into both CTAs' shared memory, which cuts
// the L2->SM operand traffic per CTA from 48 KB to 32 KB per k-block (the 128x256 tile is L2-bandwidth bound).
// MN = true ("both operands MN-major"): A is [K, M] and B is [K, N] row-major, i.e. C = A^T B with the reduction
// running over ROWS — the conv wgrad dW[Cout, k] = sum_pixels dz[pixel, Cout] * col[pixel, k] reads dz and col as they
// are, without the two transposes a K-major kernel needs.  A stage holds 64 reduction rows; every 64-column block of
// the tile is one [64 x 64] TMA box (columns past the matrix edge are zero-filled by TMA = the Cout padding).
// BMN alone (A K-major [M, K], B [K, N] row-major): the conv dgrad dcol[pixel, k] = sum_co dz[pixel, co] * Wp[co, k]
// reads the packed weights Wp[Cout, K] as they are, so no W^T copy has to be kept in step with the SGD updates.
template <int BN, int CL, bool MN = false, bool BMN = MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    int M, int N, int K, GemmEpilogue ep) {
  using C = Cfg<BN>;
  const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
  constexpr int kStages = C::kStages, kTmemCols = C::kTmemCols;
  constexpr uint32_t kStageBytesA = C::kStageBytesA, kStageBytesB = C::kStageBytesB, kStageBytes = C::kStageBytes;
  COLEARN_DYN_SMEM_UNALIGNED(uint8_t, smem_raw);
  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kStageBytesA;
  SharedBarriers* bars = reinterpret_cast<SharedBarriers*>(smem + kStages * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = M / BM, n_tiles = N / BN, num_kb = K / BK;
  // work unit = one tile (CL == 1) or a vertical pair of tiles handled by the two CTAs of a cluster (CL == 2)
  const int m_units = m_tiles / CL;
  // split-K (CL == 1 only): work unit = (output tile, K slice); slices of one tile are adjacent work units, so they
  // run concurrently on different SMs and each stores a raw fp32 partial (see GemmEpilogue::split_k)
  const int S = (CL == 1 && ep.split_k > 1) ? ep.split_k : 1;
  const int total_work = m_units * n_tiles * S;
  const int work0 = blockIdx.x / CL, work_stride = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmap_a);
    tmap_prefetch(&tmap_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], CL); }
    for (int i = 0; i < kAccStages; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], kEpilogueWarps); }
    fence_mbarrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // the peer may multicast into my smem / arrive on my barriers from now on
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // programmatic dependent launch: everything above overlapped the predecessor's tail; nothing below may touch global
  // memory before the grids this launch depends on have completed
  if (ep.pdl) COLEARN_PDL_PROLOGUE();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = work0; tile < total_work; tile += work_stride) {
        int mu, nb;
        work_to_tile(tile / S, m_units, n_tiles, mu, nb);
        const int ks = tile % S;
        const int kb_lo = split_kb(num_kb, ks, S), kb_hi = split_kb(num_kb, ks + 1, S);
        const int m0 = (mu * CL + (int)crank) * BM, n0 = nb * BN;
        if (ep.ready_flags != nullptr) {
          const uint32_t want = ep.ready_epoch_ptr ? ld_acquire_sys(ep.ready_epoch_ptr) : ep.ready_epoch;
          // wait for the peer-written weight rows [n0, n0+BN) of this round, then make them
          // visible to the async proxy before TMA touches them
          const int64_t c_lo = (ep.ready_elem_offset + (int64_t)n0 * K) / ep.ready_chunk_elems;
          const int64_t c_hi = (ep.ready_elem_offset + (int64_t)(n0 + BN) * K - 1) / ep.ready_chunk_elems;
          for (int64_t c = c_lo; c <= c_hi; ++c)
            spin_wait_ge(ep.ready_flags + c, want, 64, "gemm_tcgen05: ready flag of a broadcast weight chunk");
          if (ep.bias != nullptr) {  // the layer's fp32 bias directly follows its weight in the flat arena
            const int64_t b_lo = (ep.ready_elem_offset + (int64_t)N * K + n0) / ep.ready_chunk_elems;
            const int64_t b_hi = (ep.ready_elem_offset + (int64_t)N * K + n0 + BN - 1) / ep.ready_chunk_elems;
            for (int64_t c = b_lo; c <= b_hi; ++c)
              spin_wait_ge(ep.ready_flags + c, want, 64, "gemm_tcgen05: ready flag of a broadcast weight chunk");
          }
          fence_proxy_async();
        }
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
          if (CL == 1 && !MN && ep.conv.mode == 1) {
            // implicit-GEMM convolution, forward / dgrad: the A tile is the activation box of this tap (conv_ops.cuh)
            const convops::ConvBox bx = convops::conv_kblock(ep.conv, kb, m0, n0);
            tma_load_4d(smem_a + stage * kStageBytesA, &tmap_a, bx.c, bx.w, bx.h, bx.n, &bars->full[stage]);
            if (BMN) {
              // dgrad against the packed weights Wp[co, (tap, ci)] themselves (MN-major B): the K index (co) runs over
              // ROWS, the N index over the columns of this tap — the same two numbers with their roles swapped
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(smem_b + stage * kStageBytesB + j * kMnBoxBytes, &tmap_b, bx.b_row + j * 64, bx.b_col, &bars->full[stage]);
            } else {
              tma_load_2d(smem_b + stage * kStageBytesB, &tmap_b, bx.b_col, bx.b_row, &bars->full[stage]);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(smem_a + stage * kStageBytesA + j * kMnBoxBytes, &tmap_a, m0 + j * 64, kb * BK, &bars->full[stage]);
          } else {
            tma_load_2d(smem_a + stage * kStageBytesA, &tmap_a, kb * BK, m0, &bars->full[stage]);
          }
          if (BMN && ep.conv.mode == 2) {
            // implicit-GEMM convolution, wgrad: every 64-column block (tap, 64 channels) of the B tile is an activation box
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              const convops::ConvBox bx = convops::conv_nblock(ep.conv, n0 / 64 + j, kb);
              tma_load_4d(smem_b + stage * kStageBytesB + j * kMnBoxBytes, &tmap_b, bx.c, bx.w, bx.h, bx.n, &bars->full[stage]);
            }
          } else if (BMN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(smem_b + stage * kStageBytesB + j * kMnBoxBytes, &tmap_b, n0 + j * 64, kb * BK, &bars->full[stage]);
          } else if (CL == 2) {
            // my half of the shared B tile -> both CTAs (tmap_b's box is BN/2 rows in this mode)
            tma_load_2d_mcast(smem_b + stage * kStageBytesB + crank * (kStageBytesB / 2), &tmap_b, kb * BK,
                              n0 + (int)crank * (BN / 2), &bars->full[stage], (uint16_t)0x3);
          } else {
            tma_load_2d(smem_b + stage * kStageBytesB, &tmap_b, kb * BK, n0, &bars->full[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN) | (MN ? kIdescMnMajorA : 0u) | (BMN ? kIdescMnMajorB : 0u);
      const uint32_t mn_lbo = ep.mn_lbo ? (uint32_t)ep.mn_lbo : kMnBoxBytes, mn_sbo = ep.mn_sbo ? (uint32_t)ep.mn_sbo : 1024u;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = work0; tile < total_work; tile += work_stride, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int ks = tile % S;
        const int kb_lo = split_kb(num_kb, ks, S), kb_hi = split_kb(num_kb, ks + 1, S);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          // K-major operand: +16 bf16 = 32 bytes inside the 128B swizzle atom per UMMA_K (+2 in addr>>4 units);
          // MN-major operand: +16 reduction rows = 16 lines of 128 bytes = two whole swizzle atoms (+128)
          const uint64_t da = MN ? make_smem_desc_mn(smem_u32(smem_a + stage * kStageBytesA), mn_lbo, mn_sbo)
                                 : make_smem_desc(smem_u32(smem_a + stage * kStageBytesA));
          const uint64_t db = BMN ? make_smem_desc_mn(smem_u32(smem_b + stage * kStageBytesB), mn_lbo, mn_sbo)
                                  : make_smem_desc(smem_u32(smem_b + stage * kStageBytesB));
          constexpr uint32_t kStepA = MN ? 128 : 2, kStepB = BMN ? 128 : 2;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(tmem_d, da + kStepA * k, db + kStepB * k, idesc, ((kb - kb_lo) | k) != 0);
          if (CL == 2) umma_commit_mcast(&bars->empty[stage], (uint16_t)0x3);  // both producers write into this stage
          else umma_commit(&bars->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->tmem_full[as]);
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;          // 0: columns [0, BN/2), 1: [BN/2, BN)
    int local = 0;
    for (int tile = work0; tile < total_work; tile += work_stride, ++local) {
      int mu, nb;
      work_to_tile(tile / S, m_units, n_tiles, mu, nb);
      const int m0 = (mu * CL + (int)crank) * BM, n0 = nb * BN;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      // split-K: this work unit's raw accumulator goes to slice (tile % S) of the fp32 partial buffer
      float* part = (S > 1) ? ep.split_out + ((size_t)(tile % S) * M + row) * N + n0 : nullptr;
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), v);
        tmem_ld_wait();
        if (S > 1) {
          float4* op = reinterpret_cast<float4*>(part + c0);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            op[g] = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                                __uint_as_float(v[g * 4 + 3]));
        } else {
          epilogue_chunk(v, ep, row, n0 + c0, lane, q, m0, M, N);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
      // fused wgrad -> FedAvg reduce: this warp's 32 x BN/2 block of the master matrix is final (the __syncwarp above
      // ordered the 32 lanes' stores before lane 0's fence)
      if (ep.produced != nullptr && lane == 0) {
        __threadfence_system();
        produced_block(ep.produced, ep.produced_elem_offset, m0 + q * 32, 32, N, n0 + half * (BN / 2), BN / 2);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();   // nobody may still multicast into / arrive on a CTA that exits
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
#ifdef COLEARN_HOST_SHIM
  return nullptr;
#endif
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int rows, cols, box_rows;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (std::hash<int>()(k.rows) * 1000003u) ^ (std::hash<int>()(k.cols) * 10007u) ^
           (std::hash<int>()(k.box_rows) * 131u);
  }
};

bool make_tmap(const void* ptr, int rows, int cols, int box_rows, CUtensorMap* out) {
  // row-major [rows, cols] bf16; box = [box_rows rows, 64 cols] (64 bf16 = one 128-byte swizzle row)
#ifdef COLEARN_HOST_SHIM
  {
    const long long dims[2] = {cols, rows}, strides[1] = {(long long)cols * 2};
    const int box[2] = {BK, box_rows};
    shim_encode_tiled(out, ptr, 2, dims, strides, box);
    return true;
  }
#endif
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  MapKey key{ptr, rows, cols, box_rows};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) { g_last_error = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_error = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r); return false; }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

bool make_tmap_4d(const void* ptr, int n_images, int H, int W, int C, int box_images, CUtensorMap* out) {
  // NHWC bf16 activation as (c, w, h, n); box = 64 channels x whole images (SW128: one box line = 64 channels = 128 bytes)
#ifdef COLEARN_HOST_SHIM
  {
    const long long dims[4] = {C, W, H, n_images}, strides[3] = {(long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2};
    const int box[4] = {64, W, H, box_images};
    shim_encode_tiled(out, ptr, 4, dims, strides, box);
    return true;
  }
#endif
  struct Key { const void* p; int n, h, w, c, b; bool operator==(const Key& o) const { return p == o.p && n == o.n && h == o.h && w == o.w && c == o.c && b == o.b; } };
  struct KeyHash { size_t operator()(const Key& k) const {
    return std::hash<const void*>()(k.p) ^ (size_t)k.n * 1000003u ^ (size_t)k.h * 10007u ^ (size_t)k.w * 131u ^ (size_t)k.c * 31u ^ (size_t)k.b * 7u; } };
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Key key{ptr, n_images, H, W, C, box_images};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) { g_last_error = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)W, (cuuint32_t)H, (cuuint32_t)box_images};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_error = "cuTensorMapEncodeTiled (4-d) failed with CUresult " + std::to_string((int)r); return false; }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

template <int BN, int CL>
cudaError_t launch_t(const void* A, const void* B, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s) {
  using C = Cfg<BN>;
  CUtensorMap ta, tb;
  if (!make_tmap(A, M, K, BM, &ta) || !make_tmap(B, N, K, BN / CL, &tb)) return cudaErrorInvalidValue;
  static bool configured[64] = {false};
  static int num_sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = configure_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<BN, CL>), C::kSmemBytes, dev);
    if (e != cudaSuccess) { g_last_error = "cudaFuncSetAttribute(smem) failed"; return e; }
    cudaDeviceGetAttribute(&num_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured[dev & 63] = true;
  }
  const int work = (M / BM / CL) * (N / BN) * ((CL == 1 && ep.split_k > 1) ? ep.split_k : 1);
  int units = num_sms[dev & 63] / CL;          // persistent: one CTA (or CTA pair) per SM (pair)
  if (ep.max_ctas > 0 && ep.max_ctas / CL < units) units = ep.max_ctas / CL;
  if (work < units) units = work;
  if (units < 1) units = 1;
#ifdef COLEARN_HOST_SHIM
  if (CL != 1) { g_last_error = "clusters are not modelled on the host"; return cudaErrorNotSupported; }
  void (*kern)(CUtensorMap, CUtensorMap, int, int, int, GemmEpilogue) = gemm_tcgen05_kernel<BN, CL>;
  COLEARN_LAUNCH(kern, units * CL, kThreads, C::kSmemBytes, s, ta, tb, M, N, K, ep);
  return cudaSuccess;
#else
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * CL);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (ep.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, CL>, ta, tb, M, N, K, ep);
#endif
}

// plain 1-CTA launch, with the programmatic-dependent-launch attribute when ep.pdl
inline cudaError_t launch_1cta(void (*kern)(CUtensorMap, CUtensorMap, int, int, int, GemmEpilogue), int units, int smem_bytes, cudaStream_t s,
                               const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpilogue& ep) {
#ifndef COLEARN_HOST_SHIM
  if (ep.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, ep);
  }
#endif
  COLEARN_LAUNCH(kern, units, kThreads, smem_bytes, s, ta, tb, M, N, K, ep);
  return cudaGetLastError();
}

// MN-major operands: A [K, a_cols] (AMN) or [M, K]; B [b_rows >= K, N] row-major bf16, boxes of [64 rows x 64 columns]
template <int BN, bool AMN>
cudaError_t launch_mn(const void* A, int a_cols, const void* B, int b_rows, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s) {
  using C = Cfg<BN>;
  CUtensorMap ta, tb;
  if (!(AMN ? make_tmap(A, K, a_cols, 64, &ta) : make_tmap(A, M, K, BM, &ta)) || !make_tmap(B, b_rows, N, 64, &tb))
    return cudaErrorInvalidValue;
  static bool configured[64] = {false};
  static int num_sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = configure_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<BN, 1, AMN, true>), C::kSmemBytes, dev);
    if (e != cudaSuccess) { g_last_error = "cudaFuncSetAttribute(smem, mn) failed"; return e; }
    cudaDeviceGetAttribute(&num_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured[dev & 63] = true;
  }
  const int work = (M / BM) * (N / BN) * (ep.split_k > 1 ? ep.split_k : 1);
  int units = num_sms[dev & 63];
  if (ep.max_ctas > 0 && ep.max_ctas < units) units = ep.max_ctas;
  if (work < units) units = work;
  if (units < 1) units = 1;
  return launch_1cta(gemm_tcgen05_kernel<BN, 1, AMN, true>, units, C::kSmemBytes, s, ta, tb, M, N, K, ep);
}


// =====================================================================================================================
// cta_group::2 variant: a cluster of two CTAs (= two SMs) computes one 256 x 256 tile with UMMA M = 256.
//   * each CTA stages ITS 128 rows of A and ITS 128 of the 256 B rows (32 KB / stage instead of 48 KB): the tensor
//     cores of both SMs read both halves, so per-SM shared-memory traffic drops from ~190 B/clk (TMA fill 94 + UMMA
//     reads 96, over the 128 B/clk port) to ~126 B/clk — the bound the single-CTA kernel sits on (profiles/README §3);
//   * only the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2; both CTAs' TMA loads credit the LEADER's
//     full barrier (.cta_group::2 TMA + mapa'd barrier address), tcgen05.commit multicasts the "stage free" and
//     "accumulator ready" arrivals to both CTAs, and the peer's epilogue warps release the accumulator stage with
//     remote mbarrier arrives;
//   * every CTA's 8 epilogue warps drain its own 128 TMEM lanes (its 128 output rows).
// =====================================================================================================================
struct Cfg2SM {
  static constexpr int BN = 256;
  static constexpr int kStages = 6;
  static constexpr int kTmemCols = kAccStages * BN;   // 512
  static constexpr uint32_t kStageBytesA = BM * BK * 2;         // 16 KB: my 128 rows of A
  static constexpr uint32_t kStageBytesB = (BN / 2) * BK * 2;   // 16 KB: my 128 rows of B
  static constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_2sm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        int M, int N, int K, GemmEpilogue ep) {
  using C = Cfg2SM;
  constexpr int BN = C::BN, kStages = C::kStages, kTmemCols = C::kTmemCols;
  constexpr uint32_t kStageBytesA = C::kStageBytesA, kStageBytesB = C::kStageBytesB, kStageBytes = C::kStageBytes;
  COLEARN_DYN_SMEM_UNALIGNED(uint8_t, smem_raw);
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kStageBytesA;
  SharedBarriers* bars = reinterpret_cast<SharedBarriers*>(smem + kStages * kStageBytes);

  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_units = M / (2 * BM), n_tiles = N / BN, num_kb = K / BK;
  const int total_work = m_units * n_tiles;
  const int work0 = blockIdx.x >> 1, work_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tmap_prefetch(&tmap_a);
    tmap_prefetch(&tmap_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->full[i], 2);                  // (leader only is waited on) both producers arrive
      mbar_init(&bars->empty[i], 1);                 // leader's commit, multicast to both CTAs
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&bars->tmem_full[i], 1);             // leader's commit, multicast to both CTAs
      mbar_init(&bars->tmem_empty[i], 2 * kEpilogueWarps);   // (leader's is waited on) both CTAs' epilogue warps
    }
    fence_mbarrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (ep.pdl) COLEARN_PDL_PROLOGUE();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): my A rows + my half of the B rows, bytes credited to the leader's barrier =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = work0; tile < total_work; tile += work_stride) {
        int mu, nb;
        work_to_tile(tile, m_units, n_tiles, mu, nb);
        const int m0 = (mu * 2 + (int)crank) * BM, n0 = nb * BN;
        const int nb0 = n0 + (int)crank * (BN / 2);
        if (ep.ready_flags != nullptr) {
          const uint32_t want = ep.ready_epoch_ptr ? ld_acquire_sys(ep.ready_epoch_ptr) : ep.ready_epoch;
          const int64_t c_lo = (ep.ready_elem_offset + (int64_t)nb0 * K) / ep.ready_chunk_elems;
          const int64_t c_hi = (ep.ready_elem_offset + (int64_t)(nb0 + BN / 2) * K - 1) / ep.ready_chunk_elems;
          for (int64_t c = c_lo; c <= c_hi; ++c)
            spin_wait_ge(ep.ready_flags + c, want, 64, "gemm_tcgen05: ready flag of a broadcast weight chunk");
          if (ep.bias != nullptr) {
            const int64_t b_lo = (ep.ready_elem_offset + (int64_t)N * K + n0) / ep.ready_chunk_elems;
            const int64_t b_hi = (ep.ready_elem_offset + (int64_t)N * K + n0 + BN - 1) / ep.ready_chunk_elems;
            for (int64_t c = b_lo; c <= b_hi; ++c)
              spin_wait_ge(ep.ready_flags + c, want, 64, "gemm_tcgen05: ready flag of a broadcast weight chunk");
          }
          fence_proxy_async();
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          const uint32_t lead_full = mapa_shared(smem_u32(&bars->full[stage]), 0u);
          if (leader) mbar_arrive_expect_tx(&bars->full[stage], 2 * kStageBytes);   // bytes of BOTH CTAs
          else mbar_arrive_remote(lead_full);
          tma_load_2d_2sm(smem_a + stage * kStageBytesA, &tmap_a, kb * BK, m0, lead_full);
          tma_load_2d_2sm(smem_b + stage * kStageBytesB, &tmap_b, kb * BK, nb0, lead_full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only, one lane =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = work0; tile < total_work; tile += work_stride, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * kStageBytesA));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * kStageBytesB));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&bars->empty[stage], (uint16_t)0x3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&bars->tmem_full[as], (uint16_t)0x3);
      }
    }
  } else {
    // ===== epilogue warps 2..9 (both CTAs): my 128 rows of the 256-row tile =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int local = 0;
    for (int tile = work0; tile < total_work; tile += work_stride, ++local) {
      int mu, nb;
        work_to_tile(tile, m_units, n_tiles, mu, nb);
        const int m0 = (mu * 2 + (int)crank) * BM, n0 = nb * BN;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), v);
        tmem_ld_wait();
        epilogue_chunk(v, ep, row, n0 + c0, lane, q, m0, M, N);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&bars->tmem_empty[as]);
        else mbar_arrive_remote(mapa_shared(smem_u32(&bars->tmem_empty[as]), 0u));
        if (ep.produced != nullptr) {      // fused wgrad -> FedAvg reduce, see the 1-CTA kernel
          __threadfence_system();
          produced_block(ep.produced, ep.produced_elem_offset, m0 + q * 32, 32, N, n0 + half * (BN / 2), BN / 2);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

cudaError_t launch_2sm(const void* A, const void* B, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s) {
  using C = Cfg2SM;
  CUtensorMap ta, tb;
  if (!make_tmap(A, M, K, BM, &ta) || !make_tmap(B, N, K, C::BN / 2, &tb)) return cudaErrorInvalidValue;
  static bool configured[64] = {false};
  static int num_sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = configure_smem(reinterpret_cast<const void*>(gemm_tcgen05_2sm_kernel), C::kSmemBytes, dev);
    if (e != cudaSuccess) { g_last_error = "cudaFuncSetAttribute(smem, 2sm) failed"; return e; }
    cudaDeviceGetAttribute(&num_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured[dev & 63] = true;
  }
  const int work = (M / (2 * BM)) * (N / C::BN);
  int units = num_sms[dev & 63] / 2;
  if (ep.max_ctas > 0 && ep.max_ctas / 2 < units) units = ep.max_ctas / 2;
  if (work < units) units = work;
  if (units < 1) units = 1;
#ifdef COLEARN_HOST_SHIM
  (void)units;
  g_last_error = "cta_group::2 is not modelled on the host";
  return cudaErrorNotSupported;
#else
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * 2);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ep.pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_tcgen05_2sm_kernel, ta, tb, M, N, K, ep);
#endif
}

}  // namespace

const char* gemm_tcgen05_last_error() { return g_last_error.c_str(); }

template <int BN, bool WGRAD, bool BMN = WGRAD>
cudaError_t launch_conv_t(const void* act, int n_images, int H, int W, const void* other, int other_rows, int other_cols,
                          int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s) {
  using C = Cfg<BN>;
  CUtensorMap ta, tb;
  const int HW = H * W;
  if (WGRAD) {   // A = dz [K pixels, a_cols] MN-major boxes, B = activation boxes of 64 pixels
    if (!make_tmap(other, other_rows, other_cols, 64, &ta) || !make_tmap_4d(act, n_images, H, W, ep.conv.C, 64 / HW, &tb))
      return cudaErrorInvalidValue;
  } else {       // A = activation boxes of 128 pixels, B = K-major weights [other_rows, other_cols] (MN-major: 64-row boxes)
    if (!make_tmap_4d(act, n_images, H, W, ep.conv.C, 128 / HW, &ta) || !make_tmap(other, other_rows, other_cols, BMN ? 64 : BN, &tb))
      return cudaErrorInvalidValue;
  }
  static bool configured[64] = {false};
  static int num_sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = configure_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<BN, 1, WGRAD, BMN>), C::kSmemBytes, dev);
    if (e != cudaSuccess) { g_last_error = "cudaFuncSetAttribute(smem, conv) failed"; return e; }
    cudaDeviceGetAttribute(&num_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured[dev & 63] = true;
  }
  const int work = (M / BM) * (N / BN) * (ep.split_k > 1 ? ep.split_k : 1);
  int units = num_sms[dev & 63];
  if (ep.max_ctas > 0 && ep.max_ctas < units) units = ep.max_ctas;
  if (work < units) units = work;
  if (units < 1) units = 1;
  return launch_1cta(gemm_tcgen05_kernel<BN, 1, WGRAD, BMN>, units, C::kSmemBytes, s, ta, tb, M, N, K, ep);
}

cudaError_t launch_gemm_tcgen05_conv(const void* act, int n_images, int H, int W, const void* other, int other_rows, int other_cols,
                                     int M, int N, int K, const GemmEpilogue& ep_in, cudaStream_t s) {
  GemmEpilogue ep = ep_in;
  convops::ConvAddr& g = ep.conv;
  const int HW = H * W;
  // 1x1 images (the last ResNet stage at 32x32 input): with KH = KW = 2*pad + 1 only the centre tap overlaps the image, all
  // other boxes are pure padding = zeros — walk that single tap (the caller still describes the full convolution)
  const bool centre_only = g.mode == 1 && HW == 1 && g.KH == 2 * g.pad + 1 && g.KW == 2 * g.pad + 1 && g.KH * g.KW > 1 &&
                           K == g.KH * g.KW * g.C && ep.split_k <= 1;
  g.tap_lo = centre_only ? g.pad * g.KW + g.pad : 0;
  g.tap_cnt = centre_only ? 1 : g.KH * g.KW;
  if ((g.mode != 1 && g.mode != 2) || g.C <= 0 || (g.C % 64) || g.KH <= 0 || g.KW <= 0 || g.pad < 0 || g.HW != HW ||
      !(HW == 1 || HW == 4 || HW == 16 || HW == 64) || W > 256 || H > 256 || g.n_images != n_images) {
    g_last_error = "implicit conv GEMM: mode 1|2, C%64==0, H*W in {1,4,16,64}, consistent geometry";
    return cudaErrorInvalidValue;
  }
  // N % 64 for the A-operand modes: a 64-channel layer's dgrad has only 64 output columns (UMMA 128x64x16 tile)
  if (M <= 0 || N <= 0 || K <= 0 || (M % BM) || (N % (g.mode == 1 ? 64 : 128)) || (K % BK)) {
    g_last_error = "shape must satisfy M%128==0, N%128==0 (N%64==0 with a conv A operand), K%64==0";
    return cudaErrorInvalidValue;
  }
  if ((((uintptr_t)act) | ((uintptr_t)other)) & 15) { g_last_error = "operands must be 16-byte aligned"; return cudaErrorInvalidValue; }
  if (ep.cluster != 0 || ep.ready_flags != nullptr) { g_last_error = "implicit conv GEMM has no cluster / ready-flag variant"; return cudaErrorInvalidValue; }
  if (g.mode == 1) {
    // M = pixels (whole images per 128-row tile), K = taps * C
    if (M != n_images * HW || K != g.KH * g.KW * g.C) { g_last_error = "implicit conv GEMM (A): M = N*H*W, K = KH*KW*C"; return cudaErrorInvalidValue; }
    if (centre_only) K = g.C;
    if (g.b_mn && !g.flip) { g_last_error = "implicit conv GEMM: the MN-major weight operand is the dgrad's"; return cudaErrorInvalidValue; }
    const int need_n = (g.KH * g.KW - 1) * g.b_rows_per_tap + N;   // dgrad: last tap's block of N weight rows / columns
    if (g.flip ? (g.b_rows_per_tap <= 0 || (g.b_mn ? (other_cols < need_n || other_rows < g.C) : (other_rows < need_n || other_cols < g.C)))
               : (other_rows < N || other_cols < K)) {
      g_last_error = "implicit conv GEMM (A): weight matrix too small";
      return cudaErrorInvalidValue;
    }
  } else {
    // K = pixels, N = padded taps * C
    if (K != n_images * HW || N < g.KH * g.KW * g.C || other_rows != K || other_cols > M || (other_cols % 8)) {
      g_last_error = "implicit conv GEMM (B): K = N*H*W, N >= KH*KW*C, dz [K, a_cols <= M]";
      return cudaErrorInvalidValue;
    }
  }
  if (ep.split_k > 1) {
    if (ep.split_out == nullptr || K / BK < ep.split_k || (((uintptr_t)ep.split_out) & 15)) {
      g_last_error = "split_k needs a 16-byte aligned split_out and K/64 >= split_k";
      return cudaErrorInvalidValue;
    }
    if (ep.bias || ep.relu || ep.relu_mask || ep.out_bf16 || ep.out_f32 || ep.out_bf16_t || ep.sgd_master || ep.colsum || ep.addend) {
      g_last_error = "split_k stores raw partials: no other epilogue option may be set";
      return cudaErrorInvalidValue;
    }
  }
  if (ep.tile_n == 256 && (N % 256)) { g_last_error = "tile_n=256 needs N%256==0"; return cudaErrorInvalidValue; }
  const bool wide = (N % 256 == 0) && ep.tile_n != 128 &&
                    (ep.tile_n == 256 || ep.split_k > 1 || (int64_t)(M / BM) * (N / 256) >= 120);
  if (g.mode == 2) {
    if (wide) return launch_conv_t<256, true>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
    return launch_conv_t<128, true>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
  }
  if (g.b_mn) {
    if (wide) return launch_conv_t<256, false, true>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
    if (N % 128) return launch_conv_t<64, false, true>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
    return launch_conv_t<128, false, true>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
  }
  if (wide) return launch_conv_t<256, false>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
  if (N % 128) return launch_conv_t<64, false>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
  return launch_conv_t<128, false>(act, n_images, H, W, other, other_rows, other_cols, M, N, K, ep, s);
}

// a_mn = 1: C[M, N] = A^T B for A [K, a_cols] (a_cols <= M; the missing columns count as zeros) and B [K, N];
// a_mn = 0: C[M, N] = A B   for A [M, K] and B [b_rows >= K, N] (only the first K rows are read).  All row-major bf16.
cudaError_t launch_gemm_tcgen05_mn(const void* A, int a_mn, int a_cols, const void* B, int b_rows, int M, int N, int K,
                                   const GemmEpilogue& ep, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || (M % BM) || (N % 128) || (K % BK) || b_rows < K) {
    g_last_error = "mn-major shape must satisfy M%128==0, N%128==0, K%64==0, b_rows >= K";
    return cudaErrorInvalidValue;
  }
  if (a_mn && (a_cols <= 0 || a_cols > M || (a_cols % 8))) {
    g_last_error = "mn-major A needs 0 < a_cols <= M, a_cols%8==0";
    return cudaErrorInvalidValue;
  }
  if ((((uintptr_t)A) | ((uintptr_t)B)) & 15) { g_last_error = "operands must be 16-byte aligned"; return cudaErrorInvalidValue; }
  if (ep.cluster != 0 || ep.ready_flags != nullptr) { g_last_error = "mn-major mode has no cluster / ready-flag variant"; return cudaErrorInvalidValue; }
  if (ep.split_k > 1) {
    if (ep.split_out == nullptr || K / BK < ep.split_k || (((uintptr_t)ep.split_out) & 15)) {
      g_last_error = "split_k needs a 16-byte aligned split_out and K/64 >= split_k";
      return cudaErrorInvalidValue;
    }
    if (ep.bias || ep.relu || ep.relu_mask || ep.out_bf16 || ep.out_f32 || ep.out_bf16_t || ep.sgd_master || ep.colsum || ep.addend) {
      g_last_error = "split_k stores raw partials: no other epilogue option may be set";
      return cudaErrorInvalidValue;
    }
  }
  if (ep.tile_n == 256 && (N % 256)) { g_last_error = "tile_n=256 needs N%256==0"; return cudaErrorInvalidValue; }
  const bool wide = (N % 256 == 0) && ep.tile_n != 128 &&
                    (ep.tile_n == 256 || ep.split_k > 1 || (int64_t)(M / BM) * (N / 256) >= 120);
  if (a_mn) {
    if (wide) return launch_mn<256, true>(A, a_cols, B, b_rows, M, N, K, ep, s);
    return launch_mn<128, true>(A, a_cols, B, b_rows, M, N, K, ep, s);
  }
  if (wide) return launch_mn<256, false>(A, a_cols, B, b_rows, M, N, K, ep, s);
  return launch_mn<128, false>(A, a_cols, B, b_rows, M, N, K, ep, s);
}

cudaError_t launch_gemm_tcgen05(const void* A, const void* B, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || (M % BM) || (N % 128) || (K % BK)) {
    g_last_error = "shape must satisfy M%128==0, N%128==0, K%64==0";
    return cudaErrorInvalidValue;
  }
  if ((((uintptr_t)A) | ((uintptr_t)B)) & 15) { g_last_error = "operands must be 16-byte aligned"; return cudaErrorInvalidValue; }
  if (ep.produced != nullptr && (ep.sgd_master == nullptr || ep.split_k > 1)) {
    g_last_error = "produced reports come from the fused-SGD epilogue (needs sgd_master, no split_k)";
    return cudaErrorInvalidValue;
  }
  if (ep.split_k > 1) {
    // split-K: raw fp32 partials only, 1-CTA kernel; 128x256 tiles whenever N allows (half the B smem traffic per FLOP)
    if (ep.split_out == nullptr) { g_last_error = "split_k needs split_out"; return cudaErrorInvalidValue; }
    if (K / BK < ep.split_k) { g_last_error = "split_k must not exceed K/64"; return cudaErrorInvalidValue; }
    if (ep.cluster != 0 || ep.ready_flags != nullptr || ep.bias || ep.relu || ep.relu_mask || ep.out_bf16 || ep.out_f32 ||
        ep.out_bf16_t || ep.sgd_master || ep.colsum || ep.addend) {
      g_last_error = "split_k stores raw partials: no other epilogue / cluster / ready-flag option may be set";
      return cudaErrorInvalidValue;
    }
    if (((uintptr_t)ep.split_out) & 15) { g_last_error = "split_out must be 16-byte aligned"; return cudaErrorInvalidValue; }
    if (ep.tile_n == 256 && (N % 256)) { g_last_error = "tile_n=256 needs N%256==0"; return cudaErrorInvalidValue; }
    if (ep.tile_n == 256 || (ep.tile_n == 0 && N % 256 == 0)) return launch_t<256, 1>(A, B, M, N, K, ep, s);
    return launch_t<128, 1>(A, B, M, N, K, ep, s);
  }
  // BN=256 when it divides N and leaves enough tiles to fill the machine; BN=128 otherwise
  // cta_group::2 (two SMs per 256x256 tile): forced with cluster == 3, automatic when the shape allows it and there
  // are enough tile pairs to fill the machine (measured: 1 485 vs 1 263 TFLOP/s at 4096^3, equal at 1024x4096x4096)
#ifdef COLEARN_HOST_SHIM
  const bool can_2sm = false;                      // cta_group::2 / clusters are GPU-only
#else
  const bool can_2sm = (M % 256 == 0) && (N % 256 == 0);
#endif
  if (ep.cluster == 3 && !can_2sm) { g_last_error = "cta_group::2 needs M%256==0 and N%256==0"; return cudaErrorInvalidValue; }
  if (ep.cluster == 3 || (ep.cluster == 0 && ep.tile_n == 0 && can_2sm && (int64_t)(M / 256) * (N / 256) >= 60))
    return launch_2sm(A, B, M, N, K, ep, s);
  const bool wide = (N % 256 == 0) && ((int64_t)(M / BM) * (N / 256) >= 120) && ep.tile_n != 128;
  if (wide || ep.tile_n == 256) {
    if (N % 256) { g_last_error = "tile_n=256 needs N%256==0"; return cudaErrorInvalidValue; }
    // cluster of 2 with TMA multicast of the shared B tile (ep.cluster == 2).  Measured (profiles/README.md §3): it cuts
    // L2->SM traffic by a third but not the per-SM smem traffic (TMA writes 94 B/clk + UMMA reads 96 B/clk against a
    // 128 B/clk port), so it is not faster than independent CTAs; it stays opt-in until the cta_group::2 MMA lands.
    const bool pair = (M % (2 * BM) == 0) && ep.cluster == 2;
    if (ep.cluster == 2 && !pair) { g_last_error = "cluster=2 needs M%256==0"; return cudaErrorInvalidValue; }
    if (pair) return launch_t<256, 2>(A, B, M, N, K, ep, s);
    return launch_t<256, 1>(A, B, M, N, K, ep, s);
  }
  return launch_t<128, 1>(A, B, M, N, K, ep, s);
}

COLEARN_DEFINE_SPIN_LIMIT_SETTER(set_spin_limit_gemm)

}  // namespace colearn

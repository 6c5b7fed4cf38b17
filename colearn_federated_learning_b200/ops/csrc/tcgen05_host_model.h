// Functional CPU model of the Blackwell machinery gemm_tcgen05.cu is written against — mbarrier (arrival + transaction
// counts, phase parity), TMA tiled loads (2-D / 4-D boxes, SWIZZLE_128B, zero fill outside the tensor), tcgen05.mma
// reading shared memory through UMMA descriptors (K-major and MN-major canonical layouts, the 128-byte XOR swizzle),
// TMEM and tcgen05.ld — so that the kernel SOURCE runs in the CPU test-suite through host_shim.h (one OS thread per CUDA
// thread).  It is a model of the programming interface as documented (CUTLASS cute/arch/mma_sm100_desc.hpp, the
// canonical-layout comments of cute/atom/mma_traits_sm100.hpp), not of the hardware: what it proves is that the
// kernel's pipelines, barrier phases, work decomposition, split-K, descriptor arithmetic, TMA coordinates and
// epilogues are consistent with that interface and produce the right numbers — for the K-major path, which IS
// validated on a B200, the model and the silicon agree, which calibrates it for the MN-major / 4-D-box paths that
// have not run on a GPU yet.  asynchronous units (TMA, tensor core) execute synchronously at issue.
//
// Included by gemm_tcgen05.cu INSTEAD of its PTX wrappers when COLEARN_HOST_SHIM is defined (inside namespace colearn).
#pragma once

#ifndef __grid_constant__
#define __grid_constant__
#endif

// ---- shared-window addresses ---------------------------------------------------------------------------------------
// smem_u32(p) = byte offset of p from the 1024-aligned base of the CTA's dynamic shared memory (+ 1024, so that 0 is
// never a valid address).  The swizzle is a function of these address bits, as on the device.
inline uint8_t* shim_smem_base() {
  uintptr_t p = reinterpret_cast<uintptr_t>(::colearn_shim::dyn_smem());
  return reinterpret_cast<uint8_t*>((p + 1023) & ~(uintptr_t)1023);
}
inline uint32_t smem_u32(const void* p) { return (uint32_t)(reinterpret_cast<const uint8_t*>(p) - shim_smem_base()) + 1024u; }
inline uint8_t* shim_smem_ptr(uint32_t saddr) { return shim_smem_base() + (saddr - 1024u); }

// ---- mbarrier ------------------------------------------------------------------------------------------------------
// one 64-bit word: [0] phase parity | [1,17) pending arrivals | [17,33) arrivals per phase | [33,64) pending tx bytes (signed)
struct ShimMbar {
  static uint64_t pack(uint64_t phase, uint64_t pending, uint64_t init, int64_t tx) {
    return (phase & 1) | ((pending & 0xFFFF) << 1) | ((init & 0xFFFF) << 17) | ((uint64_t)(tx & 0x7FFFFFFF) << 33);
  }
  static void unpack(uint64_t w, uint64_t& phase, uint64_t& pending, uint64_t& init, int64_t& tx) {
    phase = w & 1;
    pending = (w >> 1) & 0xFFFF;
    init = (w >> 17) & 0xFFFF;
    tx = (int64_t)((w >> 33) & 0x7FFFFFFF);
    if (tx & 0x40000000) tx -= 0x80000000ll;     // sign-extend 31 bits
  }
  // arrivals -= d_arrive, tx += d_tx; the phase completes (flips, arrivals re-armed) when both reach zero
  static void update(uint64_t* bar, int d_arrive, int64_t d_tx) {
    uint64_t old = __atomic_load_n(bar, __ATOMIC_ACQUIRE), want;
    do {
      uint64_t phase, pending, init;
      int64_t tx;
      unpack(old, phase, pending, init, tx);
      if ((int)pending < d_arrive) abort();      // more arrivals than the barrier was initialised for: a protocol bug
      pending -= (uint64_t)d_arrive;
      tx += d_tx;
      if (pending == 0 && tx == 0) {
        phase ^= 1;
        pending = init;
      }
      want = pack(phase, pending, init, tx);
    } while (!__atomic_compare_exchange_n(bar, &old, want, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE));
  }
};
inline void mbar_init(uint64_t* bar, uint32_t count) { __atomic_store_n(bar, ShimMbar::pack(0, count, count, 0), __ATOMIC_RELEASE); }
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { ShimMbar::update(bar, 1, (int64_t)bytes); }
inline void mbar_arrive(uint64_t* bar) { ShimMbar::update(bar, 1, 0); }
inline uint32_t mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  const uint64_t w = __atomic_load_n(reinterpret_cast<uint64_t*>(shim_smem_ptr(bar_addr)), __ATOMIC_ACQUIRE);
  return ((w & 1) != (parity & 1)) ? 1u : 0u;    // the phase with this parity has completed
}
constexpr unsigned long long kMbarTimeoutNs = 20ull * 1000 * 1000 * 1000;
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const auto t0 = std::chrono::steady_clock::now();
  while (!mbar_try_wait(addr, parity)) {
    std::this_thread::yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::nanoseconds(kMbarTimeoutNs)) abort();   // the device traps here
  }
}

// ---- tensor maps + TMA -----------------------------------------------------------------------------------------------
// What cuTensorMapEncodeTiled would encode, kept in the opaque 128 bytes of a CUtensorMap (bf16, SWIZZLE_128B, zero fill)
struct ShimTensorMap {
  const __nv_bfloat16* ptr;
  int rank;
  long long dims[4];            // elements, innermost first
  long long strides[4];         // bytes; strides[0] = 2
  int box[4];
};
static_assert(sizeof(ShimTensorMap) <= sizeof(CUtensorMap), "mock tensor map must fit the opaque descriptor");
inline void shim_encode_tiled(CUtensorMap* out, const void* ptr, int rank, const long long* dims, const long long* strides_bytes, const int* box) {
  ShimTensorMap m;
  memset(&m, 0, sizeof(m));
  m.ptr = static_cast<const __nv_bfloat16*>(ptr);
  m.rank = rank;
  for (int i = 0; i < rank; ++i) {
    m.dims[i] = dims[i];
    m.strides[i] = i == 0 ? 2 : strides_bytes[i - 1];
    m.box[i] = box[i];
  }
  if (box[0] != 64) abort();    // SWIZZLE_128B: one box line = 64 bf16
  memset(out, 0, sizeof(*out));
  memcpy(out, &m, sizeof(m));
}
// box lines (all outer coordinates flattened, innermost outer dim fastest) land as 128-byte lines; line r, 16-byte chunk c
// is stored at chunk position c ^ (r % 8) (the address bits [4,7) ^ [7,10) of a 1024-aligned destination)
inline void shim_tma_load(void* smem_dst, const CUtensorMap* tmap, const int* coord, uint64_t* bar) {
  ShimTensorMap m;
  memcpy(&m, tmap, sizeof(m));
  uint8_t* dst = static_cast<uint8_t*>(smem_dst);
  const uint32_t dst_addr = smem_u32(dst);
  if ((dst_addr - 1024u) % 1024u) abort();       // swizzle atoms need 1024-byte aligned boxes
  long long lines = 1;
  for (int i = 1; i < m.rank; ++i) lines *= m.box[i];
  for (long long r = 0; r < lines; ++r) {
    long long rem = r, off_bytes = 0;
    bool inside = true;
    for (int i = 1; i < m.rank; ++i) {
      const long long idx = coord[i] + rem % m.box[i];
      rem /= m.box[i];
      if (idx < 0 || idx >= m.dims[i]) inside = false;
      off_bytes += idx * m.strides[i];
    }
    for (int c = 0; c < 64; ++c) {
      const long long ci = (long long)coord[0] + c;
      __nv_bfloat16 v = __float2bfloat16(0.f);
      if (inside && ci >= 0 && ci < m.dims[0]) v = *reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const uint8_t*>(m.ptr) + off_bytes + ci * 2);
      const uint32_t lin = (uint32_t)r * 128u + (uint32_t)c * 2u;
      const uint32_t addr = dst_addr - 1024u + lin;
      const uint32_t sw = addr ^ (((addr >> 7) & 7u) << 4);
      *reinterpret_cast<__nv_bfloat16*>(shim_smem_base() + sw) = v;
    }
  }
  ShimMbar::update(bar, 0, -(long long)lines * 128);
}
inline void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  const int coord[2] = {c0, c1};
  shim_tma_load(smem_dst, tmap, coord, bar);
}
inline void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  const int coord[4] = {c0, c1, c2, c3};
  shim_tma_load(smem_dst, tmap, coord, bar);
}
inline void tmap_prefetch(const CUtensorMap*) {}
inline void fence_mbarrier_init() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void fence_proxy_async() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- TMEM + tcgen05.mma / commit / ld --------------------------------------------------------------------------------
// 128 lanes x 512 columns of fp32 per CTA (a per-block heap array)
inline float* shim_tmem() { return ::colearn_shim::t_block->scratch(128 * 512); }
inline void tmem_alloc(uint32_t* smem_dst, uint32_t /*ncols*/) {
  if (::colearn_shim::t_linear % 32 == 0) *smem_dst = 0u;       // base address: lane 0, column 0
  ::colearn_shim::syncwarp();
}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void tc_fence_after() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void umma_commit(uint64_t* bar) { ShimMbar::update(bar, 1, 0); }   // the MMAs ran synchronously: they are complete

inline float shim_ld_bf16(uint32_t addr) {   // addr: un-swizzled byte offset in the shared window
  const uint32_t sw = addr ^ (((addr >> 7) & 7u) << 4);
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(shim_smem_base() + sw));
}
// one tcgen05.mma.cta_group::1.kind::f16: D[M, N] (+)= A[M, 16] * B[N, 16]^T, operands addressed through their descriptors
inline void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != 1 || ((idesc >> 10) & 7) != 1) abort();   // F32 accumulate, BF16 x BF16
  auto field = [](uint64_t d, int lo, int bits) { return (uint32_t)((d >> lo) & ((1ull << bits) - 1)); };
  auto operand = [&](uint64_t d, bool mn_major, int rows, float* out /* [rows][16] */) {
    if (field(d, 61, 3) != 2 || field(d, 46, 2) != 1) abort();                     // SWIZZLE_128B, descriptor version 1
    const uint32_t start = (field(d, 0, 14) << 4) - 1024u, lbo = field(d, 16, 14) << 4, sbo = field(d, 32, 14) << 4;
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < 16; ++k) {
        uint32_t addr;
        if (mn_major)   // ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 64 MN elements per line, 8 k-lines per atom
          addr = start + (uint32_t)(r % 64) * 2u + (uint32_t)(r / 64) * lbo + (uint32_t)(k % 8) * 128u + (uint32_t)(k / 8) * sbo;
        else            // K-major: row r = 128-byte line, 8-row groups SBO apart; the start address carries the k offset
          addr = start + (uint32_t)(r % 8) * 128u + (uint32_t)(r / 8) * sbo + (uint32_t)k * 2u;
        out[r * 16 + k] = shim_ld_bf16(addr);
      }
  };
  std::vector<float> A((size_t)M * 16), B((size_t)N * 16);
  operand(desc_a, a_mn, M, A.data());
  operand(desc_b, b_mn, N, B.data());
  float* tm = shim_tmem();
  const uint32_t col0 = tmem_d & 0xFFFFu, lane0 = tmem_d >> 16;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      float acc = accumulate ? tm[(size_t)(lane0 + i) * 512 + col0 + j] : 0.f;
      const float* a = &A[(size_t)i * 16];
      const float* b = &B[(size_t)j * 16];
      for (int k = 0; k < 16; ++k) acc += a[k] * b[k];
      tm[(size_t)(lane0 + i) * 512 + col0 + j] = acc;
    }
}
inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  const uint32_t lane = (taddr >> 16) + ::colearn_shim::t_linear % 32, col = taddr & 0xFFFFu;
  const float* tm = shim_tmem();
  for (int j = 0; j < 32; ++j) memcpy(&v[j], &tm[(size_t)lane * 512 + col + j], 4);
}
inline void tmem_ld_wait() {}
inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

// ---- cluster / cta_group::2 forms: not modelled (the tests drive the 1-CTA kernels) --------------------------------------
inline uint32_t cluster_ctarank() { return 0u; }
inline void cluster_sync_all() {}
inline void tma_load_2d_mcast(void*, const CUtensorMap*, int, int, uint64_t*, uint16_t) { abort(); }
inline void umma_commit_mcast(uint64_t*, uint16_t) { abort(); }
inline uint32_t mapa_shared(uint32_t a, uint32_t) { return a; }
inline void mbar_arrive_remote(uint32_t) { abort(); }
inline void tma_load_2d_2sm(void*, const CUtensorMap*, int, int, uint32_t) { abort(); }
inline void tmem_alloc_2sm(uint32_t*, uint32_t) { abort(); }
inline void tmem_dealloc_2sm(uint32_t, uint32_t) {}
inline void umma_commit_2sm(uint64_t*, uint16_t) { abort(); }
inline void umma_bf16_2sm(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) { abort(); }

// Memory-bound kernels: SGD step, FedAvg on flat arenas, fused loss fwd+bwd, eval, argmax,
// min-max scaling, keyed permutation, dtype conversion.  All are single-pass, 128-bit vectorised
// where alignment allows, and sized as a persistent grid (148 SMs x 8 CTAs) with grid-stride loops.
//
// Parity map: K9 sigmoid+BCE (client_federated.py:65,79), K10 softmax-xent (BASELINE
// 2-logit configs), K11 SSE (cf.py:112,158), K12 SGD (fc.py:355; cf.py:206-207), K13 MinMax
// (datasets.py:31-32), K14 shuffled sampling (cf.py:203), K15 eval (cf.py:233-253), K16 argmax
// (fc.py:252-256), K3 FedAvg (fc.py:373,568).
#include "colearn_kernels.h"

#include <cuda_bf16.h>
#include <math.h>

namespace colearn {
namespace {

constexpr int kThreads = 256;
inline int grid_for(int64_t n, int per_thread = 4) {
  int64_t b = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  const int64_t cap = 148 * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* scratch /*[32]*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? scratch[lane] : 0.f;
    v = warp_sum(v);
  }
  __syncthreads();
  return v;  // valid in warp 0
}

// ---- SGD ---------------------------------------------------------------------------------------
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float lr, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g)) & 15) == 0;
  if (aligned) {
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t j = i; j < n4; j += stride) {
      float4 a = p4[j];
      const float4 b = __ldg(g4 + j);
      a.x = fmaf(-lr, b.x, a.x); a.y = fmaf(-lr, b.y, a.y);
      a.z = fmaf(-lr, b.z, a.z); a.w = fmaf(-lr, b.w, a.w);
      p4[j] = a;
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) p[j] = fmaf(-lr, g[j], p[j]);
  } else {
    for (; i < n; i += stride) p[i] = fmaf(-lr, g[i], p[i]);
  }
}

// p *= scale in place: a rank pre-scales its trained arena by its FedAvg weight n_k / sum n (0 for a rank that was not
// selected) so that the in-switch multimem.ld_reduce of the two-shot kernel — which sums every member of the multicast group
// with equal weight — yields the weighted / subset average
__global__ void scale_inplace_kernel(float* __restrict__ p, float scale, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)p) & 15) == 0) {
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int64_t j = i; j < n4; j += stride) {
      float4 a = p4[j];
      a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
      p4[j] = a;
    }
    for (int64_t j = (n4 << 2) + i; j < n; j += stride) p[j] *= scale;
  } else {
    for (int64_t j = i; j < n; j += stride) p[j] *= scale;
  }
}

__global__ void sgd_step_bf16grad_kernel(float* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                                         float lr, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    p[i] = fmaf(-lr, __bfloat162float(g[i]), p[i]);
}

// ---- FedAvg on flat arenas ------------------------------------------------------------------------
template <bool APPLY>
__global__ void fedavg_kernel(float* __restrict__ theta, const float* __restrict__ slots,
                              int64_t slot_stride, const float* __restrict__ weights, int k,
                              float server_lr, int64_t n) {
  // no cap on the number of workers (the coordinator's upper bound is 100, selection 'all' applies none): the first
  // kSw weights are staged in shared memory, any beyond that are read through the read-only cache
  constexpr int kSw = 256;
  __shared__ float sw[kSw];
  for (int c = threadIdx.x; c < k && c < kSw; c += blockDim.x) sw[c] = weights[c];
  __syncthreads();
  auto wt = [&](int c) -> float { return c < kSw ? sw[c] : __ldg(weights + c); };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool aligned = (((uintptr_t)theta | (uintptr_t)slots) & 15) == 0 && (slot_stride & 3) == 0;
  if (aligned) {
    const int64_t n4 = n >> 2;
    for (int64_t j = i0; j < n4; j += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < k; ++c) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(slots + c * slot_stride) + j);
        const float w = wt(c);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      float4* t4 = reinterpret_cast<float4*>(theta) + j;
      if (APPLY) {
        float4 t = *t4;
        t.x = fmaf(server_lr, acc.x - t.x, t.x); t.y = fmaf(server_lr, acc.y - t.y, t.y);
        t.z = fmaf(server_lr, acc.z - t.z, t.z); t.w = fmaf(server_lr, acc.w - t.w, t.w);
        *t4 = t;
      } else {
        *t4 = acc;
      }
    }
    for (int64_t j = (n4 << 2) + i0; j < n; j += stride) {
      float acc = 0.f;
      for (int c = 0; c < k; ++c) acc = fmaf(wt(c), slots[c * slot_stride + j], acc);
      theta[j] = APPLY ? fmaf(server_lr, acc - theta[j], theta[j]) : acc;
    }
  } else {
    for (int64_t j = i0; j < n; j += stride) {
      float acc = 0.f;
      for (int c = 0; c < k; ++c) acc = fmaf(wt(c), slots[c * slot_stride + j], acc);
      theta[j] = APPLY ? fmaf(server_lr, acc - theta[j], theta[j]) : acc;
    }
  }
}

// ---- losses ---------------------------------------------------------------------------------------
__global__ void sigmoid_bce_kernel(const float* __restrict__ z, const float* __restrict__ y,
                                   float* __restrict__ dz, float* __restrict__ loss_out, int64_t n) {
  __shared__ float scratch[32];
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float p = 1.f / (1.f + __expf(-z[i]));
    const float t = y[i];
    const float lp = fmaxf(__logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
    acc -= t * lp + (1.f - t) * l1p;
    if (dz) dz[i] = (p - t) * inv_n;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(loss_out, acc * inv_n);
}

__global__ void sse_kernel(const float* __restrict__ out, const float* __restrict__ y,
                           float* __restrict__ dz, float* __restrict__ loss_out, int64_t n, float scale) {
  __shared__ float scratch[32];
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = out[i] - y[i];
    acc = fmaf(d, d, acc);
    if (dz) dz[i] = 2.f * d * scale;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(loss_out, acc * scale);
}

// One warp per row; cols <= 1024.  Reads logits once, writes dlogits (fp32 and/or bf16).
template <bool BF16_IN>
__global__ void softmax_xent_kernel(const void* __restrict__ logits_, const int64_t* __restrict__ labels,
                                    float* __restrict__ dl, __nv_bfloat16* __restrict__ dl_bf16,
                                    float* __restrict__ loss_out, int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_rows = 1.f / (float)rows;
  float loss_acc = 0.f;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    auto load = [&](int c) -> float {
      if (BF16_IN) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(logits_)[(size_t)r * cols + c]);
      return reinterpret_cast<const float*>(logits_)[(size_t)r * cols + c];
    };
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, load(c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += __expf(load(c) - m);
    s = warp_sum(s);
    const int label = (int)labels[r];
    const float inv_s = 1.f / s;
    for (int c = lane; c < cols; c += 32) {
      const float z = load(c);
      const float g = (__expf(z - m) * inv_s - (c == label ? 1.f : 0.f)) * inv_rows;
      if (dl) dl[(size_t)r * cols + c] = g;
      if (dl_bf16) dl_bf16[(size_t)r * cols + c] = __float2bfloat16(g);
      if (c == label) loss_acc += (__logf(s) + m) - z;
    }
  }
  loss_acc = warp_sum(loss_acc);
  if (lane == 0 && loss_acc != 0.f) atomicAdd(loss_out, loss_acc * inv_rows);
}

// Loss head of the GEMM-shaped trainers in ONE launch and ONE block: softmax cross-entropy of the valid rows / columns of a PADDED
// fp32 (or bf16) logits matrix, dL/dlogits written straight into the padded bf16 operand of the backward GEMMs (and / or an fp32
// matrix), the head's bias gradient (column sums of dL/dlogits) and the mean loss.  Replaces slice-copy + loss + cast-copy +
// at::sum + copy (three ATen launches, one of them an 18 us single-block reduction) around the old kernel.  One block, fixed
// summation order: bit-reproducible, no memset, no atomics.  cols <= kHeadMaxCols.
constexpr int kHeadMaxCols = 128;
template <bool BF16_IN>
__global__ void __launch_bounds__(1024)
softmax_xent_head_kernel(const void* __restrict__ logits_, int ld_in, const int64_t* __restrict__ labels, float* __restrict__ dl, int ld_dl,
                         __nv_bfloat16* __restrict__ dl_bf16, int ld_bf16, float* __restrict__ db, float* __restrict__ loss_out, int rows,
                         int cols) {
  __shared__ float s_col[32][kHeadMaxCols + 1];
  __shared__ float s_loss[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float inv_rows = 1.f / (float)rows;
  float colacc[kHeadMaxCols / 32];
#pragma unroll
  for (int j = 0; j < kHeadMaxCols / 32; ++j) colacc[j] = 0.f;
  float loss_acc = 0.f;
  for (int r = warp; r < rows; r += nwarps) {
    auto load = [&](int c) -> float {
      if (BF16_IN) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(logits_)[(size_t)r * ld_in + c]);
      return reinterpret_cast<const float*>(logits_)[(size_t)r * ld_in + c];
    };
    float z[kHeadMaxCols / 32];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kHeadMaxCols / 32; ++j) {
      const int c = lane + 32 * j;
      z[j] = c < cols ? load(c) : -INFINITY;
      m = fmaxf(m, z[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kHeadMaxCols / 32; ++j)
      if (lane + 32 * j < cols) s += __expf(z[j] - m);
    s = warp_sum(s);
    const int label = (int)labels[r];
    const float inv_s = 1.f / s;
#pragma unroll
    for (int j = 0; j < kHeadMaxCols / 32; ++j) {
      const int c = lane + 32 * j;
      if (c < cols) {
        const float g = (__expf(z[j] - m) * inv_s - (c == label ? 1.f : 0.f)) * inv_rows;
        if (dl) dl[(size_t)r * ld_dl + c] = g;
        if (dl_bf16) dl_bf16[(size_t)r * ld_bf16 + c] = __float2bfloat16(g);
        colacc[j] += g;
        if (c == label) loss_acc += (__logf(s) + m) - z[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kHeadMaxCols / 32; ++j) s_col[warp][lane + 32 * j] = colacc[j];
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) s_loss[warp] = loss_acc;
  __syncthreads();
  if (db) {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += s_col[w][c];
      db[c] = t;
    }
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += s_loss[w];
    *loss_out = t * inv_rows;
  }
}

__global__ void eval_binary_kernel(const float* __restrict__ p, const float* __restrict__ y,
                                   float* __restrict__ loss_sum, int* __restrict__ correct, int64_t n) {
  __shared__ float scratch[32];
  float acc = 0.f;
  int ok = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float q = p[i], t = y[i];
    const float lp = fmaxf(__logf(q), -100.f), l1p = fmaxf(log1pf(-q), -100.f);
    acc -= t * lp + (1.f - t) * l1p;
    ok += (rintf(q) == t) ? 1 : 0;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(loss_sum, acc);
  ok = __reduce_add_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && ok) atomicAdd(correct, ok);
}

__global__ void argmax_rows_kernel(const float* __restrict__ x, int64_t* __restrict__ out, int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) {
      const float v = x[(size_t)r * cols + c];
      if (v > best || (v == best && c < bi)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) out[r] = bi;
  }
}

// MinMax: one block per column pass 1 (min/max), pass 2 scale.  cols is tiny (10).
__global__ void minmax_scale_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols) {
  __shared__ float smin[32], smax[32];
  const int c = blockIdx.x;
  float lo = INFINITY, hi = -INFINITY;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float v = x[(size_t)r * cols + c];
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { smin[w] = lo; smax[w] = hi; }
  __syncthreads();
  lo = smin[0]; hi = smax[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
  float rng = hi - lo;
  if (rng == 0.f) rng = 1.f;
  const float inv = 1.f / rng;
  for (int r = threadIdx.x; r < rows; r += blockDim.x)
    out[(size_t)r * cols + c] = (x[(size_t)r * cols + c] - lo) * inv;
}

// Keyed random permutations, one row per epoch (the bijection itself: feistel_index in colearn_kernels.h)
__global__ void feistel_perm_kernel(int* __restrict__ out, int n, int rows, uint64_t seed) {
  const FeistelDomain dom = feistel_domain((uint32_t)n);
  const int64_t total = (int64_t)n * rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int row = (int)(e / n);
    out[e] = (int)feistel_index((uint32_t)(e - (int64_t)row * n), (uint32_t)n, dom, seed, row);
  }
}

__global__ void fp32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((((uintptr_t)in) & 15) == 0) && ((((uintptr_t)out) & 7) == 0)) {
    const int64_t n4 = n >> 2;
    for (int64_t j = i0; j < n4; j += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in) + j);
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(out)[j] = pk;
    }
    for (int64_t j = (n4 << 2) + i0; j < n; j += stride) out[j] = __float2bfloat16(in[j]);
  } else {
    for (int64_t j = i0; j < n; j += stride) out[j] = __float2bfloat16(in[j]);
  }
}

__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

// ---- int64 ring ops for the SMPC path: Z_{2^64} arithmetic on CUDA cores ---------------------
__global__ void fix_precision_kernel(const float* __restrict__ x, long long* __restrict__ out, int64_t n, double base) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (long long)rint((double)x[i] * base);   // round-half-even, same as torch.round
}
__global__ void float_precision_kernel(const long long* __restrict__ x, float* __restrict__ out, int64_t n, double inv_base) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (float)((double)x[i] * inv_base);
}
// C[m,n] = sum_k A[m,k] * B[k,n]  (mod 2^64; unsigned wrap-around == two's complement ring arithmetic)
__global__ void ring_matmul_kernel(const unsigned long long* __restrict__ A, const unsigned long long* __restrict__ B,
                                   unsigned long long* __restrict__ C, int M, int K, int N) {
  __shared__ unsigned long long sa[16][17], sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  unsigned long long acc = 0ull;
  for (int k0 = 0; k0 < K; k0 += 16) {
    sa[ty][tx] = (row < M && k0 + tx < K) ? A[(size_t)row * K + k0 + tx] : 0ull;
    sb[ty][tx] = (k0 + ty < K && col < N) ? B[(size_t)(k0 + ty) * N + col] : 0ull;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[k][tx];
    __syncthreads();
  }
  if (row < M && col < N) C[(size_t)row * N + col] = acc;
}

__global__ void bias_sgd_from_partials_kernel(float* __restrict__ bias, const float* __restrict__ partials, int rows, int n,
                                              int64_t row_stride, float lr, float* __restrict__ grad_out, int n_bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float g = 0.f;
  for (int r = 0; r < rows; ++r) g += partials[(size_t)r * row_stride + c];
  if (grad_out) grad_out[c] = g;
  if (bias && c < n_bias) bias[c] = fmaf(-lr, g, bias[c]);
}

__global__ void l2_flush_kernel(float* __restrict__ buf, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = (float)i;
}

}  // namespace

cudaError_t launch_sgd_step(float* p, const float* g, float lr, int64_t n, cudaStream_t s) {
  COLEARN_LAUNCH(sgd_step_kernel, grid_for(n), kThreads, 0, s, p, g, lr, n);
  return cudaGetLastError();
}
cudaError_t launch_scale_inplace(float* p, float scale, int64_t n, cudaStream_t s) {
  COLEARN_LAUNCH(scale_inplace_kernel, grid_for(n), kThreads, 0, s, p, scale, n);
  return cudaGetLastError();
}
cudaError_t launch_sgd_step_bf16grad(float* p, const void* g, float lr, int64_t n, cudaStream_t s) {
  COLEARN_LAUNCH(sgd_step_bf16grad_kernel, grid_for(n), kThreads, 0, s, p, reinterpret_cast<const __nv_bfloat16*>(g), lr, n);
  return cudaGetLastError();
}
cudaError_t launch_fedavg_apply(float* theta, const float* slots, int64_t slot_stride, const float* weights,
                                int k, float server_lr, int64_t n, cudaStream_t s) {
  if (k < 1) return cudaErrorInvalidValue;
  COLEARN_LAUNCH(fedavg_kernel<true>, grid_for(n), kThreads, 0, s, theta, slots, slot_stride, weights, k, server_lr, n);
  return cudaGetLastError();
}
cudaError_t launch_fedavg_flat(float* out, const float* slots, int64_t slot_stride, const float* weights,
                               int k, int64_t n, cudaStream_t s) {
  if (k < 1) return cudaErrorInvalidValue;
  COLEARN_LAUNCH(fedavg_kernel<false>, grid_for(n), kThreads, 0, s, out, slots, slot_stride, weights, k, 1.f, n);
  return cudaGetLastError();
}
cudaError_t launch_sigmoid_bce(const float* z, const float* y, float* dz, float* loss_out, int64_t n, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(loss_out, 0, sizeof(float), s);
  if (e != cudaSuccess) return e;
  COLEARN_LAUNCH(sigmoid_bce_kernel, grid_for(n, 1), kThreads, 0, s, z, y, dz, loss_out, n);
  return cudaGetLastError();
}
cudaError_t launch_sse(const float* out, const float* y, float* dz, float* loss_out, int64_t n, float scale, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(loss_out, 0, sizeof(float), s);
  if (e != cudaSuccess) return e;
  COLEARN_LAUNCH(sse_kernel, grid_for(n, 1), kThreads, 0, s, out, y, dz, loss_out, n, scale);
  return cudaGetLastError();
}
cudaError_t launch_softmax_xent(const void* logits, int is_bf16, const int64_t* labels, float* dl, void* dl_bf16,
                                float* loss_out, int rows, int cols, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(loss_out, 0, sizeof(float), s);
  if (e != cudaSuccess) return e;
  const int wpb = kThreads / 32;
  int blocks = (rows + wpb - 1) / wpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (is_bf16)
    COLEARN_LAUNCH(softmax_xent_kernel<true>, blocks, kThreads, 0, s, logits, labels, dl, reinterpret_cast<__nv_bfloat16*>(dl_bf16), loss_out, rows, cols);
  else
    COLEARN_LAUNCH(softmax_xent_kernel<false>, blocks, kThreads, 0, s, logits, labels, dl, reinterpret_cast<__nv_bfloat16*>(dl_bf16), loss_out, rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_softmax_xent_head(const void* logits, int is_bf16, int ld_in, const int64_t* labels, float* dl, int ld_dl, void* dl_bf16,
                                     int ld_bf16, float* db, float* loss_out, int rows, int cols, cudaStream_t s) {
  if (cols < 1 || cols > kHeadMaxCols || rows < 1) return cudaErrorInvalidValue;
  int threads = 32 * (rows < 32 ? rows : 32);          // one warp per row in flight, at most 32 warps
  if (is_bf16)
    COLEARN_LAUNCH(softmax_xent_head_kernel<true>, 1, threads, 0, s, logits, ld_in, labels, dl, ld_dl, reinterpret_cast<__nv_bfloat16*>(dl_bf16),
                   ld_bf16, db, loss_out, rows, cols);
  else
    COLEARN_LAUNCH(softmax_xent_head_kernel<false>, 1, threads, 0, s, logits, ld_in, labels, dl, ld_dl, reinterpret_cast<__nv_bfloat16*>(dl_bf16),
                   ld_bf16, db, loss_out, rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_eval_binary(const float* p, const float* y, float* loss_sum, int* correct, int64_t n, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(loss_sum, 0, sizeof(float), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(correct, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  COLEARN_LAUNCH(eval_binary_kernel, grid_for(n, 1), kThreads, 0, s, p, y, loss_sum, correct, n);
  return cudaGetLastError();
}
cudaError_t launch_argmax_rows(const float* x, int64_t* out, int rows, int cols, cudaStream_t s) {
  const int wpb = kThreads / 32;
  int blocks = (rows + wpb - 1) / wpb;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  COLEARN_LAUNCH(argmax_rows_kernel, blocks, kThreads, 0, s, x, out, rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_minmax_scale(const float* x, float* out, int rows, int cols, cudaStream_t s) {
  COLEARN_LAUNCH(minmax_scale_kernel, cols, 1024, 0, s, x, out, rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_feistel_permutation(int* out, int n, int rows, uint64_t seed, cudaStream_t s) {
  COLEARN_LAUNCH(feistel_perm_kernel, grid_for((int64_t)n * rows, 1), kThreads, 0, s, out, n, rows, seed);
  return cudaGetLastError();
}
cudaError_t launch_fp32_to_bf16(const float* in, void* out, int64_t n, cudaStream_t s) {
  COLEARN_LAUNCH(fp32_to_bf16_kernel, grid_for(n), kThreads, 0, s, in, reinterpret_cast<__nv_bfloat16*>(out), n);
  return cudaGetLastError();
}
cudaError_t launch_transpose_bf16(const void* in, void* out, int rows, int cols, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  COLEARN_LAUNCH(transpose_bf16_kernel, grid, block, 0, s, reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), rows, cols);
  return cudaGetLastError();
}
cudaError_t launch_fix_precision(const float* x, long long* out, int64_t n, double base, cudaStream_t s) {
  COLEARN_LAUNCH(fix_precision_kernel, grid_for(n, 1), kThreads, 0, s, x, out, n, base);
  return cudaGetLastError();
}
cudaError_t launch_float_precision(const long long* x, float* out, int64_t n, double inv_base, cudaStream_t s) {
  COLEARN_LAUNCH(float_precision_kernel, grid_for(n, 1), kThreads, 0, s, x, out, n, inv_base);
  return cudaGetLastError();
}
cudaError_t launch_ring_matmul(const long long* A, const long long* B, long long* C, int M, int K, int N, cudaStream_t s) {
  dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
  COLEARN_LAUNCH(ring_matmul_kernel, grid, block, 0, s, reinterpret_cast<const unsigned long long*>(A),
                 reinterpret_cast<const unsigned long long*>(B), reinterpret_cast<unsigned long long*>(C), M, K, N);
  return cudaGetLastError();
}
cudaError_t launch_bias_sgd_from_partials(float* bias, const float* partials, int rows, int n, int64_t row_stride, float lr,
                                          float* grad_out, int n_bias, cudaStream_t s) {
  COLEARN_LAUNCH(bias_sgd_from_partials_kernel, (n + 255) / 256, 256, 0, s, bias, partials, rows, n, row_stride, lr, grad_out, n_bias);
  return cudaGetLastError();
}
cudaError_t launch_l2_flush(float* buf, int64_t n, cudaStream_t s) {
  COLEARN_LAUNCH(l2_flush_kernel, 148 * 8, kThreads, 0, s, buf, n);
  return cudaGetLastError();
}

}  // namespace colearn

// CPU emulator of the convnet.cu kernels: the same host+device bodies (conv_ops.cuh), driven by plain loops
// that mimic the grid (maps: one iteration per work item; two-phase blocks: all threads of phase 1, then all
// threads of phase 2, sharing a block-local array).  Built with g++ only for the CPU test-suite
// (ops/build.py::build_emul); never used on a GPU box.
#include <torch/extension.h>

#include <cstring>
#include <vector>

#include "conv_ops.cuh"

namespace py = pybind11;

namespace {
using namespace colearn::convops;

template <class Args, class Items, class Body>
inline void run_map(const Args& a, Items items, Body body) {
  const long long n = items(a);
  for (long long i = 0; i < n; ++i) body(a, i);
}

struct ConvHostExec {
  static constexpr bool kCuda = false;
  static void im2col(const Im2colArgs& a) { run_map(a, im2col_items, im2col_body); }
  static void col2im(const Col2imArgs& a) { run_map(a, col2im_items, col2im_body); }
  static void bn_reduce(const BnReduceArgs& a) {
    std::vector<float> smem(kBnSmemFloats);
    const int nseg = bn_nseg(a);
    for (int by = 0; by < nseg; ++by)
      for (int bx = 0; bx < a.C / kBnCols; ++bx) {
        for (int tid = 0; tid < kBnThreads; ++tid) bn_reduce_phase1(a, bx, by, tid, smem.data());
        for (int tid = 0; tid < kBnThreads; ++tid) bn_reduce_phase2(a, bx, by, tid, smem.data());
      }
  }
  static void bn_finalize(const BnFinalizeArgs& a) {
    for (int c = 0; c < a.C; ++c) bn_finalize_body<false>(a, c);
  }
  // blocks run one after the other here, so "the last ticket" is simply the last block of each column group
  static void bn_reduce_finalize(const BnFusedArgs& a) {
    std::vector<float> smem(kBnSmemFloats);
    const int nseg = bn_nseg(a.r);
    for (int by = 0; by < nseg; ++by)
      for (int bx = 0; bx < a.r.C / kBnCols; ++bx) {
        for (int tid = 0; tid < kBnThreads; ++tid) bn_reduce_phase1(a.r, bx, by, tid, smem.data());
        for (int tid = 0; tid < kBnThreads; ++tid) bn_reduce_phase2(a.r, bx, by, tid, smem.data());
        if (++a.counters[bx] == (unsigned)nseg) {
          for (int tid = 0; tid < kBnThreads; ++tid) bn_fused_phase3(a, bx, tid);
          a.counters[bx] = 0;
        }
      }
  }
  static void bn_apply(const BnApplyArgs& a) { run_map(a, bn_apply_items, bn_apply_body); }
  static void bn_bwd(const BnBwdArgs& a) { run_map(a, bn_bwd_items, bn_bwd_body); }
  static void maxpool_fwd(const PoolArgs& a) { run_map(a, maxpool_fwd_items, maxpool_fwd_body); }
  static void maxpool_bwd(const PoolArgs& a) { run_map(a, maxpool_bwd_items, maxpool_bwd_body); }
  static void avgpool_fwd(const AvgPoolArgs& a) { run_map(a, avgpool_fwd_items, avgpool_fwd_body); }
  static void avgpool_bwd(const AvgPoolArgs& a) { run_map(a, avgpool_bwd_items, avgpool_bwd_body); }
  static void pack(const PackArgs& a) {
    for (long long e = 0; e < a.total; ++e) pack_body(a, e);
  }
  static void splitk_reduce(const SplitKReduceArgs& a) { run_map(a, splitk_reduce_items, splitk_reduce_body); }
};
}  // namespace

#include "conv_bindings.inc"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "CPU emulator of the colearn conv / BatchNorm / pooling kernels (tests only)";
  convbind::register_ops<ConvHostExec>(m);
}

// CPU build of the persistent-MLP kernels (mlp_persistent.cu + mlp_v2.inc) through host_shim.h: the same source the GPU
// runs, one OS thread per CUDA thread.  Tests only (ops/build.py::build_simt_emul, tests/test_simt_emul.py).
#define COLEARN_HOST_SHIM 1
#include <torch/extension.h>

#include "host_shim.h"

#include "mlp_persistent.cu"   // NOLINT(bugprone-suspicious-include): the kernel source itself

namespace py = pybind11;

namespace {
using torch::Tensor;

const float* fptr(const Tensor& t, const char* name) {
  TORCH_CHECK(!t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), name, ": contiguous CPU fp32");
  return t.data_ptr<float>();
}

// K clients (one emulated CTA each, run one after the other).  theta_outs[i] may alias theta_in.
// Returns losses [K, 2] = {last, mean}; flags (int32 [K], optional) receive signal_value = 7 + i.
Tensor mlp_local_sgd(int64_t net_kind, Tensor theta_in, std::vector<Tensor> theta_outs, std::vector<Tensor> xs, std::vector<Tensor> ys,
                     std::vector<c10::optional<Tensor>> perms, int64_t batch_size, int64_t epochs, int64_t max_steps, int64_t loss,
                     double lr, int64_t variant, std::vector<double> out_scales, bool delta_mode, c10::optional<Tensor> flags) {
  const size_t K = xs.size();
  TORCH_CHECK(K >= 1 && ys.size() == K && perms.size() == K && theta_outs.size() == K && out_scales.size() == K, "one entry per client");
  const int P = colearn::mlp_net_num_params((int)net_kind);
  TORCH_CHECK(P > 0 && theta_in.numel() == P, "theta_in must hold ", P, " parameters");
  Tensor losses = torch::zeros({(int64_t)K, 2}, torch::kFloat);
  std::vector<colearn::ClientDesc> descs(K);
  for (size_t i = 0; i < K; ++i) {
    colearn::ClientDesc& d = descs[i];
    memset(&d, 0, sizeof(d));
    d.x = fptr(xs[i], "x");
    d.y = fptr(ys[i], "y");
    TORCH_CHECK(xs[i].dim() == 2 && ys[i].dim() == 2 && xs[i].size(0) == ys[i].size(0), "x [n, d_in], y [n, y_dim]");
    d.n = (int)xs[i].size(0);
    d.y_dim = (int)ys[i].size(1);
    d.perm_rows = 1;
    if (perms[i].has_value()) {
      const Tensor& p = *perms[i];
      TORCH_CHECK(!p.is_cuda() && p.scalar_type() == at::kInt && p.is_contiguous() && p.dim() == 2 && p.size(1) == d.n, "perm int32 [rows, n]");
      d.perm = p.data_ptr<int>();
      d.perm_rows = (int)p.size(0);
    }
    d.theta_in = fptr(theta_in, "theta_in");
    TORCH_CHECK(theta_outs[i].numel() == P, "theta_out size");
    d.theta_out = const_cast<float*>(fptr(theta_outs[i], "theta_out"));
    d.loss_out = losses.data_ptr<float>() + 2 * i;
    d.out_scale = (float)out_scales[i];
    d.delta_mode = delta_mode ? 1 : 0;
    if (flags.has_value()) {
      TORCH_CHECK(flags->scalar_type() == at::kInt && flags->is_contiguous() && flags->numel() >= (int64_t)K, "flags int32 [K]");
      d.signal_flag = reinterpret_cast<uint32_t*>(flags->data_ptr<int>()) + i;
      d.signal_value = 7u + (uint32_t)i;
    }
  }
  colearn::SgdHyper hp;
  hp.batch_size = (int)batch_size;
  hp.epochs = (int)epochs;
  hp.max_steps = (int)max_steps;
  hp.loss = (int)loss;
  hp.lr = (float)lr;
  hp.variant = (int)variant;
  {
    py::gil_scoped_release nogil;
    cudaError_t e = colearn::launch_mlp_local_sgd((int)net_kind, descs.data(), (int)K, hp, nullptr);
    TORCH_CHECK(e == cudaSuccess, "launch_mlp_local_sgd (host shim) failed");
  }
  return losses;
}

Tensor mlp_forward(int64_t net_kind, Tensor theta, Tensor x, int64_t d_out) {
  TORCH_CHECK(x.dim() == 2, "x [n, d_in]");
  Tensor out = torch::zeros({x.size(0), d_out}, torch::kFloat);
  {
    py::gil_scoped_release nogil;
    cudaError_t e = colearn::launch_mlp_forward((int)net_kind, fptr(theta, "theta"), fptr(x, "x"), out.data_ptr<float>(), (int)x.size(0), nullptr);
    TORCH_CHECK(e == cudaSuccess, "launch_mlp_forward (host shim) failed");
  }
  return out;
}
}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "persistent-MLP CUDA kernels compiled for the CPU through a SIMT shim (tests only)";
  m.def("mlp_local_sgd", &mlp_local_sgd);
  m.def("mlp_forward", &mlp_forward);
  m.def("mlp_net_params", [](int64_t kind) { return (int64_t)colearn::mlp_net_num_params((int)kind); });
}

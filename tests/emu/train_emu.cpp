// CPU emulation of the training-step kernel chain (TEST INFRASTRUCTURE ONLY; never linked into libcatre_b200.so).
// Compiles catre_b200/csrc/train_kernels.cuh + train_chain.cuh with CATRE_HOST_EMU: every kernel functor runs as
// nested loops over (block, thread), so the kernels' indexing and the host orchestration are checked against the
// oracle without a GPU (tests/test_train_emu.py).  Build: g++ -O2 -shared -fPIC -DCATRE_HOST_EMU train_emu.cpp
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../catre_b200/csrc/train_chain.cuh"

using namespace catre_train;

namespace {
struct EmuOps {
  long launches = 0;
  double gemm_macs = 0.0;  // multiply-adds of all GEMM launches (the training step's algorithmic work)
  template <class KF>
  void run(const KF& k, unsigned gx, unsigned gy, unsigned gz, unsigned nt) {
    ++launches;
    for (unsigned z = 0; z < gz; ++z)
      for (unsigned y = 0; y < gy; ++y)
        for (unsigned x = 0; x < gx; ++x)
          for (unsigned t = 0; t < nt; ++t) k(Idx{(int)x, (int)y, (int)z, (int)t, (int)nt});
  }
  void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
  void gemm(const GemmP& p, int batch_or_splits) {
    gemm_macs += (double)p.M * p.N * p.K * (p.splits > 1 ? 1 : batch_or_splits);
    run(KGemmNaive{p}, (unsigned)((p.M + 3) / 4), (unsigned)((p.N + 63) / 64), (unsigned)batch_or_splits, 256);
  }
};
}  // namespace

extern "C" int emu_train_step(const float* const* weights, int B, int N, const float* pcl, const float* kps, const float* pose,
                              const float* scale, const float* K, const float* gt_pose, const float* gt_scale,
                              const unsigned char* is_sym, const float* sym_rots, int n_rots, float* pose_out, float* scale_out,
                              float* losses, float* const* grads, long* launches, const float* x_pm, const float* tfd_pm, double* gemm_macs, const float* loss_w) {
  if (n_rots > TrainWs::kMaxSymRots) return -1;
  TrainWs w;
  const size_t bytes = ws_layout(w, B, N, nullptr);
  std::vector<char> mem(bytes + 256);
  char* base = reinterpret_cast<char*>(((uintptr_t)mem.data() + 255) & ~(uintptr_t)255);
  ws_layout(w, B, N, base);
  memcpy(w.is_sym, is_sym, B);
  memcpy(w.sym_rots, sym_rots, (size_t)n_rots * 9 * sizeof(float));
  int n_sym = 0;
  for (int b = 0; b < B; ++b) n_sym += is_sym[b] != 0;
  EmuOps ops;
  Chain<EmuOps> c{ops, w, weights, N};
  TrainIn in{pcl, kps, pose, scale, K, gt_pose, gt_scale, B, n_rots, n_sym, B - n_sym, pose_out, scale_out};
  in.x_pm = x_pm; in.tfd_pm = tfd_pm;
  if (loss_w) { in.w_pm = loss_w[0]; in.w_rot = loss_w[1]; in.w_trans = loss_w[2]; in.w_scale = loss_w[3]; }
  c.forward(in);
  c.loss(in);
  c.backward(in);
  memcpy(losses, w.losses, 6 * sizeof(float));
  for (int i = 0; i < W_COUNT; ++i) memcpy(grads[i], w.G[i], weight_numel(i, N) * sizeof(float));
  if (launches) *launches = ops.launches;
  if (gemm_macs) *gemm_macs = ops.gemm_macs;
  return 0;
}

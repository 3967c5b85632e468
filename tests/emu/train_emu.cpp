// CPU emulation of the training-step kernel chain (TEST INFRASTRUCTURE ONLY; never linked into libcatre_b200.so).
// Compiles catre_b200/csrc/train_kernels.cuh + train_chain.cuh with CATRE_HOST_EMU: every kernel functor runs as
// nested loops over (block, thread), so the kernels' indexing and the host orchestration are checked against the
// oracle without a GPU (tests/test_train_emu.py).  Build: g++ -O2 -shared -fPIC -DCATRE_HOST_EMU train_emu.cpp
#include <stdlib.h>
#include <string.h>

#include <vector>

#ifdef CATRE_EMU_ASAN
#include <sanitizer/asan_interface.h>
#endif
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
  void zero(void* p, size_t bytes) {
#ifdef CATRE_EMU_ASAN  // the product clears the whole gradient arena in one memset; here the red zones in it stay poisoned
    char* c = static_cast<char*>(p);
    for (size_t i = 0; i < bytes; ++i)
      if (!__asan_address_is_poisoned(c + i)) c[i] = 0;
#else
    memset(p, 0, bytes);
#endif
  }
  bool gemm_colmax(const GemmP&, int, float*, int*, int, size_t) { return false; }  // no fused pooling here: layer + colmax
  void fork() {}  // lanes are an execution detail of the GPU launcher: here everything runs in program order
  void lane(int) {}
  void join() {}
  bool folds_bias_grad(const GemmP&) { return false; }                               // nor bias gradients inside the GEMM
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
#ifdef CATRE_EMU_ASAN  // sanitizer build: every workspace slice is followed by a poisoned 4 KB red zone
  const size_t gap = 4096;
#else
  const size_t gap = 0;
#endif
  const size_t bytes = ws_layout(w, B, N, nullptr, gap);
  std::vector<char> mem(bytes + 256);
  char* base = reinterpret_cast<char*>(((uintptr_t)mem.data() + 255) & ~(uintptr_t)255);
  ws_layout(w, B, N, base, gap);
#ifdef CATRE_EMU_ASAN
  {
    // poison everything, then unpoison exactly the bytes each slice owns (recomputed with the same walk)
    __asan_poison_memory_region(base, bytes);
    struct Slice { char* p; size_t n; };
    std::vector<Slice> sl;
    const size_t Bz = B, S = 2 * Bz, R = S * N;
    auto F = [&](float* p, size_t n) { sl.push_back({reinterpret_cast<char*>(p), n * sizeof(float)}); };
    auto I = [&](int* p, size_t n) { sl.push_back({reinterpret_cast<char*>(p), n * sizeof(int)}); };
    F(w.q, R * 3); F(w.s64, R * 64); F(w.s128, R * 128); F(w.zbuf, R * 1024); F(w.smax, S * 1024); I(w.sarg, S * 1024);
    F(w.sfc1, S * 512); F(w.sfc2, S * 256); F(w.t3, S * 9); F(w.qp, R * 3); F(w.h1, R * 64); F(w.f64, R * 64); F(w.f128, R * 128);
    F(w.fmax, S * 1024); I(w.farg, S * 1024); F(w.ffc1, S * 512); F(w.ffc2, S * 256); F(w.t64, S * 4096); F(w.pf, R * 64);
    F(w.a128, R * 128); F(w.a512, R * 512); F(w.g, S * 1024); I(w.garg, S * 1024); F(w.pfmax, S * 64); I(w.pfarg, S * 64);
    F(w.ts_in, Bz * 1091); F(w.ts_y0, Bz * 256); F(w.ts_u0, Bz * 256); F(w.ts_y1, Bz * 256); F(w.ts_u1, Bz * 256);
    F(w.ts_st0, Bz * 64); F(w.ts_st1, Bz * 64); F(w.dts, Bz * 6);
    for (int h = 0; h < 2; ++h) {
      F(w.cset[h], S * 256); F(w.ry0[h], R * 256); F(w.ru0[h], R * 256); F(w.ry1[h], R * 256); F(w.rst0[h], Bz * 64);
      F(w.rst1[h], Bz * 64); F(w.wsum[h], Bz * 256); F(w.ru1[h], R * 256);
    }
    F(w.r6, Bz * 6); F(w.swp, 2);
    F(w.lossp, Bz * 6); F(w.losses, 8); F(w.dpose, Bz * 15); F(w.d_r6, Bz * 6); F(w.d_dts, Bz * 6); F(w.tsd_u, Bz * 256);
    F(w.tsd_u0, Bz * 256); F(w.ts_din, Bz * 1091); F(w.gn_m, Bz * 64); F(w.gnp_g, Bz * 256); F(w.gnp_b, Bz * 256);
    F(w.dg, S * 1024); F(w.dpfmax, S * 64); F(w.dpf, R * 64);
    for (int h = 0; h < 2; ++h) { F(w.e[h], Bz * 256); F(w.du[h], R * 256); F(w.du0[h], R * 256); F(w.dcset[h], S * 256); }
    F(w.d512, R * 512); F(w.d128, R * 128); F(w.d64, R * 64); F(w.dh1, R * 64); F(w.dt64, S * 4096);
    F(w.dfc2, S * 256); F(w.dfc1, S * 512); F(w.dmax, S * 1024); F(w.dqp, R * 3); F(w.dt3, S * 9);
    F(w.partial, TrainWs::kPartialFloats); F(w.cs_partial, TrainWs::kCsFloats); F(w.loss_gs, Bz * 9);
    F(w.partial_ts, TrainWs::kPartialFloats); F(w.cs_partial_ts, TrainWs::kCsFloats); F(w.gn_m_ts, Bz * 64); F(w.gnp_g_ts, Bz * 256);
    F(w.gnp_b_ts, Bz * 256);
    sl.push_back({reinterpret_cast<char*>(w.gn_part_ts), Bz * TrainWs::kGnChunks * 32 * 18 * sizeof(double)});
    F(w.partial_l2, TrainWs::kPartialFloats); F(w.cs_partial_l2, TrainWs::kCsFloats); F(w.gn_m_l2, Bz * 64); F(w.gnp_g_l2, Bz * 256);
    F(w.gnp_b_l2, Bz * 256);
    sl.push_back({reinterpret_cast<char*>(w.gn_part_l2), Bz * 32 * 18 * sizeof(double)});
    for (int i = 0; i < 3; ++i) { I(w.mb_start[i], S * N); I(w.mb_cnt[i], S * N); I(w.mb_list[i], S * 1024); I(w.mb_key[i], S * 1024); }
    sl.push_back({reinterpret_cast<char*>(w.gn_part), Bz * TrainWs::kGnChunks * 32 * 18 * sizeof(double)});
    sl.push_back({reinterpret_cast<char*>(w.is_sym), Bz});
    F(w.sym_rots, (size_t)TrainWs::kMaxSymRots * 9);
    for (int i = 0; i < W_COUNT; ++i) F(w.G[i], weight_numel(i, N));
    for (auto& x : sl) __asan_unpoison_memory_region(x.p, x.n);
  }
#endif
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

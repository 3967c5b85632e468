// Host orchestration of the training step: forward with saved activations, losses, backward.
// Mirrors oracle/train_oracle.py::manual_forward / loss_backward / pose_backward / manual_backward line by line.
// Written against the small `Ops` interface so the same code drives CUDA launches in libcatre_b200.so and plain
// loops in the CPU emulation build used by tests/ (CATRE_HOST_EMU; see train_kernels.cuh).
#pragma once
#include <stddef.h>
#include <string.h>

#include "train_kernels.cuh"

namespace catre_train {

// checkpoint order of the 74 tensors (catre_api.cu kWeights)
enum : int {
  W_STN = 0, W_CONV1 = 12, W_CONV2 = 14, W_CONV3 = 16, W_CONV4 = 18, W_FSTN = 20, W_ROT_X = 32, W_ROT_Y = 46, W_TS = 60,
  W_COUNT = 74
};
// offsets inside a T-Net block (weight; bias = +1)
enum : int { T_CONV1 = 0, T_CONV2 = 2, T_CONV3 = 4, T_FC1 = 6, T_FC2 = 8, T_FC3 = 10 };
// offsets inside a rotation-head block
enum : int { R_L0 = 2, R_GN0 = 4, R_L3 = 6, R_GN1 = 8, R_NECK = 10, R_CONVP = 12 };
// offsets inside the ts-head block (base 60)
enum : int { S_L0 = 2, S_GN0 = 4, S_L3 = 6, S_GN1 = 8, S_FCT = 10, S_FCS = 12 };

// number of elements of checkpoint tensor i for N points per set
inline size_t weight_numel(int i, int N) {
  static const int kStn[12] = {64 * 3, 64, 128 * 64, 128, 1024 * 128, 1024, 512 * 1024, 512, 256 * 512, 256, 9 * 256, 9};
  static const int kFstn[12] = {64 * 64, 64, 128 * 64, 128, 1024 * 128, 1024, 512 * 1024, 512, 256 * 512, 256, 4096 * 256, 4096};
  static const int kTrunk[8] = {64 * 3, 64, 128 * 64, 128, 512 * 128, 512, 1024 * 512, 1024};
  static const int kRot[14] = {256, 256, 256 * 1088, 256, 256, 256, 256 * 256, 256, 256, 256, 3 * 256, 3, -1, 1};
  static const int kTs[14] = {256, 256, 256 * 1091, 256, 256, 256, 256 * 256, 256, 256, 256, 3 * 256, 3, 3 * 256, 3};
  if (i < 12) return kStn[i];
  if (i < 20) return kTrunk[i - 12];
  if (i < 32) return kFstn[i - 20];
  if (i < 60) { const int v = kRot[(i - 32) % 14]; return v < 0 ? (size_t)2 * N : (size_t)v; }
  return kTs[i - 60];
}

struct TrainWs {
  int maxB = 0, N = 0;
  // forward, saved
  float *q, *s64, *s128, *zbuf, *smax, *sfc1, *sfc2, *t3, *qp, *h1, *f64, *f128, *fmax, *ffc1, *ffc2, *t64, *pf, *a128, *a512, *g,
      *pfmax;
  int *sarg, *farg, *garg, *pfarg;
  float *ts_in, *ts_y0, *ts_u0, *ts_y1, *ts_u1, *ts_st0, *ts_st1, *dts;
  float *cset[2], *ry0[2], *ru0[2], *ry1[2], *rst0[2], *rst1[2], *wsum[2], *ru1[2], *r6, *swp;
  // loss / backward
  float *lossp, *losses, *dpose, *d_r6, *d_dts, *tsd_u, *tsd_u0, *ts_din, *gn_m, *gnp_g, *gnp_b;
  float *dg, *dpfmax, *dpf, *e[2], *du[2], *du0[2], *dcset[2], *d512, *d128, *d64, *dh1, *dt64, *dfc2, *dfc1, *dmax, *dqp, *dt3;
  float *partial, *cs_partial, *loss_gs;
  // second scratch set: the ts head and the y rotation head run as a side lane next to the x rotation head (Chain::forward /
  // backward); the per-head buffers above ([2]) exist for the same reason
  float *partial_ts, *cs_partial_ts, *gn_m_ts, *gnp_g_ts, *gnp_b_ts;
  double* gn_part_ts;
  // third set: the ts head's own lane (GroupNorm over one "point": one chunk)
  float *partial_l2, *cs_partial_l2, *gn_m_l2, *gnp_g_l2, *gnp_b_l2;
  double* gn_part_l2;
  // inverse arg-max maps of the sparse max-pool backward, one per max-pooled layer (stn, fstn, trunk): [S, N], [S, N], [S, 1024] x 2
  int *mb_start[3], *mb_cnt[3], *mb_list[3], *mb_key[3];
  double* gn_part;  // [maxB, kGnChunks, 32, 18]; also the point-matching partials of the loss [maxB, kLossChunks, 13]
  unsigned char* is_sym;
  float* sym_rots;
  float* G[W_COUNT];  // gradient arena, checkpoint order
  size_t grad_floats = 0;
  static constexpr size_t kPartialFloats = (size_t)4 << 20;
  static constexpr int kMaxSymRots = 1024;
  static constexpr int kGnChunks = 256;
  static constexpr int kLossChunks = 32;
  static constexpr size_t kCsFloats = (size_t)64 * 4096;
};

// Assigns every workspace pointer from `base` (256-byte aligned slices) and returns the bytes needed; with
// base == nullptr only the size is computed.
// `gap` bytes are left unused after every slice (0 in the product; the sanitizer build of tests/emu poisons them so an
// out-of-range access of any kernel is caught on the CPU).
inline size_t ws_layout(TrainWs& w, int maxB, int N, char* base, size_t gap = 0) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* {
    char* p = base ? base + off : nullptr;
    off += ((bytes + 255) & ~(size_t)255) + gap;
    return p;
  };
  auto F = [&](float*& p, size_t n) { p = reinterpret_cast<float*>(take(n * sizeof(float))); };
  auto I = [&](int*& p, size_t n) { p = reinterpret_cast<int*>(take(n * sizeof(int))); };
  const size_t B = maxB, S = 2 * B, R = S * N;
  w.maxB = maxB; w.N = N;
  F(w.q, R * 3); F(w.s64, R * 64); F(w.s128, R * 128); F(w.zbuf, R * 1024); F(w.smax, S * 1024); I(w.sarg, S * 1024);
  F(w.sfc1, S * 512); F(w.sfc2, S * 256); F(w.t3, S * 9); F(w.qp, R * 3); F(w.h1, R * 64); F(w.f64, R * 64); F(w.f128, R * 128);
  F(w.fmax, S * 1024); I(w.farg, S * 1024); F(w.ffc1, S * 512); F(w.ffc2, S * 256); F(w.t64, S * 4096); F(w.pf, R * 64);
  F(w.a128, R * 128); F(w.a512, R * 512); F(w.g, S * 1024); I(w.garg, S * 1024); F(w.pfmax, S * 64); I(w.pfarg, S * 64);
  F(w.ts_in, B * 1091); F(w.ts_y0, B * 256); F(w.ts_u0, B * 256); F(w.ts_y1, B * 256); F(w.ts_u1, B * 256);
  F(w.ts_st0, B * 64); F(w.ts_st1, B * 64); F(w.dts, B * 6);
  for (int h = 0; h < 2; ++h) {
    F(w.cset[h], S * 256); F(w.ry0[h], R * 256); F(w.ru0[h], R * 256); F(w.ry1[h], R * 256); F(w.rst0[h], B * 64);
    F(w.rst1[h], B * 64); F(w.wsum[h], B * 256); F(w.ru1[h], R * 256);
  }
  F(w.r6, B * 6); F(w.swp, 2);
  F(w.lossp, B * 6); F(w.losses, 8); F(w.dpose, B * 15); F(w.d_r6, B * 6); F(w.d_dts, B * 6); F(w.tsd_u, B * 256);
  F(w.tsd_u0, B * 256); F(w.ts_din, B * 1091); F(w.gn_m, B * 64); F(w.gnp_g, B * 256); F(w.gnp_b, B * 256);
  F(w.dg, S * 1024); F(w.dpfmax, S * 64); F(w.dpf, R * 64);
  for (int h = 0; h < 2; ++h) { F(w.e[h], B * 256); F(w.du[h], R * 256); F(w.du0[h], R * 256); F(w.dcset[h], S * 256); }
  F(w.d512, R * 512); F(w.d128, R * 128); F(w.d64, R * 64); F(w.dh1, R * 64); F(w.dt64, S * 4096);
  F(w.dfc2, S * 256); F(w.dfc1, S * 512); F(w.dmax, S * 1024); F(w.dqp, R * 3); F(w.dt3, S * 9);
  F(w.partial, TrainWs::kPartialFloats); F(w.cs_partial, TrainWs::kCsFloats); F(w.loss_gs, B * 9);
  F(w.partial_ts, TrainWs::kPartialFloats); F(w.cs_partial_ts, TrainWs::kCsFloats); F(w.gn_m_ts, B * 64); F(w.gnp_g_ts, B * 256);
  F(w.gnp_b_ts, B * 256);
  w.gn_part_ts = reinterpret_cast<double*>(take(B * TrainWs::kGnChunks * 32 * 18 * sizeof(double)));
  F(w.partial_l2, TrainWs::kPartialFloats); F(w.cs_partial_l2, TrainWs::kCsFloats); F(w.gn_m_l2, B * 64); F(w.gnp_g_l2, B * 256);
  F(w.gnp_b_l2, B * 256);
  w.gn_part_l2 = reinterpret_cast<double*>(take(B * 32 * 18 * sizeof(double)));
  for (int i = 0; i < 3; ++i) { I(w.mb_start[i], S * N); I(w.mb_cnt[i], S * N); I(w.mb_list[i], S * 1024); I(w.mb_key[i], S * 1024); }
  w.gn_part = reinterpret_cast<double*>(take(B * TrainWs::kGnChunks * 32 * 18 * sizeof(double)));
  w.is_sym = reinterpret_cast<unsigned char*>(take(B));
  F(w.sym_rots, (size_t)TrainWs::kMaxSymRots * 9);
  const size_t g0 = off;
  for (int i = 0; i < W_COUNT; ++i) F(w.G[i], weight_numel(i, N));
  w.grad_floats = (off - g0) / sizeof(float);
  return off;
}

struct TrainIn {
  // device pointers (host in the emulation).  Points: either the raw clouds (pcl = observed cloud in the camera
  // frame, kps = normalised prior; re-posed here like batching.py:127-140), or -- when x_pm / tfd_pm are given -- the
  // already re-posed sets the reference's forward receives (kps is then only the loss's obj_kps).
  const float *pcl, *kps, *pose, *scale, *K, *gt_pose, *gt_scale;
  int B, n_rots, n_sym, n_nosym;                                    // is_sym / sym_rots already staged in the workspace
  float *pose_out, *scale_out;
  const float *x_pm = nullptr, *tfd_pm = nullptr;
  float w_pm = 1.0f, w_rot = 1.0f, w_trans = 1.0f, w_scale = 1.0f;  // LOSS_CFG.*_LW
};

template <class Ops>
struct Chain {
  Ops& o;
  TrainWs& w;
  const float* const* W;  // the 74 checkpoint tensors
  int N;
  int gemm_f16 = 0;  // operand type of the tensor-core GEMM (GemmP::f16): 1 during forward(), 0 during backward()
  // Scratch of the lane that is being recorded.  Between the encoder and the pose update the ts head and the two rotation heads
  // are independent of each other; most of their launches are a few microseconds long (GroupNorm merges, column sums, the
  // tail), so the ts head and the y head are issued as a SIDE lane next to the x head (Ops::fork / lane / join: a second stream
  // on the GPU, parallel graph branches under capture, plain program order in the emulation).  The side lane has its own
  // split-K / column-sum / GroupNorm scratch and the heads have their own activation and gradient buffers; the two products
  // with which a head adds to the SHARED gradients (dg, dpf) are issued after the join for the y head, so the accumulation
  // order -- x, y, ts -- and with it every bit of the result is the same with one lane or two.
  int lane = 0;       // lane being recorded: 0 main, 1 side (y head, per-layer parameter gradients, forward-time index builds), 2 ts head
  bool split_lin_bwd = false;  // lin_bwd may use the side lane for its parameter gradients (set while no other side lane runs)
  float* sc_partial() const { return lane == 0 ? w.partial : (lane == 1 ? w.partial_ts : w.partial_l2); }
  float* sc_cs_partial() const { return lane == 0 ? w.cs_partial : (lane == 1 ? w.cs_partial_ts : w.cs_partial_l2); }
  double* sc_gn_part() const { return lane == 0 ? w.gn_part : (lane == 1 ? w.gn_part_ts : w.gn_part_l2); }
  float* sc_gn_m() const { return lane == 0 ? w.gn_m : (lane == 1 ? w.gn_m_ts : w.gn_m_l2); }
  float* sc_gnp_g() const { return lane == 0 ? w.gnp_g : (lane == 1 ? w.gnp_g_ts : w.gnp_g_l2); }
  float* sc_gnp_b() const { return lane == 0 ? w.gnp_b : (lane == 1 ? w.gnp_b_ts : w.gnp_b_l2); }
  void begin_side(int i = 1) { o.fork(); o.lane(i); lane = i; }
  void end_side() { lane = 0; o.lane(0); }

  static unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

  // db: for a weight-gradient product (A = dy^T), the bias gradient db[M] += column sums of dy; returns true when the launcher
  // folded them into the GEMM (tensor-core path), false when the caller still has to compute them
  bool gemm(const float* A, long long sam, long long sak, const float* Bm, long long sbk, long long sbn, float* C, long long scm,
            long long scn, int M, int Nn, int K, const float* bias, int relu, int acc, int batch = 1, long long sab = 0,
            long long sbb = 0, long long scb = 0, long long sbias_b = 0, float* db = nullptr) {
    GemmP p{};
    p.A = A; p.sam = sam; p.sak = sak; p.sab = sab; p.B = Bm; p.sbk = sbk; p.sbn = sbn; p.sbb = sbb;
    p.C = C; p.scm = scm; p.scn = scn; p.scb = scb; p.bias = bias; p.sbias_b = sbias_b;
    p.M = M; p.N = Nn; p.K = K; p.relu = relu; p.accumulate = acc; p.splits = 1; p.k_per = K; p.partial = sc_partial();
    p.f16 = gemm_f16;
    if (batch == 1 && !bias && !relu && K >= 4096) {  // weight gradients: reduce over many rows -> split K
      // few output tiles (at most 128 x 128 outputs): more, shorter splits so that the launch still fills the device
      const int per = (M <= 128 && Nn <= 128) ? 256 : 1024, cap = (M <= 128 && Nn <= 128) ? 128 : 32;
      int splits = K / per < cap ? K / per : cap;
      while (splits > 1 && (size_t)splits * M * Nn > TrainWs::kPartialFloats) --splits;
      if (splits > 1) {
        p.splits = splits;
        p.k_per = ((K + splits - 1) / splits + 63) & ~63;
        const bool fold = db && (size_t)splits * M * (Nn + 1) <= TrainWs::kPartialFloats && o.folds_bias_grad(p);
        if (fold) p.bias_grad_partial = sc_partial() + (size_t)splits * M * Nn;
        o.gemm(p, splits);
        KSplitReduce red{sc_partial(), C, scm, scn, M, Nn, splits, acc};
        if (fold) { red.bg_partial = p.bias_grad_partial; red.bg_out = db; }
        o.run(red, cdiv((long long)M * Nn, 256), 1, 1, 256);
        return fold;
      }
    }
    o.gemm(p, batch);
    return false;
  }
  // out[rows, C] = act(x[rows, K] W^T + b), W = checkpoint tensor wi [C, K] (+ bias wi+1)
  void layer(const float* x, int K, int wi, int C, float* out, long long rows, int relu) {
    gemm(x, K, 1, W[wi], 1, K, out, C, 1, (int)rows, C, K, W[wi + 1], relu, 0);
  }
  // act(x W^T + b) of checkpoint layer wi, max-pooled over the N points of each of the S sets, with the arg-max.  The launcher
  // may fuse the pooling into the GEMM (gemm_colmax: the [S N, C] activations are never stored); otherwise layer + colmax.
  void layer_max(const float* x, int K, int wi, int C, float* vmax, int* arg, int S, int relu) {
    GemmP p{};
    p.A = x; p.sam = K; p.sak = 1; p.B = W[wi]; p.sbk = 1; p.sbn = K; p.C = nullptr; p.scm = C; p.scn = 1; p.bias = W[wi + 1];
    p.M = (int)((long long)S * N); p.N = C; p.K = K; p.relu = relu; p.splits = 1; p.k_per = K; p.partial = w.partial; p.f16 = gemm_f16;
    if (o.gemm_colmax(p, N, vmax, arg, S, TrainWs::kPartialFloats)) return;
    layer(x, K, wi, C, w.zbuf, (long long)S * N, relu);
    colmax(w.zbuf, vmax, arg, S, C);
  }
  void colsum(const float* d, long long rows, int C, long long ld, float* out, int acc) {
    if (rows > 256) {  // two stages; enough chunks to fill the device, few enough for a short second stage
      long long chunks = rows >= 16384 ? 256 : (rows >= 2048 ? 64 : 16);
      while (chunks > 1 && (size_t)chunks * C > TrainWs::kCsFloats) chunks >>= 1;
      const long long per = (rows + chunks - 1) / chunks;
      chunks = (rows + per - 1) / per;
      o.run(KColSum{d, sc_cs_partial(), rows, C, per, 0, ld}, cdiv(C, 256), (unsigned)chunks, 1, 256);
      o.run(KColSum{sc_cs_partial(), out, chunks, C, chunks, acc, C}, cdiv(C, 256), 1, 1, 256);
    } else {
      o.run(KColSum{d, out, rows, C, rows, acc, ld}, cdiv(C, 256), 1, 1, 256);
    }
  }
  // out[g, c] = sum over the rows of group g (groups of `per_group` consecutive rows), two stages
  void group_colsum(const float* d, int groups, long long per_group, int C, float* out) {
    int sub = 16;
    while (sub > 1 && ((size_t)groups * sub * C > TrainWs::kCsFloats || per_group % sub != 0)) sub >>= 1;
    if (sub == 1) {
      o.run(KColSum{d, out, (long long)groups * per_group, C, per_group, 0, C}, cdiv(C, 256), groups, 1, 256);
      return;
    }
    o.run(KColSum{d, sc_cs_partial(), (long long)groups * per_group, C, per_group / sub, 0, C}, cdiv(C, 256), groups * sub, 1, 256);
    o.run(KColSum{sc_cs_partial(), out, (long long)groups * sub, C, sub, 0, C}, cdiv(C, 256), groups, 1, 256);
  }
  // column max + arg-max over the N points of each of S sets (first index wins ties), two stages when the scratch allows
  void colmax(const float* z, float* vmax, int* arg, int S, int C) {
    int chunks = 16;
    while (chunks > 1 && ((size_t)S * chunks * C > TrainWs::kPartialFloats / 2 || N % chunks != 0)) chunks >>= 1;
    if (chunks == 1) {
      o.run(KColMaxArg{z, vmax, arg, N, C}, cdiv(C, 256), S, 1, C < 256 ? 64 : 256);
      return;
    }
    float* pv = sc_partial();
    int* pi = reinterpret_cast<int*>(sc_partial() + TrainWs::kPartialFloats / 2);
    const unsigned nt = C < 256 ? 64 : 256;
    o.run(KColMaxArgPart{z, pv, pi, N, C, N / chunks}, cdiv(C, nt), chunks, S, nt);
    o.run(KColMaxArgMerge{pv, pi, vmax, arg, chunks, C}, cdiv(C, nt), S, 1, nt);
  }
  // sparse backward of a max-pooled layer: dx[(s, n), :] from d(max) [S, C] (KMaxBwdIndex + KMaxBwdGather)
  // The inverse arg-max map depends on the forward pass only, so it is built there, on the side lane, right after the layer's
  // arg-max exists (the main lane goes on with the next layers; the first join after it -- the heads' -- covers it): three
  // rank + range launch pairs leave the backward's critical path.  One slot per max-pooled layer.
  int mb_slot(const int* arg) const { return arg == w.sarg ? 0 : (arg == w.farg ? 1 : 2); }
  void build_max_index(const int* arg, int S, int C) {
    const int i = mb_slot(arg);
    begin_side();
    o.run(KMaxBwdRank{arg, w.mb_list[i], w.mb_key[i], C}, cdiv(C, 128), S, 1, 128);
    o.run(KMaxBwdRange{w.mb_key[i], w.mb_start[i], w.mb_cnt[i], N, C}, cdiv(N, 128), S, 1, 128);
    end_side();
  }
  // `act`: the layer's input activation [S N, K]; dx is zeroed where it is not positive (the ReLU below the layer, folded in)
  void max_bwd_dx(const float* dmax, const float* relu_max, const float* Wt, const int* arg, float* dx, int S, int C, int K,
                  const float* act) {
    const int i = mb_slot(arg);
    const int kq = K / 4, nt = 128;
    const int ppb = (kq < nt && nt % kq == 0) ? nt / kq : 1;  // points per block
    o.run(KMaxBwdGather{dmax, relu_max, Wt, w.mb_start[i], w.mb_cnt[i], w.mb_list[i], dx, act, N, C, K, ppb}, ppb > 1 ? 1 : cdiv(kq, nt),
          cdiv(N, ppb), S, nt);
  }
  // backward of y = x W^T + b for checkpoint tensor wi viewed as [C, ldw] with the K input columns at woff:
  // dW += dy^T x, db += colsum(dy) (when with_bias), dx (=|+=) dy W.
  void lin_bwd(int wi, const float* x, int K, const float* dy, int C, long long ldy, long long rows, float* dx, int dx_acc,
               int ldw = 0, int woff = 0, bool with_bias = true) {
    if (ldw == 0) ldw = K;
    // the parameter gradients (dW, db) and the data gradient (dx) only share their inputs: where the side lane is free (the
    // encoder part of the backward, split_lin_bwd) the former are issued on it next to the latter and joined right here, so
    // nothing outside this function sees a difference
    const bool sub = split_lin_bwd && lane == 0 && dx != nullptr;
    if (sub) begin_side();
    const bool folded = gemm(dy, 1, ldy, x, K, 1, w.G[wi] + woff, ldw, 1, C, K, (int)rows, nullptr, 0, 1, 1, 0, 0, 0, 0,
                             with_bias ? w.G[wi + 1] : nullptr);
    if (with_bias && !folded) colsum(dy, rows, C, ldy, w.G[wi + 1], 1);
    if (sub) end_side();
    if (dx) gemm(dy, ldy, 1, W[wi] + woff, ldw, 1, dx, K, 1, (int)rows, K, C, nullptr, 0, dx_acc);
    if (sub) o.join();
  }
  // wsum[b, c] = sum_p wp[p] u1[b, p, c]: chunked over the points, then the chunks of an object in order
  void rot_wsum(const float* u1, const float* wp, float* wsum, int B, int P) {
    int chunks = 32;
    while (chunks > 1 && ((size_t)B * chunks * 256 > TrainWs::kCsFloats || P % chunks != 0)) chunks >>= 1;
    if (chunks == 1) { o.run(KRotWsum{u1, wp, wsum, P}, B, 1, 1, 256); return; }
    o.run(KRotWsumPart{u1, wp, sc_cs_partial(), P, P / chunks}, 1, chunks, B, 256);
    o.run(KColSum{sc_cs_partial(), wsum, (long long)B * chunks, 256, chunks, 0, 256}, 1, B, 1, 256);
  }
  void relu_mask(float* d, const float* act, long long n) { o.run(KReluMask{d, act, n}, cdiv(n, 256), 1, 1, 256); }
  static int gn_chunks(int P) {  // at least 4 points per chunk
    int c = TrainWs::kGnChunks;
    while (c > 1 && P < 4 * c) c >>= 1;
    return c;
  }
  void gn_fwd(const float* y, float* st, const float* ga, const float* be, float* u, int B, int P) {
    const int chunks = gn_chunks(P), per = (P + chunks - 1) / chunks;
    o.run(KGnStatsPart{y, sc_gn_part(), P, chunks, per}, B, chunks, 1, 32);
    o.run(KGnStats{sc_gn_part(), st, P, chunks}, B, 1, 1, 32);
    const long long n = (long long)B * P * 256;
    o.run(KGnGeluFwd{y, st, ga, be, u, P, n}, cdiv(n / 4, 256), 1, 1, 256);
  }
  // du -> dy in place; gamma / beta gradients accumulated into G[gi], G[gi + 1]
  void gn_bwd(float* du, const float* y, const float* st, int gi, int B, int P) {
    const int chunks = gn_chunks(P), per = (P + chunks - 1) / chunks;
    o.run(KGnBwdPart{du, y, st, W[gi], W[gi + 1], sc_gn_part(), P, chunks, per}, B, chunks, 1, 32);
    o.run(KGnBwdSums{sc_gn_part(), sc_gn_m(), sc_gnp_g(), sc_gnp_b(), P, chunks}, B, 1, 1, 32 * 18);
    colsum(sc_gnp_g(), B, 256, 256, w.G[gi], 1);
    colsum(sc_gnp_b(), B, 256, 256, w.G[gi + 1], 1);
    const long long n = (long long)B * P * 256;
    o.run(KGnBwdApply{du, y, st, W[gi], W[gi + 1], sc_gn_m(), P, n}, cdiv(n / 4, 256), 1, 1, 256);
  }

  // T-Net forward (pointnets/pointnet.py:24-41, 57-78)
  void tnet_fwd(const float* x, int Kin, int wb, int k, float* c64, float* c128, float* vmax, int* arg, float* fc1, float* fc2,
                float* tout, int S) {
    const long long R = (long long)S * N;
    layer(x, Kin, wb + T_CONV1, 64, c64, R, 1);
    layer(c64, 64, wb + T_CONV2, 128, c128, R, 1);
    layer_max(c128, 128, wb + T_CONV3, 1024, vmax, arg, S, 1);
    build_max_index(arg, S, 1024);
    layer(vmax, 1024, wb + T_FC1, 512, fc1, S, 1);
    layer(fc1, 512, wb + T_FC2, 256, fc2, S, 1);
    layer(fc2, 256, wb + T_FC3, k * k, tout, S, 0);
    o.run(KAddEye{tout, k}, cdiv(k, 64), S, 1, 64);
  }
  // T-Net backward from d(T) [S, k*k]; d(x) is accumulated into dx_out when given (oracle: tnet_bwd)
  void tnet_bwd(const float* x, int Kin, int wb, int k, const float* c64, const float* c128, const float* vmax, const int* arg,
                const float* fc1, const float* fc2, const float* dT, float* dx_out, int S) {
    const long long R = (long long)S * N;
    lin_bwd(wb + T_FC3, fc2, 256, dT, k * k, k * k, S, w.dfc2, 0);
    relu_mask(w.dfc2, fc2, (long long)S * 256);
    lin_bwd(wb + T_FC2, fc1, 512, w.dfc2, 256, 256, S, w.dfc1, 0);
    relu_mask(w.dfc1, fc1, (long long)S * 512);
    lin_bwd(wb + T_FC1, vmax, 1024, w.dfc1, 512, 512, S, w.dmax, 0);
    // the layer's parameter gradients next to its data gradient (both read d(max); joined here)
    const bool sub = split_lin_bwd && lane == 0;
    if (sub) begin_side();
    o.run(KMaxBwdDw{w.dmax, vmax, c128, arg, w.G[wb + T_CONV3], w.G[wb + T_CONV3 + 1], S, N, 1024, 128}, cdiv(128, 128), 1024, 1, 128);
    if (sub) end_side();
    max_bwd_dx(w.dmax, vmax, W[wb + T_CONV3], arg, w.d128, S, 1024, 128, c128);  // incl. the ReLU mask of conv2's output
    if (sub) o.join();
    lin_bwd(wb + T_CONV2, c64, 64, w.d128, 128, 128, R, w.d64, 0);
    relu_mask(w.d64, c64, R * 64);
    lin_bwd(wb + T_CONV1, x, Kin, w.d64, 64, 64, R, dx_out, 1);
  }

  // one rotation head (heads/conv_out_per_rot_head.py:126-140) with the layer-0 split; h = 0: x, 1: y
  void rot_fwd(int h, int B) {
    const int S = 2 * B, P = 2 * N;
    const long long R = (long long)S * N;
    const int rb = h ? W_ROT_Y : W_ROT_X;
    gemm(w.g, 1024, 1, W[rb + R_L0], 1, 1088, w.cset[h], 256, 1, S, 256, 1024, W[rb + R_L0 + 1], 0, 0);
    gemm(w.pf, 64, 1, W[rb + R_L0] + 1024, 1, 1088, w.ry0[h], 256, 1, N, 256, 64, w.cset[h], 0, 0, S, (long long)N * 64, 0,
         (long long)N * 256, 256);
    gn_fwd(w.ry0[h], w.rst0[h], W[rb + R_GN0], W[rb + R_GN0 + 1], w.ru0[h], B, P);
    layer(w.ru0[h], 256, rb + R_L3, 256, w.ry1[h], R, 0);
    gn_fwd(w.ry1[h], w.rst1[h], W[rb + R_GN1], W[rb + R_GN1 + 1], w.ru1[h], B, P);
    rot_wsum(w.ru1[h], W[rb + R_CONVP], w.wsum[h], B, P);
    colsum(W[rb + R_CONVP], P, 1, 1, w.swp + h, 0);  // sum_p wp[p], also read by the backward of this step
    o.run(KRotOut{w.wsum[h], W[rb + R_NECK], W[rb + R_NECK + 1], w.swp, W[rb + R_CONVP + 1], w.r6, P, h}, B, 1, 1, 32);
  }
  // backward of one rotation head down to its own weights' gradients and to du0 / dcset (the gradients at the layer-0 output,
  // per point and summed per set); what it adds to the shared dg / dpf is rot_bwd_shared
  void rot_bwd(int h, int B) {
    const int S = 2 * B, P = 2 * N;
    const long long R = (long long)S * N;
    const int rb = h ? W_ROT_Y : W_ROT_X;
    const float* wp = W[rb + R_CONVP];
    o.run(KRotTailBwd{w.d_r6, w.wsum[h], W[rb + R_NECK], w.swp, w.e[h], w.G[rb + R_NECK], w.G[rb + R_NECK + 1], w.G[rb + R_CONVP + 1], B, P, h},
          1, 1, 1, 256);
    const long long n = (long long)B * P * 256;
    o.run(KGnGeluFwd{w.ry1[h], w.rst1[h], W[rb + R_GN1], W[rb + R_GN1 + 1], w.ru1[h], P, n}, cdiv(n / 4, 256), 1, 1, 256);  // recompute u1
    if ((size_t)B * P * 8 <= TrainWs::kPartialFloats) {
      o.run(KRotDwpPart{w.ru1[h], w.e[h], sc_partial(), P}, cdiv(8 * P, 256), B, 1, 256);
      o.run(KRotDwpSum{sc_partial(), w.d_r6, W[rb + R_NECK + 1], w.G[rb + R_CONVP], B, P, h}, cdiv(P, 128), 1, 1, 128);
    } else {
      o.run(KRotDwp{w.ru1[h], w.e[h], w.d_r6, W[rb + R_NECK + 1], w.G[rb + R_CONVP], B, P, h}, cdiv(P, 128), 1, 1, 128);
    }
    o.run(KRotDu1{wp, w.e[h], w.du[h], P, n}, cdiv(n, 256), 1, 1, 256);
    gn_bwd(w.du[h], w.ry1[h], w.rst1[h], rb + R_GN1, B, P);
    lin_bwd(rb + R_L3, w.ru0[h], 256, w.du[h], 256, 256, R, w.du0[h], 0);
    gn_bwd(w.du0[h], w.ry0[h], w.rst0[h], rb + R_GN0, B, P);
    // layer 0: point-feature columns per point, global-feature columns once per set
    group_colsum(w.du0[h], S, N, 256, w.dcset[h]);  // dcset[s] = sum over the set's points
    gemm(w.dcset[h], 1, 256, w.g, 1024, 1, w.G[rb + R_L0], 1088, 1, 256, 1024, S, nullptr, 0, 1);
    gemm(w.du0[h], 1, 256, w.pf, 64, 1, w.G[rb + R_L0] + 1024, 1088, 1, 256, 64, (int)R, nullptr, 0, 1);
    colsum(w.dcset[h], S, 256, 256, w.G[rb + R_L0 + 1], 1);
  }
  void rot_bwd_shared(int h, int B) {
    const int S = 2 * B;
    const long long R = (long long)S * N;
    const int rb = h ? W_ROT_Y : W_ROT_X;
    gemm(w.dcset[h], 256, 1, W[rb + R_L0], 1088, 1, w.dg, 1024, 1, S, 1024, 256, nullptr, 0, 1);
    gemm(w.du0[h], 256, 1, W[rb + R_L0] + 1024, 1088, 1, w.dpf, 64, 1, (int)R, 64, 256, nullptr, 0, 1);
  }

  void forward(const TrainIn& in) {
    const int B = in.B, S = 2 * B, P = 2 * N;
    const long long R = (long long)S * N;
    gemm_f16 = 1;
    if (in.x_pm) o.run(KInterleave{in.x_pm, in.tfd_pm, w.q, N}, cdiv(3 * N, 256), B, 1, 256);
    else o.run(KUpdatePoints{in.pcl, in.kps, in.pose, in.scale, w.q, N}, cdiv(N, 256), B, 1, 256);
    // encoder (pointnets/pointnet.py:97-116), all 2B sets at once
    tnet_fwd(w.q, 3, W_STN, 3, w.s64, w.s128, w.smax, w.sarg, w.sfc1, w.sfc2, w.t3, S);
    gemm(w.q, 3, 1, w.t3, 3, 1, w.qp, 3, 1, N, 3, 3, nullptr, 0, 0, S, (long long)N * 3, 9, (long long)N * 3);  // q' = q . T3
    layer(w.qp, 3, W_CONV1, 64, w.h1, R, 1);
    tnet_fwd(w.h1, 64, W_FSTN, 64, w.f64, w.f128, w.fmax, w.farg, w.ffc1, w.ffc2, w.t64, S);
    gemm(w.h1, 64, 1, w.t64, 64, 1, w.pf, 64, 1, N, 64, 64, nullptr, 0, 0, S, (long long)N * 64, 4096, (long long)N * 64);  // pf = h1 . T64
    begin_side();  // max over the points of the 64 point features (ts-head input): next to the trunk's wide layers; its
    colmax(w.pf, w.pfmax, w.pfarg, S, 64);  // consumers are the side lane itself (ts head) and the backward
    end_side();
    layer(w.pf, 64, W_CONV2, 128, w.a128, R, 1);
    layer(w.a128, 128, W_CONV3, 512, w.a512, R, 1);
    layer_max(w.a512, 512, W_CONV4, 1024, w.g, w.garg, S, 0);  // no ReLU after conv4 (pointnet.py:114)
    build_max_index(w.garg, S, 1024);
    // translation / size head (heads/fc_trans_size_head.py:61-70): its own lane; the y rotation head: the side lane
    o.join();  // the point-feature max (side lane) feeds the ts head (another lane)
    begin_side(2);
    o.run(KTsGather{w.g, w.pfmax, in.scale, w.ts_in}, cdiv(1091, 256), B, 1, 256);
    layer(w.ts_in, 1091, W_TS + S_L0, 256, w.ts_y0, B, 0);
    gn_fwd(w.ts_y0, w.ts_st0, W[W_TS + S_GN0], W[W_TS + S_GN0 + 1], w.ts_u0, B, 1);
    layer(w.ts_u0, 256, W_TS + S_L3, 256, w.ts_y1, B, 0);
    gn_fwd(w.ts_y1, w.ts_st1, W[W_TS + S_GN1], W[W_TS + S_GN1 + 1], w.ts_u1, B, 1);
    gemm(w.ts_u1, 256, 1, W[W_TS + S_FCT], 1, 256, w.dts, 6, 1, B, 3, 256, W[W_TS + S_FCT + 1], 0, 0);
    gemm(w.ts_u1, 256, 1, W[W_TS + S_FCS], 1, 256, w.dts + 3, 6, 1, B, 3, 256, W[W_TS + S_FCS + 1], 0, 0);
    end_side();
    begin_side(1);
    rot_fwd(1, B);
    end_side();
    rot_fwd(0, B);
    o.join();
    o.run(KPoseFwd{w.r6, w.dts, in.pose, in.scale, in.K, in.pose_out, in.scale_out, B}, cdiv(B, 64), 1, 1, 64);
  }

  void loss(const TrainIn& in) {
    int chunks = TrainWs::kLossChunks;
    while (chunks > 1 && N % chunks != 0) chunks >>= 1;
    // the search over the symmetric copies in chunks of 8 rotations (cs_partial is free between forward and backward)
    int sel_per = 8;
    while ((size_t)in.B * cdiv(in.n_rots, sel_per) * 2 > TrainWs::kCsFloats) sel_per *= 2;
    const int sel_chunks = (int)cdiv(in.n_rots, sel_per);
    if (sel_chunks > 0)
      o.run(KLossSelPart{in.pose_out, in.gt_pose, w.sym_rots, w.is_sym, w.cs_partial, in.B, in.n_rots, sel_chunks, sel_per},
            cdiv(sel_chunks, 32), in.B, 1, 32);
    o.run(KLossSel{in.pose_out, in.gt_pose, w.sym_rots, w.is_sym, w.cs_partial, w.loss_gs, in.B, in.n_rots, sel_chunks},
          cdiv(in.B, 32), 1, 1, 32);
    o.run(KLossPm{in.pose_out, in.scale_out, in.gt_scale, in.kps, w.loss_gs, w.gn_part, in.B, N, chunks, N / chunks, in.w_pm},
          cdiv(chunks, 32), in.B, 1, 32);
    o.run(KLoss{in.pose_out, in.scale_out, in.gt_pose, in.gt_scale, w.gn_part, w.is_sym, w.lossp, w.dpose, in.B, N, chunks,
                in.n_sym, in.n_nosym, in.w_pm, in.w_rot, in.w_trans, in.w_scale}, cdiv(in.B, 32), 1, 1, 32);
    o.run(KLossSum{w.lossp, w.losses, in.B}, 1, 1, 1, 32);
  }

  void backward(const TrainIn& in) {
    const int B = in.B, S = 2 * B, P = 2 * N;
    const long long R = (long long)S * N;
    gemm_f16 = 0;
    o.zero(w.G[0], w.grad_floats * sizeof(float));
    o.zero(w.dg, (size_t)S * 1024 * sizeof(float));
    o.zero(w.dpfmax, (size_t)S * 64 * sizeof(float));
    o.zero(w.dpf, (size_t)R * 64 * sizeof(float));
    o.run(KPoseBwd{w.dpose, w.r6, w.dts, in.pose, in.K, w.d_r6, w.d_dts, B}, cdiv(B, 64), 1, 1, 64);
    // ---- the ts head on its own lane (its gradient w.r.t. the global feature is added after the join), the y rotation head on
    //      the side lane
    begin_side(2);
    lin_bwd(W_TS + S_FCT, w.ts_u1, 256, w.d_dts, 3, 6, B, w.tsd_u, 0);
    lin_bwd(W_TS + S_FCS, w.ts_u1, 256, w.d_dts + 3, 3, 6, B, w.tsd_u, 1);
    gn_bwd(w.tsd_u, w.ts_y1, w.ts_st1, W_TS + S_GN1, B, 1);
    lin_bwd(W_TS + S_L3, w.ts_u0, 256, w.tsd_u, 256, 256, B, w.tsd_u0, 0);
    gn_bwd(w.tsd_u0, w.ts_y0, w.ts_st0, W_TS + S_GN0, B, 1);
    lin_bwd(W_TS + S_L0, w.ts_in, 1091, w.tsd_u0, 256, 256, B, w.ts_din, 0);
    end_side();
    begin_side(1);
    rot_bwd(1, B);
    end_side();
    // ---- main lane: the x rotation head, then (after the join) the shared-gradient products of both heads in the order x, y
    rot_bwd(0, B);
    rot_bwd_shared(0, B);
    o.join();
    rot_bwd_shared(1, B);
    // ---- encoder
    split_lin_bwd = true;
    o.run(KTsScatter{w.ts_din, w.dg, w.dpfmax}, cdiv(1088, 256), B, 1, 256);  // += on top of the rotation heads' share
    o.run(KScatterMax{w.dpfmax, w.pfarg, w.dpf, N, 64}, 1, S, 1, 64);
    begin_side();
    o.run(KMaxBwdDw{w.dg, nullptr, w.a512, w.garg, w.G[W_CONV4], w.G[W_CONV4 + 1], S, N, 1024, 512}, cdiv(512, 128), 1024, 1, 128);
    end_side();
    max_bwd_dx(w.dg, nullptr, W[W_CONV4], w.garg, w.d512, S, 1024, 512, w.a512);  // incl. the ReLU mask of conv3's output
    o.join();
    lin_bwd(W_CONV3, w.a128, 128, w.d512, 512, 512, R, w.d128, 0);
    relu_mask(w.d128, w.a128, R * 128);
    lin_bwd(W_CONV2, w.pf, 64, w.d128, 128, 128, R, w.dpf, 1);
    // pf = h1 . T64 per set:  dh1 = dpf . T64^T,  dT64 = h1^T . dpf
    begin_side();
    gemm(w.h1, 1, 64, w.dpf, 64, 1, w.dt64, 64, 1, 64, 64, N, nullptr, 0, 0, S, (long long)N * 64, (long long)N * 64, 4096);
    end_side();
    gemm(w.dpf, 64, 1, w.t64, 1, 64, w.dh1, 64, 1, N, 64, 64, nullptr, 0, 0, S, (long long)N * 64, 4096, (long long)N * 64);
    o.join();
    tnet_bwd(w.h1, 64, W_FSTN, 64, w.f64, w.f128, w.fmax, w.farg, w.ffc1, w.ffc2, w.dt64, w.dh1, S);
    relu_mask(w.dh1, w.h1, R * 64);
    lin_bwd(W_CONV1, w.qp, 3, w.dh1, 64, 64, R, w.dqp, 0);
    {  // dT3 = q^T . dq' per set (3 x 3 outputs over N points: chunked sums, then the chunks of a set in order)
      int chunks = 32;
      while (chunks > 1 && ((size_t)S * chunks * 9 > TrainWs::kCsFloats || N % chunks != 0)) chunks >>= 1;
      o.run(KSet3x3Part{w.q, w.dqp, w.cs_partial, N, N / chunks}, 1, chunks, S, 32);
      o.run(KColSum{w.cs_partial, w.dt3, (long long)S * chunks, 9, chunks, 0, 9}, 1, S, 1, 32);
    }
    tnet_bwd(w.q, 3, W_STN, 3, w.s64, w.s128, w.smax, w.sarg, w.sfc1, w.sfc2, w.dt3, nullptr, S);
    split_lin_bwd = false;
  }
};

}  // namespace catre_train

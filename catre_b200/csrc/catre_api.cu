// libcatre_b200.so -- C ABI (include/catre_b200.h) and the host-side driver of the kernel chain.
// One engine per device; all work is enqueued on the caller's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/catre_b200.h"
#include "simt_kernels.cuh"
#include "tc_kernels.cuh"
#include "tc_fused.cuh"
#include "fc_chain.cuh"
#include "cloud_kernels.cuh"
#include "metrics_kernels.cuh"
#include "train_chain.cuh"
#include "train_gemm_tc.cuh"
#include "optim_kernels.cuh"

using namespace catre;

namespace {

thread_local std::string g_create_error;

struct WeightSpec {
  const char* name;
  int ndim;
  long long shape[3];  // -1 in shape[1] of conv_p = n_obs + n_prior
};

// the 74 checkpoint tensors in checkpoint order
const WeightSpec kWeights[] = {
    {"pcl_net.stn.conv1.weight", 3, {64, 3, 1}},      {"pcl_net.stn.conv1.bias", 1, {64}},
    {"pcl_net.stn.conv2.weight", 3, {128, 64, 1}},    {"pcl_net.stn.conv2.bias", 1, {128}},
    {"pcl_net.stn.conv3.weight", 3, {1024, 128, 1}},  {"pcl_net.stn.conv3.bias", 1, {1024}},
    {"pcl_net.stn.fc1.weight", 2, {512, 1024}},       {"pcl_net.stn.fc1.bias", 1, {512}},
    {"pcl_net.stn.fc2.weight", 2, {256, 512}},        {"pcl_net.stn.fc2.bias", 1, {256}},
    {"pcl_net.stn.fc3.weight", 2, {9, 256}},          {"pcl_net.stn.fc3.bias", 1, {9}},
    {"pcl_net.conv1.weight", 3, {64, 3, 1}},          {"pcl_net.conv1.bias", 1, {64}},
    {"pcl_net.conv2.weight", 3, {128, 64, 1}},        {"pcl_net.conv2.bias", 1, {128}},
    {"pcl_net.conv3.weight", 3, {512, 128, 1}},       {"pcl_net.conv3.bias", 1, {512}},
    {"pcl_net.conv4.weight", 3, {1024, 512, 1}},      {"pcl_net.conv4.bias", 1, {1024}},
    {"pcl_net.fstn.conv1.weight", 3, {64, 64, 1}},    {"pcl_net.fstn.conv1.bias", 1, {64}},
    {"pcl_net.fstn.conv2.weight", 3, {128, 64, 1}},   {"pcl_net.fstn.conv2.bias", 1, {128}},
    {"pcl_net.fstn.conv3.weight", 3, {1024, 128, 1}}, {"pcl_net.fstn.conv3.bias", 1, {1024}},
    {"pcl_net.fstn.fc1.weight", 2, {512, 1024}},      {"pcl_net.fstn.fc1.bias", 1, {512}},
    {"pcl_net.fstn.fc2.weight", 2, {256, 512}},       {"pcl_net.fstn.fc2.bias", 1, {256}},
    {"pcl_net.fstn.fc3.weight", 2, {4096, 256}},      {"pcl_net.fstn.fc3.bias", 1, {4096}},
    {"rot_head.rot_head_x.norm.weight", 1, {256}},    {"rot_head.rot_head_x.norm.bias", 1, {256}},
    {"rot_head.rot_head_x.layers.0.weight", 3, {256, 1088, 1}}, {"rot_head.rot_head_x.layers.0.bias", 1, {256}},
    {"rot_head.rot_head_x.layers.1.weight", 1, {256}}, {"rot_head.rot_head_x.layers.1.bias", 1, {256}},
    {"rot_head.rot_head_x.layers.3.weight", 3, {256, 256, 1}}, {"rot_head.rot_head_x.layers.3.bias", 1, {256}},
    {"rot_head.rot_head_x.layers.4.weight", 1, {256}}, {"rot_head.rot_head_x.layers.4.bias", 1, {256}},
    {"rot_head.rot_head_x.neck.0.weight", 3, {3, 256, 1}}, {"rot_head.rot_head_x.neck.0.bias", 1, {3}},
    {"rot_head.rot_head_x.conv_p.weight", 3, {1, -1, 1}}, {"rot_head.rot_head_x.conv_p.bias", 1, {1}},
    {"rot_head.rot_head_y.norm.weight", 1, {256}},    {"rot_head.rot_head_y.norm.bias", 1, {256}},
    {"rot_head.rot_head_y.layers.0.weight", 3, {256, 1088, 1}}, {"rot_head.rot_head_y.layers.0.bias", 1, {256}},
    {"rot_head.rot_head_y.layers.1.weight", 1, {256}}, {"rot_head.rot_head_y.layers.1.bias", 1, {256}},
    {"rot_head.rot_head_y.layers.3.weight", 3, {256, 256, 1}}, {"rot_head.rot_head_y.layers.3.bias", 1, {256}},
    {"rot_head.rot_head_y.layers.4.weight", 1, {256}}, {"rot_head.rot_head_y.layers.4.bias", 1, {256}},
    {"rot_head.rot_head_y.neck.0.weight", 3, {3, 256, 1}}, {"rot_head.rot_head_y.neck.0.bias", 1, {3}},
    {"rot_head.rot_head_y.conv_p.weight", 3, {1, -1, 1}}, {"rot_head.rot_head_y.conv_p.bias", 1, {1}},
    {"ts_head.norm.weight", 1, {256}},                {"ts_head.norm.bias", 1, {256}},
    {"ts_head.linears.0.weight", 2, {256, 1091}},     {"ts_head.linears.0.bias", 1, {256}},
    {"ts_head.linears.1.weight", 1, {256}},           {"ts_head.linears.1.bias", 1, {256}},
    {"ts_head.linears.3.weight", 2, {256, 256}},      {"ts_head.linears.3.bias", 1, {256}},
    {"ts_head.linears.4.weight", 1, {256}},           {"ts_head.linears.4.bias", 1, {256}},
    {"ts_head.fc_t.weight", 2, {3, 256}},             {"ts_head.fc_t.bias", 1, {3}},
    {"ts_head.fc_s.weight", 2, {3, 256}},             {"ts_head.fc_s.bias", 1, {3}},
};
constexpr int kNumWeights = sizeof(kWeights) / sizeof(kWeights[0]);

// kernel groups for launch accounting / per-group device timing
enum Grp {
  G_UPDATE_POINTS, G_FRONT3, G_STN_CONV2, G_STN_CONV3_MAX, G_TNET_FC, G_FSTN_CONV1, G_FSTN_CONV2,
  G_FSTN_CONV3_MAX, G_FEAT_TRANSFORM, G_CONV2, G_CONV3, G_CONV4_MAX, G_ROT_GFEAT, G_ROT_LAYER0, G_GN_FINALIZE,
  G_ROT_LAYER1, G_ROT_TAIL, G_TS_POSE, G_ROT_FUSED, G_NUM
};
const char* kGrpNames[G_NUM] = {
    "update_points", "front3", "stn_conv2", "stn_conv3_max", "tnet_fc", "fstn_conv1", "fstn_conv2",
    "fstn_conv3_max", "feat_transform", "conv2", "conv3", "conv4_max", "rot_gfeat", "rot_layer0", "gn_finalize",
    "rot_layer1", "rot_tail", "ts_pose", "rot_fused"};

}  // namespace

struct catre_engine {
  catre_cfg cfg{};
  std::string err;
  int N = 0;          // observed points per object (n_obs)
  int Np = 0;         // prior points per object (n_prior); rows per object P = N + Np
  int maxB = 0;
  bool packed = false;
  std::map<std::string, std::vector<float>> hw;  // host copies of the checkpoint tensors
  std::vector<void*> dev_allocs;

  // ---- device weights (fp32)
  std::map<std::string, float*> dw;
  float *stn_fc3_bI = nullptr, *fstn_fc3_bI = nullptr;
  float *rot_w0g = nullptr, *rot_b0 = nullptr, *rot_w0p = nullptr;
  float *rot_gn0_g = nullptr, *rot_gn0_b = nullptr, *rot_gn1_g = nullptr, *rot_gn1_b = nullptr;
  float *rot_b1 = nullptr, *neck_w = nullptr, *neck_b = nullptr, *wp = nullptr, *convp_b = nullptr;
  float *ts_w0t = nullptr, *ts_w1t = nullptr, *ts_w0g = nullptr;
  // bf16 hi/lo weight copies + tensor maps (tensor-core modes); "MA" maps have 128-row boxes (M side of
  // the MMA), "NB" maps have BN-row boxes (N side)
  TcPair tw_stn_c2, tw_stn_c3, tw_fstn_c1, tw_fstn_c2, tw_fstn_c3, tw_conv2, tw_conv3, tw_conv4, tw_rot0;
  TcPair tw_rot1s;            // both heads' layers.3 weights stacked [512, 256] (fused rot kernel)
  cudaStream_t side = nullptr;           // the ts head runs here, underneath the rot-head kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t side2 = nullptr;          // second side lane of the training chain (the ts head's)
  cudaEvent_t ev_join2 = nullptr;
  float* dts = nullptr;                  // [B, 6] raw ts-head outputs
  CUtensorMap tw_rot0_nb[2];  // rot layer-0 point-feature weights [512, 64] as an N-side operand (256-row boxes)
  float *fstn_fc3_wT = nullptr;  // fstn.fc3 rows permuted so the FC emits T64^T (row j = output channel j of pf = h1 . T64)
  int num_sms = 148;

  // ---- workspace
  float *q = nullptr, *h64a = nullptr, *h64b = nullptr, *h128 = nullptr, *h512 = nullptr, *a0 = nullptr, *a1 = nullptr;
  int *gmax_all = nullptr, *gmax_stn = nullptr, *gmax_fstn = nullptr, *gmax_g = nullptr, *gmax_pf = nullptr;
  float *ts0 = nullptr, *t3 = nullptr, *t64 = nullptr, *cset = nullptr;
  float *fc512 = nullptr, *fc256 = nullptr;  // fp32 intermediates of the tiled FC path (large batches)
  int fct_min_rows = 1 << 30;                // sets from which ALL FC layers run as tiled GEMMs (measured slower: off)
  int fct_fc3_min_rows = 16;                 // sets from which fstn.fc3 (256 -> 4096, 62 % of the fstn chain) leaves the chain
  // fused FC chains (fc_chain.cuh): weights packed [8 ranks][K][C/8] fp32; [0] = stn, [1] = fstn
  float *fcc_fc1[2] = {nullptr, nullptr}, *fcc_fc2[2] = {nullptr, nullptr}, *fcc_fc3[2] = {nullptr, nullptr};
  float *fcc_cset = nullptr, *fcc_ts0 = nullptr;
  int trunk_group = 0;  // objects per conv3 -> conv4 group launch (0: the whole batch at once); CATRE_TRUNK_GROUP
  int* rot_count = nullptr;   // [B][2] publication counters of the fused rot kernel (zeroed by gn_finalize_set_kernel)
  bool fused_tail = false;    // CATRE_ROT_TAIL=fused: the rot tail runs out of TMEM inside the fused rot kernel (no a1T); default: split
  bool debug_taps = false;    // CATRE_DEBUG_TAPS=1: keep optional intermediate copies for catre_debug_read (tests/test_stages_gpu.py)
  int fcc_ranks = 8;  // CTAs per FC-chain cluster (16 where the device co-schedules them), fixed per engine
  TcPair t64s;  // tensor-core modes: bf16 hi/lo of T64^T, [S*64, 64]
  float *stats0 = nullptr, *stats1 = nullptr, *gn0 = nullptr, *gn1 = nullptr, *rot_partial = nullptr;
  // bf16 hi/lo activations of the tensor-core path and their tensor maps
  TcPair x64, f64, a128, pf16, a512;               // .map_* = MA view (128-row boxes)
  CUtensorMap a128_nb[2], pf_nb[2], a512_nb[2];   // NB views (256-row boxes): [hi, lo]
  size_t ws_bytes = 0;
  // staging for catre_refine_host
  float *st_pcl = nullptr, *st_prior = nullptr, *st_pose = nullptr, *st_scale = nullptr, *st_K = nullptr;
  float *st_oposes = nullptr, *st_oscales = nullptr;
  int32_t* st_cls = nullptr;
  int st_prior_rows = 0;  // rows ([N,3] each) st_prior holds: max(max_batch, 16), so a small class table always fits
  static constexpr int kMaxHostIter = 16;

  // ---- training step: workspace grown on demand, weights refreshed device-to-device
  catre_train::TrainWs tws;
  char* tws_mem = nullptr;
  int tws_maxB = 0;
  std::map<std::string, bool> dw_dirty;  // tensors refreshed by catre_train_set_weight since the last pack
  float loss_w[4] = {1.0f, 1.0f, 1.0f, 1.0f};  // PM_LW, ROT_LW, TRANS_LW, SCALE_LW
  bool train_naive_gemm = false;         // CATRE_TRAIN_NAIVE_GEMM=1: one-thread-per-output GEMM (debug reference)
  // CUDA graphs of the training chain (default; CATRE_TRAIN_GRAPH=0 launches the chain kernel by kernel).  The chain's launch
  // geometry depends on (B, number of symmetric objects, number of symmetry rotations, loss weights, GEMM mode) only; its
  // inputs / outputs are engine-owned static buffers (tg_io) filled / drained by plain copies around the graph launch, so a
  // graph stays valid whatever tensors the caller passes.  The first step of a key runs kernel by kernel (it also configures
  // the kernels' attributes), the second is captured on `cap` (torch's default stream is the legacy stream, which cannot be
  // captured) and every later one is a single cudaGraphLaunch on the caller's stream.
  struct TrainGraphKey {
    int B, n_sym, n_rots, mode; float lw[4];
    bool operator<(const TrainGraphKey& o) const { return memcmp(this, &o, sizeof(*this)) < 0; }
  };
  struct TrainGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int seen = 0; };
  std::map<TrainGraphKey, TrainGraph> train_graphs;
  cudaStream_t cap = nullptr;
  float* tg_io = nullptr;                // [x_pm | tfd_pm | obj_kps | pose | scale | K | gt_pose | gt_scale | out_pose | out_scale]

  // ---- accounting
  int64_t launches = 0;
  bool prof_on = false;
  struct Ev { int grp; cudaEvent_t a, b; };
  std::vector<Ev> ev_pending;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[G_NUM] = {0};
  int64_t prof_n[G_NUM] = {0};
};

namespace {

int fail(catre_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CU_TRY(e, call)                                                                           \
  do {                                                                                            \
    cudaError_t _st = (call);                                                                     \
    if (_st != cudaSuccess)                                                                       \
      return fail(e, CATRE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), __FILE__, __LINE__); \
  } while (0)

template <typename T>
int dalloc(catre_engine* e, T** p, size_t n) {
  void* v = nullptr;
  size_t bytes = n * sizeof(T);
  if (bytes == 0) bytes = 16;
  CU_TRY(e, cudaMalloc(&v, bytes));
  CU_TRY(e, cudaMemset(v, 0, bytes));  // padded operand rows (beyond the live sets) must hold finite values
  e->dev_allocs.push_back(v);
  e->ws_bytes += bytes;
  *p = reinterpret_cast<T*>(v);
  return 0;
}

int upload(catre_engine* e, float** p, const std::vector<float>& h) {
  int rc = dalloc(e, p, h.size());
  if (rc) return rc;
  CU_TRY(e, cudaMemcpy(*p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// ---- launch bookkeeping ------------------------------------------------------------------------
struct Launch {
  catre_engine* e;
  cudaStream_t s;
  int grp;
  cudaEvent_t a = nullptr, b = nullptr;
  Launch(catre_engine* e_, cudaStream_t s_, int g) : e(e_), s(s_), grp(g) {
    e->launches++;
    if (e->prof_on) {
      auto get = [&]() {
        cudaEvent_t ev;
        if (!e->ev_pool.empty()) { ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
        else cudaEventCreate(&ev);
        return ev;
      };
      a = get(); b = get();
      cudaEventRecord(a, s);
    }
  }
  ~Launch() {
    if (e->prof_on) {
      cudaEventRecord(b, s);
      e->ev_pending.push_back({grp, a, b});
    }
  }
};

int check_launch(catre_engine* e, const char* what) {
  cudaError_t st = cudaPeekAtLastError();
  if (st != cudaSuccess) {
    cudaGetLastError();
    return fail(e, CATRE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(st));
  }
  return 0;
}

GemmP gemm_args(const float* A, int lda, const float* W, int K, int C, const float* bias, float* out, int ldo,
                long long R, int relu) {
  GemmP p{};
  p.A = A; p.lda = lda; p.W = W; p.wcs = K; p.w_set_stride = 0;
  p.bias = bias; p.rowvec = nullptr; p.ldrv = 0; p.out = out; p.ldo = ldo; p.gmax = nullptr; p.stats = nullptr;
  p.stats_ld = 0; p.stats_goff = 0;
  p.gn_scale = p.gn_shift = nullptr; p.ldgn = 0;
  p.R = (int)R; p.C = C; p.K = K; p.rows_per_set = 1; p.rows_per_obj = 1; p.relu = relu;
  return p;
}

template <int BN, int AMODE>
int run_gemm(catre_engine* e, cudaStream_t s, int grp, const GemmP& p) {
  dim3 grid((p.C + BN - 1) / BN, (p.R + 127) / 128);
  {
    Launch l(e, s, grp);
    launch_pdl(pw_gemm_kernel<BN, AMODE>, dim3(grid), dim3(256), (size_t)(0), s, p);
  }
  return check_launch(e, kGrpNames[grp]);
}

const float* W(catre_engine* e, const char* name) { return e->dw.at(name); }

// ---- tensor-core launch helpers ------------------------------------------------------------------
template <int ORIENT, int EPI, int BN>
int tc_run(catre_engine* e, cudaStream_t s, int grp, const CUtensorMap& ma_hi, const CUtensorMap& ma_lo,
           const CUtensorMap& nb_hi, const CUtensorMap& nb_lo, const TcGemmP& p, const TcPair* out = nullptr) {
  cudaError_t st;
  // point-on-lanes layers TMA-store their bf16 hi/lo output through the destination's [64 x 128-row] maps
  const CUtensorMap& oh = out ? out->map_hi : ma_hi;
  const CUtensorMap& ol = out ? out->map_lo : ma_lo;
  {
    Launch l(e, s, grp);
    if (e->cfg.precision == CATRE_PREC_BF16) st = tc_launch<ORIENT, EPI, BN, 1>(ma_hi, ma_lo, nb_hi, nb_lo, oh, ol, p, e->num_sms, s);
    else st = tc_launch<ORIENT, EPI, BN, 3>(ma_hi, ma_lo, nb_hi, nb_lo, oh, ol, p, e->num_sms, s);
  }
  if (st != cudaSuccess) {
    cudaGetLastError();
    return fail(e, CATRE_ERR_CUDA, "launch of tc %s failed: %s", kGrpNames[grp], cudaGetErrorString(st));
  }
  return 0;
}

// point-wise layer, points on TMEM lanes: act [R,K] (hi/lo) x W [C,K] -> relu(. + bias) re-split to bf16 hi/lo [R,C]
template <int BN>
int tc_split_layer(catre_engine* e, cudaStream_t s, int grp, const TcPair& act, const TcPair& w, int K, int C,
                   const float* bias, const TcPair& out, long long R, long long in_row0 = 0) {
  TcGemmP p{};
  p.K = K; p.m_tiles = (int)(R / 128); p.n_tiles = C / BN; p.rows_per_set = e->N; p.rows_per_obj = e->N + e->Np;
  p.bias = bias; p.relu = 1; p.mi_in0 = (int)(in_row0 / 128);
  return tc_run<PT_ON_LANES, EPI_SPLIT, BN>(e, s, grp, act.map_hi, act.map_lo, w.map_hi, w.map_lo, p, &out);
}

// point-wise layer + column max over the points of each set, channels on TMEM lanes
int tc_max_layer(catre_engine* e, cudaStream_t s, int grp, const TcPair& w, const CUtensorMap* act_nb, int K, int C,
                 const float* bias, int relu, int* gmax, long long R, long long set_row0 = 0) {
  TcGemmP p{};
  p.K = K; p.m_tiles = C / 128; p.n_tiles = (int)(R / 256);
  p.bias = bias; p.relu = relu; p.gmax = gmax; p.C = C; p.rows_per_set = e->N; p.rows_per_obj = e->N + e->Np;
  p.set_row0 = set_row0;
  return tc_run<CH_ON_LANES, EPI_MAX, 256>(e, s, grp, w.map_hi, w.map_lo, act_nb[0], act_nb[1], p);
}

// fused 64 -> 128 -> 1024 + column max of a T-Net (see enc_fused_kernel)
int tc_tnet_trunk(catre_engine* e, cudaStream_t s, int grp, const TcPair& act, const TcPair& w2, const float* b2, const TcPair& w3,
                  const float* b3, int* gmax, long long R) {
  EncFusedP p{};
  p.tiles = (int)(R / 256); p.rows_per_set = e->N; p.rows_per_obj = e->N + e->Np; p.bias2 = b2; p.bias3 = b3; p.gmax = gmax;
  cudaError_t st;
  {
    Launch l(e, s, grp);
    if (e->cfg.precision == CATRE_PREC_BF16)
      st = enc_fused_launch<1>(act.map_hi, act.map_lo, w2.map_hi, w2.map_lo, w3.map_hi, w3.map_lo, p, e->num_sms, s);
    else
      st = enc_fused_launch<3>(act.map_hi, act.map_lo, w2.map_hi, w2.map_lo, w3.map_hi, w3.map_lo, p, e->num_sms, s);
  }
  if (st != cudaSuccess) {
    cudaGetLastError();
    return fail(e, CATRE_ERR_CUDA, "launch of %s failed: %s", kGrpNames[grp], cudaGetErrorString(st));
  }
  return 0;
}

int tc_front(catre_engine* e, cudaStream_t s, const float* t3, const char* conv, long long R) {
  std::string c(conv);
  {
    Launch l(e, s, G_FRONT3);
    launch_pdl(e->cfg.precision == CATRE_PREC_F16X3 ? front3_split_kernel<true> : front3_split_kernel<false>, dim3((unsigned)((R + FRONT_PTS - 1) / FRONT_PTS)), dim3(256), (size_t)(0), s, e->q, t3, e->dw.at(c + ".weight"), e->dw.at(c + ".bias"),
                                                                 e->x64.hi, e->x64.lo, (int)R, e->N + e->Np, e->N);
  }
  return check_launch(e, "front3_split");
}

// ---- fused FC chains (fc_chain.cuh): one cluster launch per chain, fp32, same arithmetic at every batch size
FccLayer fcc_layer_desc(const catre_engine* e, const float* wp, const float* bias, int K, int C, int relu) {
  FccLayer l{};
  l.wp = wp; l.bias = bias; l.K = K; l.C = C; l.NC = fcc_nc(C, e->fcc_ranks); l.relu = relu;
  return l;
}

int run_chain(catre_engine* e, cudaStream_t s, int grp, const FccProblem& p0, const FccProblem* p1) {
  FccBatch b{};
  b.p[0] = p0;
  b.clusters0 = (p0.rows + FCC_ROWS - 1) / FCC_ROWS;
  int clusters = b.clusters0;
  if (p1) { b.p[1] = *p1; clusters += (p1->rows + FCC_ROWS - 1) / FCC_ROWS; }
  for (int q = 0; q < (p1 ? 2 : 1); ++q)
    for (int l = 0; l < b.p[q].n_layers; ++l)
      if (!fcc_layer_ok(b.p[q].L[l], e->fcc_ranks)) return fail(e, CATRE_ERR_UNSUPPORTED, "fc chain: unsupported layer geometry");
  cudaError_t st;
  {
    Launch l(e, s, grp);
    st = fcc_launch(b, clusters, e->fcc_ranks, s);
  }
  if (st != cudaSuccess) {
    cudaGetLastError();
    return fail(e, CATRE_ERR_CUDA, "launch of fc chain (%s) failed: %s", kGrpNames[grp], cudaGetErrorString(st));
  }
  return 0;
}

// one FC layer as a tiled GEMM in the chain's summation order (large row counts; bit-identical to the chain)
int run_fct(catre_engine* e, cudaStream_t s, int grp, const int* keys, long long lda_keys, const float* a32, int lda32, const float* w,
            const float* bias, int rows, int C, int K, int relu, float* out32, unsigned short* out_hi, unsigned short* out_lo) {
  FctP p{};
  p.keys = keys; p.lda_keys = lda_keys; p.a32 = a32; p.lda32 = lda32; p.w = w; p.ldw = K; p.bias = bias;
  p.rows = rows; p.C = C; p.K = K; p.relu = relu; p.out32 = out32; p.ldo = C; p.out_hi = out_hi; p.out_lo = out_lo;
  p.out_f16 = e->cfg.precision == CATRE_PREC_F16X3;
  fct_geometry(K, fcc_nc(C, e->fcc_ranks), &p.KS, &p.per, &p.kc);
  if (!fct_ok(p)) return fail(e, CATRE_ERR_UNSUPPORTED, "tiled fc: unsupported layer geometry (K=%d C=%d)", K, C);
  cudaError_t st;
  {
    Launch l(e, s, grp);
    st = fct_launch(p, s);
  }
  if (st != cudaSuccess) {
    cudaGetLastError();
    return fail(e, CATRE_ERR_CUDA, "launch of tiled fc (%s) failed: %s", kGrpNames[grp], cudaGetErrorString(st));
  }
  return 0;
}

// T-Net FC chain: keys [S,1024] -> 512 -> 256 -> kk (+I)  (pointnets/pointnet.py:32-40, 66-77); which = 0 stn, 1 fstn
int tnet_fc_chain(catre_engine* e, cudaStream_t s, const int* keys, int S, int which) {
  const std::string pf = which ? "pcl_net.fstn" : "pcl_net.stn";
  const bool tc = e->cfg.precision != CATRE_PREC_FP32_SIMT;
  if (S >= e->fct_min_rows) {  // large batch: three tiled GEMM launches, same bits as the cluster chain below
    int rc;
    if ((rc = run_fct(e, s, G_TNET_FC, keys, 1024, nullptr, 0, W(e, (pf + ".fc1.weight").c_str()), W(e, (pf + ".fc1.bias").c_str()), S, 512,
                      1024, 1, e->fc512, nullptr, nullptr))) return rc;
    if ((rc = run_fct(e, s, G_TNET_FC, nullptr, 0, e->fc512, 512, W(e, (pf + ".fc2.weight").c_str()), W(e, (pf + ".fc2.bias").c_str()), S, 256,
                      512, 1, e->fc256, nullptr, nullptr))) return rc;
    if (which == 0) {
      {
        Launch l(e, s, G_TNET_FC);
        launch_pdl(fc_small_kernel, dim3((unsigned)((S * 9 + 127) / 128)), dim3(128), (size_t)0, s, (const float*)e->fc256, 256,
                   W(e, "pcl_net.stn.fc3.weight"), 256, (const float*)e->stn_fc3_bI, S, 9, 256, e->t3);
      }
      return check_launch(e, "fc_small");
    }
    return run_fct(e, s, G_TNET_FC, nullptr, 0, e->fc256, 256, e->fstn_fc3_wT, e->fstn_fc3_bI, S, 4096, 256, 0, tc ? nullptr : e->t64,
                   tc ? reinterpret_cast<unsigned short*>(e->t64s.hi) : nullptr, tc ? reinterpret_cast<unsigned short*>(e->t64s.lo) : nullptr);
  }
  if (which == 1 && S >= e->fct_fc3_min_rows) {
    // fstn: fc1 + fc2 in the cluster chain (few columns, long K: needs the in-CTA k-slicing to fill an SM), fc3 as a
    // tiled GEMM (4096 columns, K = 256: a thousand independent tiles); same bits either way
    FccProblem p2{};
    p2.keys = keys; p2.lda = 1024; p2.rows = S; p2.n_layers = 2;
    p2.L[0] = fcc_layer_desc(e, e->fcc_fc1[1], W(e, "pcl_net.fstn.fc1.bias"), 1024, 512, 1);
    p2.L[1] = fcc_layer_desc(e, e->fcc_fc2[1], W(e, "pcl_net.fstn.fc2.bias"), 512, 256, 1);
    p2.out32 = e->fc256;
    int rc = run_chain(e, s, G_TNET_FC, p2, nullptr);
    if (rc) return rc;
    return run_fct(e, s, G_TNET_FC, nullptr, 0, e->fc256, 256, e->fstn_fc3_wT, e->fstn_fc3_bI, S, 4096, 256, 0, tc ? nullptr : e->t64,
                   tc ? reinterpret_cast<unsigned short*>(e->t64s.hi) : nullptr, tc ? reinterpret_cast<unsigned short*>(e->t64s.lo) : nullptr);
  }
  FccProblem p{};
  p.keys = keys; p.lda = 1024; p.rows = S; p.n_layers = 3;
  p.L[0] = fcc_layer_desc(e, e->fcc_fc1[which], W(e, (pf + ".fc1.bias").c_str()), 1024, 512, 1);
  p.L[1] = fcc_layer_desc(e, e->fcc_fc2[which], W(e, (pf + ".fc2.bias").c_str()), 512, 256, 1);
  if (which == 0) {
    p.L[2] = fcc_layer_desc(e, e->fcc_fc3[0], e->stn_fc3_bI, 256, 9, 0);
    p.out32 = e->t3;
  } else {  // fc3's rows are permuted at pack time so the chain emits T64^T (the feature transform's N-side operand)
    p.L[2] = fcc_layer_desc(e, e->fcc_fc3[1], e->fstn_fc3_bI, 256, 4096, 0);
    if (tc) {
      p.out_hi = reinterpret_cast<unsigned short*>(e->t64s.hi);
      p.out_lo = reinterpret_cast<unsigned short*>(e->t64s.lo);
      p.out_f16 = e->cfg.precision == CATRE_PREC_F16X3;
    } else {
      p.out32 = e->t64;
    }
  }
  return run_chain(e, s, G_TNET_FC, p, nullptr);
}

// One refinement iteration on a chunk of B objects whose points are already in e->q.
// head_done: the iteration's head kernel (iter_head_kernel) has already run stn.conv1.  defer_pose != nullptr: do not launch
// pose_update_kernel; return its arguments instead -- the next iteration's head kernel updates the pose itself.
int iteration(catre_engine* e, cudaStream_t s, int B, const float* pose_in, const float* scale_in, const float* K,
              float* pose_out, float* scale_out, const int* prior_cls = nullptr, int n_cls = 0, bool head_done = false,
              TsPoseP* defer_pose = nullptr) {
  const int N = e->N, S = 2 * B, P = e->N + e->Np;
  const long long R = (long long)B * P;
  int rc;
  const bool tc = e->cfg.precision != CATRE_PREC_FP32_SIMT;

  // column-max key buffers are laid out for the current S; the point kernel that filled e->q has reset them
  e->gmax_stn = e->gmax_all;
  e->gmax_fstn = e->gmax_all + (size_t)S * 1024;
  e->gmax_g = e->gmax_all + (size_t)S * 2048;
  e->gmax_pf = e->gmax_all + (size_t)S * 3072;

  // ---- E1: STN3d (pointnets/pointnet.py:24-41)
  if (tc) {
    if (!head_done && (rc = tc_front(e, s, nullptr, "pcl_net.stn.conv1", R))) return rc;
    if ((rc = tc_tnet_trunk(e, s, G_STN_CONV3_MAX, e->x64, e->tw_stn_c2, W(e, "pcl_net.stn.conv2.bias"), e->tw_stn_c3,
                            W(e, "pcl_net.stn.conv3.bias"), e->gmax_stn, R))) return rc;
  } else {
    {
      Launch l(e, s, G_FRONT3);
      launch_pdl(front3_kernel, dim3((unsigned)((R * 16 + 255) / 256)), dim3(256), (size_t)(0), s, e->q, nullptr, W(e, "pcl_net.stn.conv1.weight"),
                                                                    W(e, "pcl_net.stn.conv1.bias"), e->h64a, R, P, N);
    }
    if ((rc = check_launch(e, "front3"))) return rc;
    GemmP p = gemm_args(e->h64a, 64, W(e, "pcl_net.stn.conv2.weight"), 64, 128, W(e, "pcl_net.stn.conv2.bias"), e->h128,
                        128, R, 1);
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_STN_CONV2, p))) return rc;
    p = gemm_args(e->h128, 128, W(e, "pcl_net.stn.conv3.weight"), 128, 1024, W(e, "pcl_net.stn.conv3.bias"), nullptr, 0,
                  R, 1);
    p.gmax = e->gmax_stn; p.rows_per_set = N; p.rows_per_obj = P;
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_STN_CONV3_MAX, p))) return rc;
  }
  if ((rc = tnet_fc_chain(e, s, e->gmax_stn, S, 0))) return rc;

  // ---- E2: input transform + conv1 (pointnet.py:97-103)
  if (tc) {
    if ((rc = tc_front(e, s, e->t3, "pcl_net.conv1", R))) return rc;
  } else {
    {
      Launch l(e, s, G_FRONT3);
      launch_pdl(front3_kernel, dim3((unsigned)((R * 16 + 255) / 256)), dim3(256), (size_t)(0), s, e->q, e->t3, W(e, "pcl_net.conv1.weight"),
                                                                    W(e, "pcl_net.conv1.bias"), e->h64a, R, P, N);
    }
    if ((rc = check_launch(e, "front3"))) return rc;
  }

  // ---- E3: STNkd (pointnet.py:57-78).  fc3's rows are permuted at pack time so it emits T64^T.
  if (tc) {
    if ((rc = tc_split_layer<64>(e, s, G_FSTN_CONV1, e->x64, e->tw_fstn_c1, 64, 64, W(e, "pcl_net.fstn.conv1.bias"), e->f64, R))) return rc;
    if ((rc = tc_tnet_trunk(e, s, G_FSTN_CONV3_MAX, e->f64, e->tw_fstn_c2, W(e, "pcl_net.fstn.conv2.bias"), e->tw_fstn_c3,
                            W(e, "pcl_net.fstn.conv3.bias"), e->gmax_fstn, R))) return rc;
  } else {
    GemmP p = gemm_args(e->h64a, 64, W(e, "pcl_net.fstn.conv1.weight"), 64, 64, W(e, "pcl_net.fstn.conv1.bias"), e->h64b,
                        64, R, 1);
    if ((rc = run_gemm<64, A_PLAIN>(e, s, G_FSTN_CONV1, p))) return rc;
    p = gemm_args(e->h64b, 64, W(e, "pcl_net.fstn.conv2.weight"), 64, 128, W(e, "pcl_net.fstn.conv2.bias"), e->h128, 128,
                  R, 1);
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_FSTN_CONV2, p))) return rc;
    p = gemm_args(e->h128, 128, W(e, "pcl_net.fstn.conv3.weight"), 128, 1024, W(e, "pcl_net.fstn.conv3.bias"), nullptr,
                  0, R, 1);
    p.gmax = e->gmax_fstn; p.rows_per_set = N; p.rows_per_obj = P;
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_FSTN_CONV3_MAX, p))) return rc;
  }
  if ((rc = tnet_fc_chain(e, s, e->gmax_fstn, S, 1))) return rc;

  // ---- E4: feature transform pf = h1 . T64 (per set), trunk conv2-4, global max (pointnet.py:105-116)
  if (tc) {
    TcGemmP p{};
    p.K = 64; p.m_tiles = (int)(R / 128); p.n_tiles = 1; p.rows_per_set = N; p.rows_per_obj = P; p.nb_per_set = 1;
    p.gmax = e->gmax_pf; p.C = 64;
    if ((rc = tc_run<PT_ON_LANES, EPI_SPLIT_MAX, 64>(e, s, G_FEAT_TRANSFORM, e->x64.map_hi, e->x64.map_lo, e->t64s.map_hi,
                                                      e->t64s.map_lo, p, &e->pf16))) return rc;
    if ((rc = tc_split_layer<128>(e, s, G_CONV2, e->pf16, e->tw_conv2, 64, 128, W(e, "pcl_net.conv2.bias"), e->a128, R))) return rc;
    // conv3 -> conv4 in groups of trunk_group objects: the group's [rows, 512] hi/lo activations (4 KB per point) are
    // written and read back while they are still in L2 and the next group overwrites the same lines, so this
    // intermediate -- the largest of the chain -- costs no HBM traffic and the workspace holds one group of it
    const int G = e->trunk_group > 0 ? e->trunk_group : B;
    for (int g0 = 0; g0 < B; g0 += G) {
      const long long row0 = (long long)g0 * P, Rg = (long long)((B - g0 < G) ? B - g0 : G) * P;
      if ((rc = tc_split_layer<128>(e, s, G_CONV3, e->a128, e->tw_conv3, 128, 512, W(e, "pcl_net.conv3.bias"), e->a512, Rg, row0))) return rc;
      if ((rc = tc_max_layer(e, s, G_CONV4_MAX, e->tw_conv4, e->a512_nb, 512, 1024, W(e, "pcl_net.conv4.bias"), 0, e->gmax_g, Rg, row0))) return rc;
    }
  } else {
    GemmP p = gemm_args(e->h64a, 64, e->t64, 64, 64, nullptr, e->h64b, 64, R, 0);
    p.w_set_stride = 4096; p.rows_per_set = N; p.rows_per_obj = P;  // W[c][k] = T64^T[set][c][k]
    p.gmax = e->gmax_pf;
    if ((rc = run_gemm<64, A_PLAIN>(e, s, G_FEAT_TRANSFORM, p))) return rc;
    p = gemm_args(e->h64b, 64, W(e, "pcl_net.conv2.weight"), 64, 128, W(e, "pcl_net.conv2.bias"), e->h128, 128, R, 1);
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_CONV2, p))) return rc;
    p = gemm_args(e->h128, 128, W(e, "pcl_net.conv3.weight"), 128, 512, W(e, "pcl_net.conv3.bias"), e->h512, 512, R, 1);
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_CONV3, p))) return rc;
    p = gemm_args(e->h512, 512, W(e, "pcl_net.conv4.weight"), 512, 1024, W(e, "pcl_net.conv4.bias"), nullptr, 0, R, 0);
    p.gmax = e->gmax_g; p.rows_per_set = N; p.rows_per_obj = P;
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_CONV4_MAX, p))) return rc;
  }

  // ---- the two FC layers over the max-pooled global feature, one launch: cset = W0[:, :1024] . g_set + b0 for every
  //      set (rot layer-0 split) and ts-head layer 0 over the OBSERVED sets' g (row b -> set 2b)
  if (S >= e->fct_min_rows) {
    if ((rc = run_fct(e, s, G_ROT_GFEAT, e->gmax_g, 1024, nullptr, 0, e->rot_w0g, e->rot_b0, S, 512, 1024, 0, e->cset, nullptr, nullptr))) return rc;
    if ((rc = run_fct(e, s, G_ROT_GFEAT, e->gmax_g, 2048, nullptr, 0, e->ts_w0g, W(e, "ts_head.linears.0.bias"), B, 256, 1024, 0, e->ts0,
                      nullptr, nullptr))) return rc;
  } else {
    FccProblem pc{}, pt{};
    pc.keys = e->gmax_g; pc.lda = 1024; pc.rows = S; pc.n_layers = 1;
    pc.L[0] = fcc_layer_desc(e, e->fcc_cset, e->rot_b0, 1024, 512, 0);
    pc.out32 = e->cset;
    pt.keys = e->gmax_g; pt.lda = 2048; pt.rows = B; pt.n_layers = 1;
    pt.L[0] = fcc_layer_desc(e, e->fcc_ts0, W(e, "ts_head.linears.0.bias"), 1024, 256, 0);
    pt.out32 = e->ts0;
    if ((rc = run_chain(e, s, G_ROT_GFEAT, pc, &pt))) return rc;
  }
  // ---- H1 on the side stream (needs only the max-pooled features; independent of the rot head): the remaining
  //      67 inputs of ts layer 0 (pointfeat max, init scale) are added in ts_head_kernel.
  TsPoseP tsp{};
  {
    tsp.ts0 = e->ts0; tsp.gmax_pf = e->gmax_pf; tsp.dts = e->dts;
    tsp.w0t = e->ts_w0t; tsp.g0 = W(e, "ts_head.linears.1.weight"); tsp.be0 = W(e, "ts_head.linears.1.bias");
    tsp.w1t = e->ts_w1t; tsp.b1 = W(e, "ts_head.linears.3.bias"); tsp.g1 = W(e, "ts_head.linears.4.weight");
    tsp.be1 = W(e, "ts_head.linears.4.bias");
    tsp.wt = W(e, "ts_head.fc_t.weight"); tsp.bt = W(e, "ts_head.fc_t.bias");
    tsp.ws = W(e, "ts_head.fc_s.weight"); tsp.bs = W(e, "ts_head.fc_s.bias");
    tsp.rot_partial = e->rot_partial; tsp.rot_tiles = (tc && !e->fused_tail) ? 16 : P / 128; tsp.convp_bias = e->convp_b;
    tsp.pose_in = pose_in; tsp.scale_in = scale_in; tsp.K = K; tsp.pose_out = pose_out; tsp.scale_out = scale_out;
    tsp.cls = prior_cls; tsp.n_cls = n_cls;
    CU_TRY(e, cudaEventRecord(e->ev_fork, s));
    CU_TRY(e, cudaStreamWaitEvent(e->side, e->ev_fork, 0));
    {
      Launch l(e, e->side, G_TS_POSE);
      launch_pdl(ts_head_kernel, dim3(B), dim3(256), (size_t)(0), e->side, tsp);
    }
    if ((rc = check_launch(e, "ts_head"))) return rc;
    CU_TRY(e, cudaEventRecord(e->ev_join, e->side));
  }

  float* gn0_scale = e->gn0;
  float* gn0_shift = e->gn0 + (size_t)e->maxB * 1024;
  if (tc) {
    // pass 1: GroupNorm statistics of a0 = W0p . pf + cset (nothing stored; K = 64 makes the recompute cheap)
    TcGemmP p{};
    p.K = 64; p.m_tiles = 4; p.n_tiles = (int)(R / 256); p.rows_per_set = N; p.rows_per_obj = P;
    p.rowvec = e->cset; p.ldrv = 512;
    p.stats = e->stats0; p.stats_ld = 64; p.stats_goff = 0;  // partials per 128 points (BN = 256, two column halves)
    if ((rc = tc_run<CH_ON_LANES, EPI_STATS, 256>(e, s, G_ROT_LAYER0, e->tw_rot0.map_hi, e->tw_rot0.map_lo, e->pf_nb[0], e->pf_nb[1], p))) return rc;
    {
      Launch l(e, s, G_GN_FINALIZE);
      launch_pdl(gn_finalize_set_kernel, dim3((B * 64 + 127) / 128), dim3(128), (size_t)(0), s, e->stats0, e->rot_gn0_g, e->rot_gn0_b, e->cset, gn0_scale,
                                                                 gn0_shift, B, 512, P / 128, P, e->rot_count);
    }
    if ((rc = check_launch(e, "gn_finalize_set"))) return rc;
    {
      // layer-0 recompute + GroupNorm + GELU + bf16 split into shared memory + layer 1, one kernel
      RotFusedP pf{};
      pf.tiles = (int)(R / 128); pf.rows_per_set = N; pf.rows_per_obj = P;
      pf.gn_scale = gn0_scale; pf.gn_shift = gn0_shift; pf.bias1 = e->rot_b1; pf.stats = e->stats1;
      pf.obj_count = e->rot_count; pf.gn1_gamma = e->rot_gn1_g; pf.gn1_beta = e->rot_gn1_b;
      pf.neck_w = e->neck_w; pf.neck_b = e->neck_b; pf.wp = e->wp; pf.partial = e->rot_partial;
      pf.a1t = (!e->fused_tail || e->debug_taps) ? e->a1 : nullptr;
      cudaError_t st;
      {
        Launch l(e, s, G_ROT_FUSED);
#define CATRE_ROT_LAUNCH(NP, FT)                                                                                          \
  rot_fused_launch<NP, FT>(e->pf16.map_hi, e->pf16.map_lo, e->tw_rot0.map_hi, e->tw_rot0.map_lo, e->tw_rot1s.map_hi, \
                           e->tw_rot1s.map_lo, pf, e->num_sms, s)
        if (e->cfg.precision == CATRE_PREC_BF16) st = e->fused_tail ? CATRE_ROT_LAUNCH(1, true) : CATRE_ROT_LAUNCH(1, false);
        else st = e->fused_tail ? CATRE_ROT_LAUNCH(3, true) : CATRE_ROT_LAUNCH(3, false);
#undef CATRE_ROT_LAUNCH
      }
      if (st != cudaSuccess) {
        cudaGetLastError();
        return fail(e, CATRE_ERR_CUDA, "launch of rot_fused failed: %s", cudaGetErrorString(st));
      }
    }
  } else {
    GemmP p = gemm_args(e->h64b, 64, e->rot_w0p, 64, 512, nullptr, e->a0, 512, R, 0);
    p.rowvec = e->cset; p.ldrv = 512; p.rows_per_set = N; p.rows_per_obj = P;
    p.stats = e->stats0; p.stats_ld = 64; p.stats_goff = 0;
    if ((rc = run_gemm<128, A_PLAIN>(e, s, G_ROT_LAYER0, p))) return rc;
    {
      Launch l(e, s, G_GN_FINALIZE);
      launch_pdl(gn_finalize_kernel, dim3((B * 64 + 127) / 128), dim3(128), (size_t)(0), s, e->stats0, e->rot_gn0_g, e->rot_gn0_b, gn0_scale, gn0_shift, B,
                                                             512, P / 128, P);
    }
    if ((rc = check_launch(e, "gn_finalize"))) return rc;
    for (int h = 0; h < 2; ++h) {
      const char* wn = h == 0 ? "rot_head.rot_head_x.layers.3.weight" : "rot_head.rot_head_y.layers.3.weight";
      GemmP p1 = gemm_args(e->a0 + h * 256, 512, W(e, wn), 256, 256, e->rot_b1 + h * 256, e->a1 + h * 256, 512, R, 0);
      p1.rows_per_set = N; p1.rows_per_obj = P;
      p1.gn_scale = gn0_scale + h * 256; p1.gn_shift = gn0_shift + h * 256; p1.ldgn = 512;
      p1.stats = e->stats1; p1.stats_ld = 64; p1.stats_goff = 32 * h;
      if ((rc = run_gemm<128, A_GN_GELU>(e, s, G_ROT_LAYER1, p1))) return rc;
    }
  }
  if (!tc) {
    Launch l(e, s, G_GN_FINALIZE);
    launch_pdl(gn_finalize_kernel, dim3((B * 64 + 127) / 128), dim3(128), (size_t)(0), s, e->stats1, e->rot_gn1_g, e->rot_gn1_b, e->gn1,
               e->gn1 + (size_t)e->maxB * 512, B, 512, P / 128, P);
  }
  if ((rc = check_launch(e, "gn_finalize"))) return rc;
  if (!tc) {
    Launch l(e, s, G_ROT_TAIL);
    launch_pdl(rot_tail_kernel, dim3(P / 128, B), dim3(256), (size_t)(0), s, e->a1, e->gn1, e->gn1 + (size_t)e->maxB * 512, e->neck_w,
               e->neck_b, e->wp, e->rot_partial, P);
  } else if (!e->fused_tail) {  // split tail: finalises the GroupNorm-1 statistics itself (partials per 128 points from the fused rot
                                // kernel) and walks the objects backwards (the last ones written are still in L2)
    Launch l(e, s, G_ROT_TAIL);
    launch_pdl(rot_tail_t_kernel, dim3(16, B), dim3(256), (size_t)(P * sizeof(float)), s, e->a1, e->stats1, e->rot_gn1_g, e->rot_gn1_b,
               P / 128, e->neck_w, e->neck_b, e->wp, e->rot_partial, P, 1);
  }  // fused tail: the tail ran inside the fused rot kernel, out of TMEM
  if ((rc = check_launch(e, "rot_tail"))) return rc;

  // ---- G1 + G2 after both heads: join the side stream, then rot6d Gram-Schmidt + pose update
  CU_TRY(e, cudaStreamWaitEvent(s, e->ev_join, 0));
  if (defer_pose != nullptr) {
    *defer_pose = tsp;
    return 0;
  }
  {
    Launch l(e, s, G_TS_POSE);
    launch_pdl(pose_update_kernel, dim3((B + 7) / 8), dim3(256), (size_t)(0), s, tsp, B);
  }
  return check_launch(e, "pose_update");
}

// Head of an iteration in the tensor-core modes: [previous iteration's pose update] + point update + stn.conv1, one launch.
int iter_head(catre_engine* e, cudaStream_t s, int B, const TsPoseP* prev, const float* pose, const float* scale, const float* pcl,
              const float* prior, const int* cls, int n_cls) {
  IterHeadP h{};
  if (prev) { h.prev = *prev; h.have_prev = 1; }
  h.pose = pose; h.scale = scale; h.pcl = pcl; h.prior = prior; h.cls = cls; h.n_cls = n_cls;
  h.q = e->q; h.gmax = e->gmax_all; h.n_keys = (long long)(2 * B) * (1024 * 3 + 64);
  h.W = W(e, "pcl_net.stn.conv1.weight"); h.bias = W(e, "pcl_net.stn.conv1.bias");
  h.out_hi = e->x64.hi; h.out_lo = e->x64.lo; h.B = B; h.N = e->N; h.Np = e->Np;
  const long long R = (long long)B * (e->N + e->Np);
  {
    Launch l(e, s, G_FRONT3);
    launch_pdl(e->cfg.precision == CATRE_PREC_F16X3 ? iter_head_kernel<true> : iter_head_kernel<false>, dim3((unsigned)(R / FRONT_PTS)),
               dim3(256), (size_t)0, s, h);
  }
  return check_launch(e, "iter_head");
}

int check_ready(catre_engine* e, int B) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!e->packed) return fail(e, CATRE_ERR_NOT_PACKED, "catre_pack has not been called (or weights changed since)");
  if (B < 0) return fail(e, CATRE_ERR_INVALID_ARG, "negative batch %d", B);
  return 0;
}

}  // namespace

// ---- training step: CUDA launches behind train_chain.cuh's Ops interface
namespace {
struct CudaTrainOps {
  cudaStream_t s;
  bool naive;
  bool v2;  // CATRE_TRAIN_GEMM=v2: the 128 x BN register-prefetch CUDA-core GEMM (measured slower than the 64 x 64 one)
  bool tc;  // tcgen05 GEMM for the large shapes (default; CATRE_TRAIN_GEMM=simt / v2 / CATRE_TRAIN_NAIVE_GEMM=1 turn it off)
  int num_sms;
  int64_t launches = 0;
  cudaError_t err = cudaSuccess;
  void note(cudaError_t st) { if (err == cudaSuccess && st != cudaSuccess) err = st; }
  // lanes (train_chain.cuh: the ts head next to the rotation heads): a second stream between fork() and join(), ordered by two
  // events; under stream capture the same calls become parallel branches of the graph.  CATRE_TRAIN_LANES=0: one lane.
  cudaStream_t s_main = nullptr, s_side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  bool lanes = false;
  void fork() {
    if (!lanes) return;
    note(cudaEventRecord(ev_fork, s_main));
    for (int i = 0; i < 2; ++i) note(cudaStreamWaitEvent(s_side[i], ev_fork, 0));
  }
  void lane(int i) { if (lanes) s = i ? s_side[i - 1] : s_main; }
  void join() {
    if (!lanes) return;
    for (int i = 0; i < 2; ++i) {
      note(cudaEventRecord(ev_join[i], s_side[i]));
      note(cudaStreamWaitEvent(s_main, ev_join[i], 0));
    }
    s = s_main;
  }
  // carve: ask for the maximum shared-memory carve-out for the small kernels too, so that the SMs do not switch their L1 /
  // shared-memory split before and after every tensor-core GEMM (193 KB of shared memory) of the chain; set once per kernel
  // instantiation and value (the attribute belongs to the function, not to the launch).  CATRE_TRAIN_CARVEOUT=0 turns it off.
  bool carve = false;
  template <class Fn>
  void set_carve(Fn* fn, int& state) {
    const int want = carve ? 100 : -1;
    if (state != want) {
      note(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, want));
      state = want;
    }
  }
  template <class KF>
  void run(const KF& k, unsigned gx, unsigned gy, unsigned gz, unsigned nt) {
    if (gx == 0 || gy == 0 || gz == 0) return;
    static int carve_state = -1;  // one per instantiation
    set_carve(catre_train::tk_run<KF>, carve_state);
    catre_train::tk_run<KF><<<dim3(gx, gy, gz), nt, 0, s>>>(k);
    ++launches;
    note(cudaPeekAtLastError());
  }
  void zero(void* p, size_t bytes) { note(cudaMemsetAsync(p, 0, bytes, s)); }
  // The tensor-core GEMM (train_gemm_tc.cuh) takes every GEMM with enough work for a 128 x 128 x 64 tile pipeline; the rest
  // (K = 3 input layers, the 3- / 6- / 9-column heads, single-row products) stays on the CUDA-core tile kernel.
  static bool tc_shape(const catre_train::GemmP& p, int bz) {
    return p.M >= 16 && p.N >= 16 && p.K >= 16 && (double)p.M * p.N * p.K * (p.splits > 1 ? 1 : bz) >= (double)(1 << 20);
  }
  // bias gradient inside a split-K weight-gradient launch (train_gemm_tc.cuh): the tensor-core path with A = dy^T (m fast)
  bool folds_bias_grad(const catre_train::GemmP& p) const { return tc && !naive && fold_bg && p.sam == 1 && p.splits > 1 && tc_shape(p, 1); }
  bool fold_bg = true;  // CATRE_TRAIN_FOLD_BIAS=0: column sums as separate launches
  void gemm_tc(const catre_train::GemmP& p, int bz) {
    note(p.f16 ? catre_train::tk_gemm_tc_launch<true>(p, bz, s) : catre_train::tk_gemm_tc_launch<false>(p, bz, s));
    ++launches;
  }
  // max-pooled layer with the pooling fused into the tensor-core GEMM's epilogue (64-bit value|~row keys + atomicMax, decoded
  // into vmax / arg afterwards); false = not available, the chain then runs the layer and the column max separately
  bool gemm_colmax(const catre_train::GemmP& p, int rows_per_set, float* vmax, int* arg, int S, size_t partial_floats) {
    if (!tc || rows_per_set % 128 != 0 || (size_t)S * p.N * 2 > partial_floats || !tc_shape(p, 1)) return false;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(p.partial);
    zero(keys, (size_t)S * p.N * sizeof(unsigned long long));
    const catre_train::TgColMax cm{keys, rows_per_set};
    note(p.f16 ? catre_train::tk_gemm_tc_launch<true>(p, 1, s, cm) : catre_train::tk_gemm_tc_launch<false>(p, 1, s, cm));
    ++launches;
    const long long n = (long long)S * p.N;
    run(catre_train::KColMaxDecode{keys, vmax, arg, n}, (unsigned)((n + 255) / 256), 1, 1, 256);
    return true;
  }
  void gemm(const catre_train::GemmP& p, int bz) {
    if (p.M <= 0 || p.N <= 0 || bz <= 0) return;
    if (naive) {
      run(catre_train::KGemmNaive{p}, (unsigned)((p.M + 3) / 4), (unsigned)((p.N + 63) / 64), (unsigned)bz, 256);
    } else if (tc && tc_shape(p, bz)) {
      // a launch of a few tiles with a long reduction (the small-M FC layers: 32 rows, K up to 1024): split K here so that
      // the tiles x splits CTAs cover the device; the partials are summed in a fixed order (+ bias / ReLU) by KSplitReduce
      const long long tiles = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128);
      if (p.splits == 1 && bz == 1 && tiles * 2 <= num_sms && p.K >= 256 && p.partial) {
        int splits = (int)(num_sms / tiles);
        if (splits > p.K / 128) splits = p.K / 128;
        while (splits > 1 && (size_t)splits * p.M * p.N > catre_train::TrainWs::kPartialFloats) --splits;
        if (splits > 1) {
          catre_train::GemmP q = p;
          q.splits = splits;
          q.k_per = ((p.K + splits - 1) / splits + 63) & ~63;
          gemm_tc(q, splits);
          catre_train::KSplitReduce red{p.partial, p.C, p.scm, p.scn, p.M, p.N, splits, p.accumulate, p.bias, p.relu};
          run(red, (unsigned)(((long long)p.M * p.N + 255) / 256), 1, 1, 256);
          return;
        }
      }
      gemm_tc(p, bz);
    } else if (v2) {
      if (p.N <= 64) catre_train::tk_gemm_tiled2<64><<<dim3((unsigned)((p.M + 127) / 128), (unsigned)((p.N + 63) / 64), (unsigned)bz), 256, 0, s>>>(p);
      else catre_train::tk_gemm_tiled2<128><<<dim3((unsigned)((p.M + 127) / 128), (unsigned)((p.N + 127) / 128), (unsigned)bz), 256, 0, s>>>(p);
      ++launches;
      note(cudaPeekAtLastError());
    } else {
      static int carve_state = -1;
      set_carve(catre_train::tk_gemm_tiled, carve_state);
      catre_train::tk_gemm_tiled<<<dim3((unsigned)((p.M + 63) / 64), (unsigned)((p.N + 63) / 64), (unsigned)bz), 256, 0, s>>>(p);
      ++launches;
      note(cudaPeekAtLastError());
    }
  }
};

int weight_index(const char* name) {
  for (int i = 0; i < kNumWeights; ++i)
    if (strcmp(kWeights[i].name, name) == 0) return i;
  return -1;
}
}  // namespace

// ================================================================================================
extern "C" {

const char* catre_version(void) { return "catre_b200 0.2 sm_100a (fp32 SIMT + tcgen05 f16x3 | bf16)"; }
int32_t catre_num_weights(void) { return kNumWeights; }
const char* catre_weight_name(int32_t i) { return (i >= 0 && i < kNumWeights) ? kWeights[i].name : nullptr; }
int32_t catre_profile_num(void) { return G_NUM; }
const char* catre_profile_name(int32_t i) { return (i >= 0 && i < G_NUM) ? kGrpNames[i] : nullptr; }

const char* catre_last_error(const catre_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int catre_create(catre_engine** out, const catre_cfg* cfg) {
  if (!out || !cfg) return fail(nullptr, CATRE_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (cfg->n_obs <= 0 || cfg->n_obs % 128 != 0 || cfg->n_prior <= 0 || cfg->n_prior % 128 != 0)
    return fail(nullptr, CATRE_ERR_UNSUPPORTED, "n_obs=%d and n_prior=%d must be positive multiples of 128", cfg->n_obs, cfg->n_prior);
  if (cfg->max_batch < 1) return fail(nullptr, CATRE_ERR_INVALID_ARG, "max_batch=%d must be >= 1", cfg->max_batch);
  if (cfg->precision < CATRE_PREC_FP32_SIMT || cfg->precision > CATRE_PREC_BF16)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "unknown precision %d", cfg->precision);
  if (cfg->precision != CATRE_PREC_FP32_SIMT && (cfg->n_obs % 256 != 0 || cfg->n_prior % 256 != 0))
    return fail(nullptr, CATRE_ERR_UNSUPPORTED, "tensor-core modes need n_obs and n_prior %% 256 == 0 (got %d, %d)", cfg->n_obs, cfg->n_prior);
  int ndev = 0;
  cudaError_t st = cudaGetDeviceCount(&ndev);
  if (st != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, CATRE_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(st));
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CATRE_ERR_INVALID_ARG, "device %d out of range", cfg->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
    return fail(nullptr, CATRE_ERR_NO_DEVICE, "device %d is not an sm_100 (Blackwell) GPU", cfg->device);
  if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(nullptr, CATRE_ERR_CUDA, "cudaSetDevice failed");

  catre_engine* e = new catre_engine();
  e->cfg = *cfg;
  e->N = cfg->n_obs;
  e->Np = cfg->n_prior;
  e->maxB = cfg->max_batch;
  const size_t B = e->maxB, S = 2 * B, N = e->N, Np = e->Np, P = N + Np, R = B * P;
  int rc = 0;
  const bool tc = cfg->precision != CATRE_PREC_FP32_SIMT;
  rc |= dalloc(e, &e->q, R * 3);
  if (!tc) {  // fp32 activations of the CUDA-core mode; the tensor-core modes keep bf16 hi/lo operands only
    rc |= dalloc(e, &e->h64a, R * 64);
    rc |= dalloc(e, &e->h64b, R * 64);
    rc |= dalloc(e, &e->h128, R * 128);
    rc |= dalloc(e, &e->h512, R * 512);
    rc |= dalloc(e, &e->a0, R * 512);
    rc |= dalloc(e, &e->t64, S * 4096);
  }
  // fp32 mode: a1 [R][512], the rot layer-1 output; tensor-core modes with the split rot tail (default): a1T [B][P/4][512][4].
  // With the fused tail (CATRE_ROT_TAIL=fused) it is never stored; CATRE_DEBUG_TAPS=1 then keeps a copy for the per-stage test
  {
    const char* rt = getenv("CATRE_ROT_TAIL");
    e->fused_tail = tc && rt && strcmp(rt, "fused") == 0;
  }
  if (!tc || !e->fused_tail || (getenv("CATRE_DEBUG_TAPS") && atoi(getenv("CATRE_DEBUG_TAPS")) != 0)) rc |= dalloc(e, &e->a1, R * 512);
  rc |= dalloc(e, &e->rot_count, 2 * B);
  rc |= dalloc(e, &e->gmax_all, S * (1024 * 3 + 64));
  e->gmax_stn = e->gmax_all;
  e->gmax_fstn = e->gmax_all + S * 1024;
  e->gmax_g = e->gmax_all + S * 2048;
  e->gmax_pf = e->gmax_all + S * 3072;
  rc |= dalloc(e, &e->fc512, S * 512);
  rc |= dalloc(e, &e->fc256, S * 256);
  rc |= dalloc(e, &e->ts0, B * 256);
  rc |= dalloc(e, &e->dts, B * 6);
  rc |= dalloc(e, &e->t3, S * 9 + 7);
  rc |= dalloc(e, &e->cset, S * 512);
  rc |= dalloc(e, &e->stats0, (R / 64) * 64 * 2);
  rc |= dalloc(e, &e->stats1, (R / 64) * 64 * 2);  // the fused rot kernel emits partials per 64 points
  rc |= dalloc(e, &e->gn0, B * 2048);  // scale | shift, per set in the tensor-core modes
  rc |= dalloc(e, &e->gn1, B * 512 * 2);
  rc |= dalloc(e, &e->rot_partial, B * (P / 128 > 16 ? P / 128 : 16) * 6);
  rc |= dalloc(e, &e->st_pcl, B * N * 3);
  e->st_prior_rows = e->maxB > 16 ? e->maxB : 16;
  rc |= dalloc(e, &e->st_prior, (size_t)e->st_prior_rows * Np * 3);
  rc |= dalloc(e, &e->st_cls, B);
  rc |= dalloc(e, &e->st_pose, B * 12);
  rc |= dalloc(e, &e->st_scale, B * 3);
  rc |= dalloc(e, &e->st_K, B * 9);
  rc |= dalloc(e, &e->st_oposes, B * 12 * (catre_engine::kMaxHostIter + 1));
  rc |= dalloc(e, &e->st_oscales, B * 3 * (catre_engine::kMaxHostIter + 1));
  e->num_sms = prop.multiProcessorCount;
  {
    // 8-CTA clusters measured faster than 16 on B200 (profiles/r02_ncu_fc_chain.txt); CATRE_FC_RANKS=16 selects the
    // non-portable size where the device can co-schedule it (experiments)
    const char* env = getenv("CATRE_FC_RANKS");
    e->fcc_ranks = (env && atoi(env) == 16) ? fcc_pick_ranks() : 8;
    const char* env2 = getenv("CATRE_FC_TILED_MIN_ROWS");  // experiments: where the tiled-GEMM FC path takes over (same bits)
    if (env2 && atoi(env2) > 0) e->fct_min_rows = atoi(env2);
    const char* env3 = getenv("CATRE_FC3_TILED_MIN_ROWS");
    if (env3 && atoi(env3) > 0) e->fct_fc3_min_rows = atoi(env3);
    const char* env5 = getenv("CATRE_DEBUG_TAPS");
    e->debug_taps = env5 && atoi(env5) != 0;
    const char* env6 = getenv("CATRE_TRUNK_GROUP");
    if (env6 && atoi(env6) >= 0) e->trunk_group = atoi(env6);
  }
  if (cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->side2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming) != cudaSuccess) rc |= 1;
  if (!rc && tc) {
    auto pair = [&](TcPair& t, size_t cols, size_t rows) {
      rc |= dalloc(e, &t.hi, rows * cols);
      rc |= dalloc(e, &t.lo, rows * cols);
      if (rc) return;
      if (!tc_make_map(&t.map_hi, t.hi, rows, cols, cols, 128) || !tc_make_map(&t.map_lo, t.lo, rows, cols, cols, 128)) rc |= 2;
    };
    // conv3's output only ever holds one object group (see iteration()): trunk_group objects instead of max_batch
    const size_t R512 = (e->trunk_group > 0 && (size_t)e->trunk_group < B) ? (size_t)e->trunk_group * P : R;
    pair(e->x64, 64, R); pair(e->f64, 64, R); pair(e->a128, 128, R); pair(e->pf16, 64, R); pair(e->a512, 512, R512);
    const size_t Spad = (S + 127) / 128 * 128;
    if (!rc) {  // T64^T per set, the N-side operand of the feature transform: [S*64, 64], 64-row boxes
      rc |= dalloc(e, &e->t64s.hi, Spad * 4096);
      rc |= dalloc(e, &e->t64s.lo, Spad * 4096);
      if (!rc && (!tc_make_map(&e->t64s.map_hi, e->t64s.hi, S * 64, 64, 64, 64) ||
                  !tc_make_map(&e->t64s.map_lo, e->t64s.lo, S * 64, 64, 64, 64))) rc |= 2;
    }
    if (!rc) {
      bool ok = true;
      ok &= tc_make_map(&e->a128_nb[0], e->a128.hi, R, 128, 128, 256) && tc_make_map(&e->a128_nb[1], e->a128.lo, R, 128, 128, 256);
      ok &= tc_make_map(&e->pf_nb[0], e->pf16.hi, R, 64, 64, 256) && tc_make_map(&e->pf_nb[1], e->pf16.lo, R, 64, 64, 256);
      ok &= tc_make_map(&e->a512_nb[0], e->a512.hi, R512, 512, 512, 256) && tc_make_map(&e->a512_nb[1], e->a512.lo, R512, 512, 512, 256);
      if (!ok) rc |= 2;
    }
    if (rc & 2) e->err = "cuTensorMapEncodeTiled failed for an activation buffer";
  }
  if (rc) {
    g_create_error = e->err;
    catre_destroy(e);
    return CATRE_ERR_CUDA;
  }
  *out = e;
  return CATRE_OK;
}

void catre_destroy(catre_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  for (void* p : e->dev_allocs) cudaFree(p);
  if (e->tws_mem) cudaFree(e->tws_mem);
  for (auto& kv : e->train_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (e->tg_io) cudaFree(e->tg_io);
  if (e->cap) cudaStreamDestroy(e->cap);
  if (e->side) cudaStreamDestroy(e->side);
  if (e->side2) cudaStreamDestroy(e->side2);
  if (e->ev_join2) cudaEventDestroy(e->ev_join2);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  for (auto& ev : e->ev_pending) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  delete e;
}

int catre_set_weight(catre_engine* e, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!name || !data || !shape) return fail(e, CATRE_ERR_INVALID_ARG, "null argument");
  const WeightSpec* spec = nullptr;
  for (int i = 0; i < kNumWeights; ++i)
    if (strcmp(kWeights[i].name, name) == 0) spec = &kWeights[i];
  if (!spec) return fail(e, CATRE_ERR_UNKNOWN_WEIGHT, "unknown weight '%s'", name);
  if (ndim != spec->ndim) return fail(e, CATRE_ERR_SHAPE, "%s: expected %d dims, got %d", name, spec->ndim, ndim);
  size_t n = 1;
  for (int d = 0; d < ndim; ++d) {
    long long want = spec->shape[d] == -1 ? (long long)(e->cfg.n_obs + e->cfg.n_prior) : spec->shape[d];
    if (shape[d] != want)
      return fail(e, CATRE_ERR_SHAPE, "%s: dim %d is %lld, expected %lld%s", name, d, (long long)shape[d], want,
                  spec->shape[d] == -1 ? " (= n_obs + n_prior: conv_p is tied to the point count)" : "");
    n *= (size_t)want;
  }
  std::vector<float>& h = e->hw[name];
  h.resize(n);
  // `data` may be a device tensor written on any of the caller's streams: wait for the device (set-up path, not hot)
  CU_TRY(e, cudaDeviceSynchronize());
  CU_TRY(e, cudaMemcpy(h.data(), data, n * sizeof(float), cudaMemcpyDefault));
  e->dw_dirty.erase(name);
  e->packed = false;
  return CATRE_OK;
}

int catre_pack(catre_engine* e, void* stream) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  // Ordering contract (include/catre_b200.h): pack overwrites the device weights in place with blocking copies, so
  // everything the caller enqueued before it -- catre_train_set_weight copies on `stream`, kernels still reading the
  // old weights on any stream, non-blocking streams included -- must have finished first.
  (void)stream;
  CU_TRY(e, cudaDeviceSynchronize());
  for (int i = 0; i < kNumWeights; ++i)
    if (!e->hw.count(kWeights[i].name))
      return fail(e, CATRE_ERR_NOT_PACKED, "weight '%s' has not been set", kWeights[i].name);
  // tensors refreshed on the device by catre_train_set_weight: pull them back so the packed copies derive from them
  for (auto& kv : e->dw_dirty) {
    std::vector<float>& h = e->hw.at(kv.first);
    CU_TRY(e, cudaMemcpy(h.data(), e->dw.at(kv.first), h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  }
  e->dw_dirty.clear();
  // packed weights are re-created on every pack; previous device copies stay allocated until destroy
  // only when shapes change (they cannot), so reuse buffers if present
  int rc = 0;
  auto up = [&](float** p, const std::vector<float>& h) {
    if (*p == nullptr) rc |= upload(e, p, h);
    else if (cudaMemcpy(*p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) rc |= 1;
  };
  for (int i = 0; i < kNumWeights; ++i) up(&e->dw[kWeights[i].name], e->hw[kWeights[i].name]);
  auto H = [&](const char* n) -> const std::vector<float>& { return e->hw.at(n); };

  std::vector<float> v, fstn_fc3_wT_host;
  v = H("pcl_net.stn.fc3.bias");  // + I3 (pointnet.py:37-40)
  for (int i = 0; i < 3; ++i) v[i * 3 + i] += 1.0f;
  up(&e->stn_fc3_bI, v);
  {  // fstn.fc3 with output rows permuted (i*64+j -> j*64+i) so the FC emits T64^T, + I64 (pointnet.py:72-77)
    const std::vector<float>& w3 = H("pcl_net.fstn.fc3.weight");
    const std::vector<float>& b3 = H("pcl_net.fstn.fc3.bias");
    std::vector<float>& wT = fstn_fc3_wT_host;
    wT.resize(w3.size());
    std::vector<float> bT(b3.size());
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < 64; ++j) {
        memcpy(&wT[(size_t)(j * 64 + i) * 256], &w3[(size_t)(i * 64 + j) * 256], 256 * sizeof(float));
        bT[j * 64 + i] = b3[i * 64 + j] + (i == j ? 1.0f : 0.0f);
      }
    up(&e->fstn_fc3_wT, wT);
    up(&e->fstn_fc3_bI, bT);
  }

  const char* hx = "rot_head.rot_head_x.";
  const char* hy = "rot_head.rot_head_y.";
  const int P = e->cfg.n_obs + e->cfg.n_prior;
  std::vector<float> w0g(512 * 1024), w0p(512 * 64), b0(512), g0(512), be0(512), b1(512), g1(512), be1(512),
      nw(2 * 3 * 256), nb(6), wp(2 * (size_t)P), cb(2);
  for (int h = 0; h < 2; ++h) {
    std::string pre = h == 0 ? hx : hy;
    const std::vector<float>& w0 = H((pre + "layers.0.weight").c_str());  // [256, 1088] = [global 1024 | pointfeat 64]
    for (int c = 0; c < 256; ++c) {
      memcpy(&w0g[(size_t)(h * 256 + c) * 1024], &w0[(size_t)c * 1088], 1024 * sizeof(float));
      memcpy(&w0p[(size_t)(h * 256 + c) * 64], &w0[(size_t)c * 1088 + 1024], 64 * sizeof(float));
    }
    auto cp = [&](std::vector<float>& dst, const char* n, size_t off) {
      const std::vector<float>& s = H((pre + n).c_str());
      memcpy(&dst[off], s.data(), s.size() * sizeof(float));
    };
    cp(b0, "layers.0.bias", h * 256); cp(g0, "layers.1.weight", h * 256); cp(be0, "layers.1.bias", h * 256);
    cp(b1, "layers.3.bias", h * 256); cp(g1, "layers.4.weight", h * 256); cp(be1, "layers.4.bias", h * 256);
    cp(nw, "neck.0.weight", h * 768); cp(nb, "neck.0.bias", h * 3);
    cp(wp, "conv_p.weight", (size_t)h * P); cp(cb, "conv_p.bias", h);
  }
  up(&e->rot_w0g, w0g); up(&e->rot_w0p, w0p); up(&e->rot_b0, b0);
  up(&e->rot_gn0_g, g0); up(&e->rot_gn0_b, be0); up(&e->rot_b1, b1);
  up(&e->rot_gn1_g, g1); up(&e->rot_gn1_b, be1);
  up(&e->neck_w, nw); up(&e->neck_b, nb); up(&e->wp, wp); up(&e->convp_b, cb);

  const std::vector<float>& t0 = H("ts_head.linears.0.weight");
  std::vector<float> t0t(1091 * 256), t1t(256 * 256);
  for (int c = 0; c < 256; ++c)
    for (int k = 0; k < 1091; ++k) t0t[(size_t)k * 256 + c] = t0[(size_t)c * 1091 + k];
  const std::vector<float>& t1 = H("ts_head.linears.3.weight");
  for (int c = 0; c < 256; ++c)
    for (int k = 0; k < 256; ++k) t1t[(size_t)k * 256 + c] = t1[(size_t)c * 256 + k];
  up(&e->ts_w0t, t0t); up(&e->ts_w1t, t1t);
  {  // ts layer-0 weights over the global feature, [256, 1024] contiguous (operand of the cluster FC)
    std::vector<float> t0g(256 * 1024);
    for (int c = 0; c < 256; ++c) memcpy(&t0g[(size_t)c * 1024], &t0[(size_t)c * 1091], 1024 * sizeof(float));
    up(&e->ts_w0g, t0g);
    // fused FC chains (fc_chain.cuh): [8 ranks][K][C/8] fp32 slices, one contiguous stream per CTA and layer
    auto pack_chain = [&](float** dst, const std::vector<float>& w, int C, int K) {
      const int NC = fcc_nc(C, e->fcc_ranks);
      std::vector<float> pk((size_t)e->fcc_ranks * K * NC);
      fcc_pack(w.data(), C, K, NC, e->fcc_ranks, pk.data());
      up(dst, pk);
    };
    pack_chain(&e->fcc_fc1[0], H("pcl_net.stn.fc1.weight"), 512, 1024);
    pack_chain(&e->fcc_fc2[0], H("pcl_net.stn.fc2.weight"), 256, 512);
    pack_chain(&e->fcc_fc3[0], H("pcl_net.stn.fc3.weight"), 9, 256);
    pack_chain(&e->fcc_fc1[1], H("pcl_net.fstn.fc1.weight"), 512, 1024);
    pack_chain(&e->fcc_fc2[1], H("pcl_net.fstn.fc2.weight"), 256, 512);
    pack_chain(&e->fcc_fc3[1], fstn_fc3_wT_host, 4096, 256);
    pack_chain(&e->fcc_cset, w0g, 512, 1024);
    pack_chain(&e->fcc_ts0, t0g, 256, 1024);
  }
  if (rc) return fail(e, CATRE_ERR_CUDA, "uploading packed weights failed: %s", cudaGetErrorString(cudaGetLastError()));

  if (e->cfg.precision != CATRE_PREC_FP32_SIMT) {
    // bf16 hi/lo split of the wide layers' weights [C, K] + tensor maps (box rows = how the layer uses W)
    auto wpair = [&](TcPair& t, const std::vector<float>& w, int C, int K, int box_rows) -> int {
      std::vector<__nv_bfloat16> hi((size_t)C * K), lo((size_t)C * K);  // 16-bit storage; fp16 patterns in the f16x3 mode
      const bool f16 = e->cfg.precision == CATRE_PREC_F16X3;
      for (size_t i = 0; i < hi.size(); ++i) {
        if (f16) {
          const __half h = __float2half_rn(w[i]);
          const __half l = __float2half_rn(w[i] - __half2float(h));
          hi[i] = __ushort_as_bfloat16(__half_as_ushort(h));
          lo[i] = __ushort_as_bfloat16(__half_as_ushort(l));
        } else {
          hi[i] = __float2bfloat16_rn(w[i]);
          lo[i] = __float2bfloat16_rn(w[i] - __bfloat162float(hi[i]));
        }
      }
      if (!t.hi) {
        if (dalloc(e, &t.hi, hi.size()) || dalloc(e, &t.lo, lo.size())) return CATRE_ERR_CUDA;
      }
      CU_TRY(e, cudaMemcpy(t.hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
      CU_TRY(e, cudaMemcpy(t.lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
      if (!tc_make_map(&t.map_hi, t.hi, C, K, K, box_rows) || !tc_make_map(&t.map_lo, t.lo, C, K, K, box_rows))
        return fail(e, CATRE_ERR_CUDA, "cuTensorMapEncodeTiled failed for a %dx%d weight", C, K);
      return 0;
    };
    int r2 = 0;
    r2 = r2 ? r2 : wpair(e->tw_stn_c2, H("pcl_net.stn.conv2.weight"), 128, 64, 128);
    r2 = r2 ? r2 : wpair(e->tw_stn_c3, H("pcl_net.stn.conv3.weight"), 1024, 128, 128);
    r2 = r2 ? r2 : wpair(e->tw_fstn_c1, H("pcl_net.fstn.conv1.weight"), 64, 64, 64);
    r2 = r2 ? r2 : wpair(e->tw_fstn_c2, H("pcl_net.fstn.conv2.weight"), 128, 64, 128);
    r2 = r2 ? r2 : wpair(e->tw_fstn_c3, H("pcl_net.fstn.conv3.weight"), 1024, 128, 128);
    r2 = r2 ? r2 : wpair(e->tw_conv2, H("pcl_net.conv2.weight"), 128, 64, 128);
    r2 = r2 ? r2 : wpair(e->tw_conv3, H("pcl_net.conv3.weight"), 512, 128, 128);
    r2 = r2 ? r2 : wpair(e->tw_conv4, H("pcl_net.conv4.weight"), 1024, 512, 128);
    r2 = r2 ? r2 : wpair(e->tw_rot0, w0p, 512, 64, 128);
    if (!r2 && (!tc_make_map(&e->tw_rot0_nb[0], e->tw_rot0.hi, 512, 64, 64, 256) ||
                !tc_make_map(&e->tw_rot0_nb[1], e->tw_rot0.lo, 512, 64, 64, 256)))
      r2 = fail(e, CATRE_ERR_CUDA, "cuTensorMapEncodeTiled failed for the rot layer-0 N-side view");
    {
      std::vector<float> w1s = H("rot_head.rot_head_x.layers.3.weight");
      const std::vector<float>& w1y = H("rot_head.rot_head_y.layers.3.weight");
      w1s.insert(w1s.end(), w1y.begin(), w1y.end());
      r2 = r2 ? r2 : wpair(e->tw_rot1s, w1s, 512, 256, 128);
    }
    if (r2) return r2;
  }
  CU_TRY(e, cudaDeviceSynchronize());
  e->packed = true;
  return CATRE_OK;
}

size_t catre_workspace_bytes(const catre_engine* e, int32_t B) {
  if (!e || B < 0) return 0;
  // the workspace is allocated once for max_batch; report the share a launch of B objects touches
  double frac = (double)(B > e->maxB ? e->maxB : B) / (double)e->maxB;
  return (size_t)((double)e->ws_bytes * frac);
}

int catre_forward_once(catre_engine* e, const float* x_pm, const float* kps_pm, const float* pose, const float* scale,
                       const float* K, int32_t B, float* out_pose, float* out_scale, void* stream) {
  int rc = check_ready(e, B);
  if (rc) return rc;
  if (B == 0) { e->launches = 0; return CATRE_OK; }
  if (!x_pm || !kps_pm || !pose || !scale || !K || !out_pose || !out_scale) return fail(e, CATRE_ERR_INVALID_ARG, "null tensor");
  cudaStream_t s = (cudaStream_t)stream;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  e->launches = 0;
  const int N = e->N;
  for (int b0 = 0; b0 < B; b0 += e->maxB) {
    int Bc = (B - b0 < e->maxB) ? B - b0 : e->maxB;
    long long total = (long long)Bc * (N + e->Np);
    {
      Launch l(e, s, G_UPDATE_POINTS);
      launch_pdl(gather_points_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), s, x_pm + (size_t)b0 * N * 3, kps_pm + (size_t)b0 * e->Np * 3,
                 e->q, Bc, N, e->Np, e->gmax_all, (long long)(2 * Bc) * (1024 * 3 + 64));
    }
    if ((rc = check_launch(e, "gather_points"))) return rc;
    if ((rc = iteration(e, s, Bc, pose + (size_t)b0 * 12, scale + (size_t)b0 * 3, K + (size_t)b0 * 9,
                        out_pose + (size_t)b0 * 12, out_scale + (size_t)b0 * 3))) return rc;
  }
  return CATRE_OK;
}

// prior_cls == nullptr: `prior` is [B, N, 3]; otherwise `prior` is the table [n_cls, N, 3] indexed by prior_cls[b]
static int refine_impl(catre_engine* e, const float* pcl, const float* prior, const int32_t* prior_cls, int32_t n_cls,
                       const float* init_pose, const float* init_scale, const float* K, int32_t B, int32_t n_iter,
                       float* out_poses, float* out_scales, void* stream) {
  int rc = check_ready(e, B);
  if (rc) return rc;
  if (n_iter < 0) return fail(e, CATRE_ERR_INVALID_ARG, "negative n_iter %d", n_iter);
  if (B == 0) { e->launches = 0; return CATRE_OK; }
  if (!pcl || !prior || !init_pose || !init_scale || !K || !out_poses || !out_scales) return fail(e, CATRE_ERR_INVALID_ARG, "null tensor");
  cudaStream_t s = (cudaStream_t)stream;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  e->launches = 0;
  const int N = e->N;
  CU_TRY(e, cudaMemcpyAsync(out_poses, init_pose, (size_t)B * 12 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CU_TRY(e, cudaMemcpyAsync(out_scales, init_scale, (size_t)B * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  for (int b0 = 0; b0 < B; b0 += e->maxB) {
    int Bc = (B - b0 < e->maxB) ? B - b0 : e->maxB;
    long long total = (long long)Bc * (N + e->Np);
    const int* pc_chunk = prior_cls ? (const int*)prior_cls + b0 : (const int*)nullptr;
    const bool fused_head = e->cfg.precision != CATRE_PREC_FP32_SIMT;
    TsPoseP prev{};
    for (int it = 1; it <= n_iter; ++it) {
      const float* pin = out_poses + ((size_t)(it - 1) * B + b0) * 12;
      const float* sin = out_scales + ((size_t)(it - 1) * B + b0) * 3;
      float* pout = out_poses + ((size_t)it * B + b0) * 12;
      float* sout = out_scales + ((size_t)it * B + b0) * 3;
      if (fused_head) {
        // tensor-core modes: the previous iteration's pose update, the point update and stn.conv1 are one launch; only the
        // last iteration ends with pose_update_kernel
        const float* pr = prior_cls ? prior : prior + (size_t)b0 * e->Np * 3;
        if ((rc = iter_head(e, s, Bc, it > 1 ? &prev : nullptr, pin, sin, pcl + (size_t)b0 * N * 3, pr, pc_chunk, (int)n_cls))) return rc;
        if ((rc = iteration(e, s, Bc, pin, sin, K + (size_t)b0 * 9, pout, sout, pc_chunk, (int)n_cls, true, it < n_iter ? &prev : nullptr)))
          return rc;
        continue;
      }
      {
        Launch l(e, s, G_UPDATE_POINTS);
        const float* pr = prior_cls ? prior : prior + (size_t)b0 * e->Np * 3;
        launch_pdl(update_points_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), s, pcl + (size_t)b0 * N * 3, pr,
                   pin, sin, e->q, Bc, N, e->Np, e->gmax_all, (long long)(2 * Bc) * (1024 * 3 + 64), pc_chunk, (int)n_cls);
      }
      if ((rc = check_launch(e, "update_points"))) return rc;
      if ((rc = iteration(e, s, Bc, pin, sin, K + (size_t)b0 * 9, pout, sout, pc_chunk, (int)n_cls))) return rc;
    }
  }
  return CATRE_OK;
}

int catre_refine(catre_engine* e, const float* pcl, const float* prior, const float* init_pose, const float* init_scale,
                 const float* K, int32_t B, int32_t n_iter, float* out_poses, float* out_scales, void* stream) {
  return refine_impl(e, pcl, prior, nullptr, 0, init_pose, init_scale, K, B, n_iter, out_poses, out_scales, stream);
}

int catre_refine_table(catre_engine* e, const float* pcl, const float* prior_table, const int32_t* prior_cls, int32_t n_cls,
                       const float* init_pose, const float* init_scale, const float* K, int32_t B, int32_t n_iter,
                       float* out_poses, float* out_scales, void* stream) {
  if (e && (n_cls < 1 || (!prior_cls && B > 0))) return fail(e, CATRE_ERR_INVALID_ARG, "prior table needs n_cls >= 1 and class ids");
  return refine_impl(e, pcl, prior_table, prior_cls, n_cls, init_pose, init_scale, K, B, n_iter, out_poses, out_scales, stream);
}

static int refine_host_impl(catre_engine* e, const float* pcl, const float* prior, const int32_t* prior_cls, int32_t n_cls,
                            const float* init_pose, const float* init_scale, const float* K, int32_t B, int32_t n_iter,
                            float* out_poses, float* out_scales, void* stream, float* packed_dev = nullptr) {
  int rc = check_ready(e, B);
  if (rc) return rc;
  if (n_iter < 0 || n_iter > catre_engine::kMaxHostIter)
    return fail(e, CATRE_ERR_INVALID_ARG, "n_iter %d outside [0, %d] for the host entry", n_iter, catre_engine::kMaxHostIter);
  if (B == 0) { e->launches = 0; return CATRE_OK; }
  if (!pcl || !prior || !init_pose || !init_scale || !K || !out_poses || !out_scales) return fail(e, CATRE_ERR_INVALID_ARG, "null tensor");
  cudaStream_t s = (cudaStream_t)stream;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  const int N = e->N;
  int64_t launches = 0;
  if (prior_cls) {  // the table travels once per call, the class ids per chunk
    if (n_cls < 1 || n_cls > e->st_prior_rows)
      return fail(e, CATRE_ERR_INVALID_ARG, "n_cls %d outside [1, %d] for the host entry", n_cls, e->st_prior_rows);
    for (int b = 0; b < B; ++b)
      if (prior_cls[b] < 0 || prior_cls[b] >= n_cls)
        return fail(e, CATRE_ERR_INVALID_ARG, "prior_cls[%d] = %d outside [0, %d)", b, prior_cls[b], n_cls);
    CU_TRY(e, cudaMemcpyAsync(e->st_prior, prior, (size_t)n_cls * e->Np * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  for (int b0 = 0; b0 < B; b0 += e->maxB) {
    int Bc = (B - b0 < e->maxB) ? B - b0 : e->maxB;
    CU_TRY(e, cudaMemcpyAsync(e->st_pcl, pcl + (size_t)b0 * N * 3, (size_t)Bc * N * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (prior_cls) CU_TRY(e, cudaMemcpyAsync(e->st_cls, prior_cls + b0, (size_t)Bc * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    else CU_TRY(e, cudaMemcpyAsync(e->st_prior, prior + (size_t)b0 * e->Np * 3, (size_t)Bc * e->Np * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CU_TRY(e, cudaMemcpyAsync(e->st_pose, init_pose + (size_t)b0 * 12, (size_t)Bc * 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    CU_TRY(e, cudaMemcpyAsync(e->st_scale, init_scale + (size_t)b0 * 3, (size_t)Bc * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CU_TRY(e, cudaMemcpyAsync(e->st_K, K + (size_t)b0 * 9, (size_t)Bc * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = refine_impl(e, e->st_pcl, e->st_prior, prior_cls ? e->st_cls : nullptr, n_cls, e->st_pose, e->st_scale, e->st_K, Bc, n_iter,
                     e->st_oposes, e->st_oscales, s);
    if (rc) return rc;
    launches += e->launches;
    if (packed_dev) {  // the final poses of this chunk, packed for the multi-GPU all-gather, stay on the device
      launch_pdl(pack_poses_kernel, dim3((unsigned)((Bc * 15 + 255) / 256)), dim3(256), (size_t)0, s, (const float*)e->st_oposes,
                 (const float*)e->st_oscales, (int)Bc, (int)n_iter, packed_dev + (size_t)b0 * 15);
      if ((rc = check_launch(e, "pack_poses"))) return rc;
      ++launches;
    }
    if (Bc == B) {  // single chunk: the staging layout [n_iter+1, B, .] is the output layout
      CU_TRY(e, cudaMemcpyAsync(out_poses, e->st_oposes, (size_t)(n_iter + 1) * B * 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
      CU_TRY(e, cudaMemcpyAsync(out_scales, e->st_oscales, (size_t)(n_iter + 1) * B * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    } else {
      for (int it = 0; it <= n_iter; ++it) {
        CU_TRY(e, cudaMemcpyAsync(out_poses + ((size_t)it * B + b0) * 12, e->st_oposes + (size_t)it * Bc * 12,
                                  (size_t)Bc * 12 * sizeof(float), cudaMemcpyDeviceToHost, s));
        CU_TRY(e, cudaMemcpyAsync(out_scales + ((size_t)it * B + b0) * 3, e->st_oscales + (size_t)it * Bc * 3,
                                  (size_t)Bc * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
      }
    }
  }
  CU_TRY(e, cudaStreamSynchronize(s));
  e->launches = launches;
  return CATRE_OK;
}

int catre_refine_host(catre_engine* e, const float* pcl, const float* prior, const float* init_pose, const float* init_scale,
                      const float* K, int32_t B, int32_t n_iter, float* out_poses, float* out_scales, void* stream) {
  return refine_host_impl(e, pcl, prior, nullptr, 0, init_pose, init_scale, K, B, n_iter, out_poses, out_scales, stream);
}

int catre_refine_host_packed(catre_engine* e, const float* pcl, const float* prior, const float* init_pose, const float* init_scale,
                             const float* K, int32_t B, int32_t n_iter, float* out_poses, float* out_scales, float* packed_dev,
                             void* stream) {
  if (e && !packed_dev && B > 0) return fail(e, CATRE_ERR_INVALID_ARG, "catre_refine_host_packed: null packed_dev");
  return refine_host_impl(e, pcl, prior, nullptr, 0, init_pose, init_scale, K, B, n_iter, out_poses, out_scales, stream, packed_dev);
}

int catre_pack_poses(const float* poses, const float* scales, int32_t B, int32_t iter, float* packed, void* stream) {
  if (B == 0) return CATRE_OK;
  if (!poses || !scales || !packed || B < 0 || iter < 0) return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_pack_poses: bad argument");
  launch_pdl(pack_poses_kernel, dim3((unsigned)((B * 15 + 255) / 256)), dim3(256), (size_t)0, (cudaStream_t)stream, poses, scales, (int)B,
             (int)iter, packed);
  CU_TRY(nullptr, cudaGetLastError());
  return CATRE_OK;
}

int catre_refine_table_host(catre_engine* e, const float* pcl, const float* prior_table, const int32_t* prior_cls, int32_t n_cls,
                            const float* init_pose, const float* init_scale, const float* K, int32_t B, int32_t n_iter,
                            float* out_poses, float* out_scales, void* stream) {
  if (e && (n_cls < 1 || (!prior_cls && B > 0))) return fail(e, CATRE_ERR_INVALID_ARG, "prior table needs n_cls >= 1 and class ids");
  return refine_host_impl(e, pcl, prior_table, prior_cls, n_cls, init_pose, init_scale, K, B, n_iter, out_poses, out_scales, stream);
}

// ---- observed-cloud producer (cloud_kernels.cuh): engine-independent, errors go to catre_last_error(NULL) ----
namespace {
struct CloudScratch { int* hist; int* chunk_off; int* kstar; };
inline int cloud_chunks(long long hw) { return (int)((hw + CLOUD_CHUNK - 1) / CLOUD_CHUNK); }
inline CloudScratch cloud_scratch(void* scratch, int B, int n_chunks) {
  CloudScratch c;
  c.hist = reinterpret_cast<int*>(scratch);
  c.chunk_off = c.hist + (size_t)B * n_chunks * CLOUD_BINS;
  c.kstar = c.chunk_off + (size_t)B * n_chunks;
  return c;
}
}  // namespace

size_t catre_cloud_scratch_bytes(int32_t B, int32_t H, int32_t W) {
  if (B < 0 || H <= 0 || W <= 0) return 0;
  const size_t nc = (size_t)cloud_chunks((long long)H * W);
  return ((size_t)B * nc * (CLOUD_BINS + 1) + (size_t)B + 16) * sizeof(int);
}

int catre_cloud_select(const float* depth, const uint8_t* masks, const float* intr, const float* centers, const float* radii,
                       int32_t n_radii, int32_t B, int32_t H, int32_t W, int32_t* sel_pix, int32_t* n_sel, void* scratch,
                       void* stream) {
  if (B == 0) return CATRE_OK;
  if (!depth || !masks || !intr || !centers || !radii || !sel_pix || !n_sel || !scratch)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_cloud_select: null argument");
  if (B < 0 || H <= 0 || W <= 0 || (long long)H * W > (1ll << 30))
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_cloud_select: bad sizes B=%d H=%d W=%d", B, H, W);
  if (n_radii < 1 || n_radii > CLOUD_MAX_RADII)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_cloud_select: n_radii=%d outside [1, %d]", n_radii, CLOUD_MAX_RADII);
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = H * W, nc = cloud_chunks(HW);
  const CloudIntr k{intr[0], intr[1], intr[2], intr[3]};
  const CloudScratch c = cloud_scratch(scratch, B, nc);
  cloud_hist_kernel<<<dim3(nc, B), CLOUD_THREADS, 0, s>>>(depth, masks, k, centers, radii, n_radii, HW, W, nc, c.hist);
  cloud_select_kernel<<<B, 32, 0, s>>>(c.hist, nc, n_radii, c.kstar, n_sel, c.chunk_off);
  cloud_compact_kernel<<<dim3(nc, B), CLOUD_THREADS, 0, s>>>(depth, masks, k, centers, radii, n_radii, HW, W, nc, c.kstar,
                                                               c.chunk_off, sel_pix);
  CU_TRY(nullptr, cudaGetLastError());
  return CATRE_OK;
}

int catre_cloud_gather(const float* depth, const float* intr, const int32_t* sel_pix, const int32_t* n_sel, const int64_t* sample_idx,
                       int32_t B, int32_t H, int32_t W, int32_t n_pts, float* pcl, void* stream) {
  if (B == 0 || n_pts == 0) return CATRE_OK;
  if (!depth || !intr || !sel_pix || !n_sel || !sample_idx || !pcl)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_cloud_gather: null argument");
  if (B < 0 || H <= 0 || W <= 0 || n_pts < 0)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_cloud_gather: bad sizes B=%d H=%d W=%d n_pts=%d", B, H, W, n_pts);
  const CloudIntr k{intr[0], intr[1], intr[2], intr[3]};
  cloud_gather_kernel<<<dim3((n_pts + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(
      depth, k, sel_pix, n_sel, reinterpret_cast<const long long*>(sample_idx), H * W, W, n_pts, pcl);
  CU_TRY(nullptr, cudaGetLastError());
  return CATRE_OK;
}

// ---- pairwise NOCS pose metrics (metrics_kernels.cuh): engine-independent ----
int catre_pair_metrics_ex(const double* pred_RT, const double* pred_scale, const int32_t* pred_cls, const double* gt_RT,
                          const double* gt_scale, const int32_t* gt_cls, const int32_t* gt_handle, const int32_t* pair_pred,
                          const int32_t* pair_gt, int32_t n_pairs, uint32_t sym_class_mask, uint32_t flip_class_mask,
                          int32_t mug_class, int32_t shift_mode, float* iou, float* deg_shift, double* deg_shift64, void* stream) {
  if (n_pairs == 0) return CATRE_OK;
  if (n_pairs < 0) return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_pair_metrics: negative n_pairs %d", n_pairs);
  if (shift_mode != 0 && shift_mode != 1) return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_pair_metrics: shift_mode %d", shift_mode);
  if (!pred_RT || !pred_scale || !pred_cls || !gt_RT || !gt_scale || !gt_cls || !gt_handle || !pair_pred || !pair_gt || !iou ||
      (!deg_shift && !deg_shift64))
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_pair_metrics: null argument");
  PairMetricsP p{pred_RT, pred_scale, pred_cls, gt_RT, gt_scale, gt_cls, gt_handle, pair_pred, pair_gt, n_pairs,
                 sym_class_mask, flip_class_mask, mug_class, iou, deg_shift, shift_mode, deg_shift64};
  pair_metrics_kernel<<<(n_pairs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  CU_TRY(nullptr, cudaGetLastError());
  return CATRE_OK;
}

int catre_pair_metrics(const double* pred_RT, const double* pred_scale, const int32_t* pred_cls, const double* gt_RT,
                       const double* gt_scale, const int32_t* gt_cls, const int32_t* gt_handle, const int32_t* pair_pred,
                       const int32_t* pair_gt, int32_t n_pairs, uint32_t sym_class_mask, uint32_t flip_class_mask,
                       int32_t mug_class, float* iou, float* deg_shift, void* stream) {
  if (n_pairs > 0 && !deg_shift) return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_pair_metrics: null argument");
  return catre_pair_metrics_ex(pred_RT, pred_scale, pred_cls, gt_RT, gt_scale, gt_cls, gt_handle, pair_pred, pair_gt, n_pairs,
                               sym_class_mask, flip_class_mask, mug_class, 0, iou, deg_shift, nullptr, stream);
}

int catre_match_greedy(int32_t mode, const int32_t* sub_pred_off, const int32_t* sub_gt_off, const int32_t* sub_pair_off,
                       int32_t n_sub, int32_t n_pred, int32_t n_gt, const float* iou, const double* deg_shift64,
                       const int32_t* order, const int32_t* n_cand, const int32_t* pred_cls, const int32_t* gt_cls,
                       const double* thr_a, int32_t n_a, const double* thr_b, int32_t n_b, int32_t* gt_match,
                       int32_t* pred_match, void* stream) {
  if (n_sub == 0 || n_a * n_b == 0) return CATRE_OK;
  if (mode != 0 && mode != 1) return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_match_greedy: mode %d", mode);
  if (n_sub < 0 || n_pred < 0 || n_gt < 0 || n_a < 1 || n_b < 1)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_match_greedy: bad sizes");
  if (!sub_pred_off || !sub_gt_off || !sub_pair_off || !thr_a || (mode == 1 && !thr_b) || !gt_match || !pred_match ||
      (mode == 0 ? !iou : !deg_shift64) || !order || !n_cand || !pred_cls || !gt_cls)
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_match_greedy: null argument");
  MatchP p{sub_pred_off, sub_gt_off, sub_pair_off, iou, deg_shift64, order, n_cand, pred_cls, gt_cls, thr_a, n_a,
           thr_b, n_b, gt_match, pred_match, n_sub, n_pred, n_gt, mode};
  const long long threads = (long long)n_sub * n_a * n_b;
  match_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
  CU_TRY(nullptr, cudaGetLastError());
  return CATRE_OK;
}

// ---- training step ---------------------------------------------------------------------
int catre_train_set_weight(catre_engine* e, const char* name, const float* src_dev, void* stream) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!name || !src_dev) return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_set_weight: null argument");
  const int i = weight_index(name);
  if (i < 0) return fail(e, CATRE_ERR_UNKNOWN_WEIGHT, "unknown weight '%s'", name);
  auto it = e->dw.find(name);
  if (it == e->dw.end() || !it->second)
    return fail(e, CATRE_ERR_NOT_PACKED, "catre_train_set_weight needs one earlier catre_pack (it allocates the device copies)");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  CU_TRY(e, cudaMemcpyAsync(it->second, src_dev, catre_train::weight_numel(i, e->N) * sizeof(float), cudaMemcpyDefault,
                            (cudaStream_t)stream));
  e->dw_dirty[name] = true;
  e->packed = false;
  return CATRE_OK;
}

int catre_train_set_weights(catre_engine* e, const float* const* src_dev, void* stream) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!src_dev) return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_set_weights: null argument");
  static_assert(kNumWeights <= catre_train::KMultiCopy::kMax, "pointer table of KMultiCopy");
  if (e->dw.size() != (size_t)kNumWeights)
    return fail(e, CATRE_ERR_NOT_PACKED, "catre_train_set_weights needs one earlier catre_pack (it allocates the device copies)");
  catre_train::KMultiCopy k{};
  int n_max = 0, n_set = 0;
  for (int i = 0; i < kNumWeights; ++i) {
    k.src[i] = src_dev[i];
    k.dst[i] = e->dw.at(kWeights[i].name);
    k.n[i] = (int)catre_train::weight_numel(i, e->N);
    if (!src_dev[i]) continue;
    ++n_set;
    if (k.n[i] > n_max) n_max = k.n[i];
    e->dw_dirty[kWeights[i].name] = true;
  }
  if (n_set == 0) return CATRE_OK;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  k.per = 8192;
  catre_train::tk_run<<<dim3((unsigned)((n_max + k.per - 1) / k.per), (unsigned)kNumWeights), 256, 0, (cudaStream_t)stream>>>(k);
  CU_TRY(e, cudaGetLastError());
  e->packed = false;
  return CATRE_OK;
}

int catre_train_set_loss_weights(catre_engine* e, float pm_lw, float rot_lw, float trans_lw, float scale_lw) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  const float w[4] = {pm_lw, rot_lw, trans_lw, scale_lw};
  for (float v : w)
    if (!(v > 0.0f) || !isfinite(v))
      return fail(e, CATRE_ERR_UNSUPPORTED, "loss weights must be finite and > 0 (got %g %g %g %g): a zero weight removes the term from "
                  "the reference's loss dict, which this engine does not implement", pm_lw, rot_lw, trans_lw, scale_lw);
  for (int i = 0; i < 4; ++i) e->loss_w[i] = w[i];
  return CATRE_OK;
}

int catre_train_step(catre_engine* e, const float* x_pm, const float* tfd_pm, const float* obj_kps, const float* pose,
                     const float* scale, const float* K, const float* gt_pose, const float* gt_scale,
                     const uint8_t* is_sym_host, const float* sym_rots_host, int32_t n_sym_rots, int32_t B, float* out_pose,
                     float* out_scale, float* out_losses, void* stream) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (B == 0) return CATRE_OK;
  if (B < 0) return fail(e, CATRE_ERR_INVALID_ARG, "negative batch %d", B);
  if (!x_pm || !tfd_pm || !obj_kps || !pose || !scale || !K || !gt_pose || !gt_scale || !is_sym_host || !out_pose || !out_scale ||
      !out_losses)
    return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_step: null argument");
  if (n_sym_rots < 0 || n_sym_rots > catre_train::TrainWs::kMaxSymRots || (n_sym_rots > 0 && !sym_rots_host))
    return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_step: n_sym_rots %d outside [0, %d] or null rotations", n_sym_rots,
                catre_train::TrainWs::kMaxSymRots);
  if (e->dw.size() != (size_t)kNumWeights)
    return fail(e, CATRE_ERR_NOT_PACKED, "catre_train_step needs the 74 tensors set and one catre_pack");
  if (e->N != e->Np)
    return fail(e, CATRE_ERR_UNSUPPORTED, "catre_train_step: the training chain needs n_obs == n_prior (got %d, %d)", e->N, e->Np);
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (B > e->tws_maxB) {  // (re)allocate the workspace; not on the steady-state path
    CU_TRY(e, cudaStreamSynchronize(s));
    for (auto& kv : e->train_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);  // they hold the old pointers
    e->train_graphs.clear();
    if (e->tws_mem) { CU_TRY(e, cudaFree(e->tws_mem)); e->tws_mem = nullptr; e->tws_maxB = 0; }
    if (e->tg_io) { CU_TRY(e, cudaFree(e->tg_io)); e->tg_io = nullptr; }
    catre_train::TrainWs probe;
    const size_t bytes = catre_train::ws_layout(probe, B, e->N, nullptr);
    void* mem = nullptr;
    cudaError_t st = cudaMalloc(&mem, bytes);
    if (st != cudaSuccess) {
      cudaGetLastError();
      return fail(e, CATRE_ERR_CUDA, "training workspace for %d objects (%.1f MB): %s", B, bytes / 1048576.0, cudaGetErrorString(st));
    }
    e->tws_mem = static_cast<char*>(mem);
    catre_train::ws_layout(e->tws, B, e->N, e->tws_mem);
    e->tws_maxB = B;
    CU_TRY(e, cudaMalloc(&mem, ((size_t)9 * B * e->N + (size_t)64 * B) * sizeof(float)));
    e->tg_io = static_cast<float*>(mem);
  }
  catre_train::TrainWs& w = e->tws;
  int n_sym = 0;
  for (int b = 0; b < B; ++b) n_sym += is_sym_host[b] != 0;
  if (n_sym > 0 && n_sym_rots == 0)
    return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_step: symmetric objects but no symmetry rotations");
  CU_TRY(e, cudaMemcpyAsync(w.is_sym, is_sym_host, (size_t)B, cudaMemcpyHostToDevice, s));
  if (n_sym_rots) CU_TRY(e, cudaMemcpyAsync(w.sym_rots, sym_rots_host, (size_t)n_sym_rots * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
  const float* Wp[catre_train::W_COUNT];
  for (int i = 0; i < kNumWeights; ++i) Wp[i] = e->dw.at(kWeights[i].name);
  const char* env = getenv("CATRE_TRAIN_NAIVE_GEMM");
  const char* env2 = getenv("CATRE_TRAIN_GEMM");
  const char* env3 = getenv("CATRE_TRAIN_GRAPH");
  const bool naive_gemm = e->train_naive_gemm || (env && env[0] == '1'), gemm_v2 = env2 && strcmp(env2, "v2") == 0;
  const bool gemm_tc = !naive_gemm && !gemm_v2 && !(env2 && strcmp(env2, "simt") == 0);
  const char* env4 = getenv("CATRE_TRAIN_CARVEOUT");
  const bool carve = gemm_tc && env4 && env4[0] == '1';  // experiment (see CudaTrainOps::carve)
  const char* env5 = getenv("CATRE_TRAIN_FOLD_BIAS");
  const bool fold_bg = !(env5 && env5[0] == '0');
  const char* env6 = getenv("CATRE_TRAIN_LANES");
  const bool lanes = !(env6 && env6[0] == '0');
  auto make_in = [&](const float* x, const float* tfd, const float* kps, const float* po, const float* sc, const float* Kz,
                     const float* gp, const float* gs, float* op, float* os) {
    catre_train::TrainIn in{nullptr, kps, po, sc, Kz, gp, gs, B, n_sym_rots, n_sym, B - n_sym, op, os};
    in.x_pm = x; in.tfd_pm = tfd;
    in.w_pm = e->loss_w[0]; in.w_rot = e->loss_w[1]; in.w_trans = e->loss_w[2]; in.w_scale = e->loss_w[3];
    return in;
  };
  auto run_chain = [&](cudaStream_t st, const catre_train::TrainIn& in, int64_t& n_launch) -> cudaError_t {
    CudaTrainOps ops{st, naive_gemm, gemm_v2, gemm_tc, e->num_sms};
    ops.carve = carve;
    ops.fold_bg = fold_bg;
    ops.lanes = lanes && e->side && e->side2 && e->ev_fork && e->ev_join && e->ev_join2;
    ops.s_main = st; ops.s_side[0] = e->side; ops.s_side[1] = e->side2; ops.ev_fork = e->ev_fork; ops.ev_join[0] = e->ev_join;
    ops.ev_join[1] = e->ev_join2;
    catre_train::Chain<CudaTrainOps> chain{ops, w, Wp, e->N};
    chain.forward(in);
    chain.loss(in);
    chain.backward(in);
    n_launch = ops.launches;
    return ops.err;
  };
  cudaError_t cerr = cudaSuccess;
  bool done = false;
  if (!(env3 && env3[0] == '0')) {
    catre_engine::TrainGraphKey key{};
    key.B = B; key.n_sym = n_sym; key.n_rots = n_sym_rots; key.mode = (naive_gemm ? 1 : (gemm_v2 ? 2 : (gemm_tc ? 0 : 3))) + (carve ? 8 : 0) + (fold_bg ? 0 : 16) + (lanes ? 0 : 32);
    for (int i = 0; i < 4; ++i) key.lw[i] = e->loss_w[i];
    if (e->train_graphs.size() > 64 && !e->train_graphs.count(key)) {  // bound the cache (each graph holds ~240 nodes)
      CU_TRY(e, cudaStreamSynchronize(s));
      for (auto& kv : e->train_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      e->train_graphs.clear();
    }
    catre_engine::TrainGraph& g = e->train_graphs[key];
    if (g.seen > 0) {
      const size_t pts = (size_t)B * e->N * 3;
      float* io = e->tg_io;
      float *sx = io, *st = sx + pts, *sk = st + pts, *spo = sk + pts, *ssc = spo + 12 * B, *sK = ssc + 4 * B, *sgp = sK + 12 * B,
            *sgs = sgp + 12 * B, *sop = sgs + 4 * B, *sos = sop + 12 * B;  // every slice a multiple of 4 floats: 16-byte aligned
      if (!g.exec) {
        if (!e->cap) CU_TRY(e, cudaStreamCreateWithFlags(&e->cap, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        CU_TRY(e, cudaStreamBeginCapture(e->cap, cudaStreamCaptureModeThreadLocal));
        cerr = run_chain(e->cap, make_in(sx, st, sk, spo, ssc, sK, sgp, sgs, sop, sos), g.launches);
        cudaError_t eend = cudaStreamEndCapture(e->cap, &graph);
        if (cerr == cudaSuccess) cerr = eend;
        if (cerr == cudaSuccess) cerr = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (cerr != cudaSuccess) {
          cudaGetLastError();
          g.exec = nullptr;
          return fail(e, CATRE_ERR_CUDA, "catre_train_step: capturing the chain as a CUDA graph: %s (CATRE_TRAIN_GRAPH=0 launches it "
                      "kernel by kernel)", cudaGetErrorString(cerr));
        }
      }
      const struct { float* dst; const float* src; size_t n; } cp[8] = {
          {sx, x_pm, pts}, {st, tfd_pm, pts}, {sk, obj_kps, pts}, {spo, pose, (size_t)12 * B}, {ssc, scale, (size_t)3 * B},
          {sK, K, (size_t)9 * B}, {sgp, gt_pose, (size_t)12 * B}, {sgs, gt_scale, (size_t)3 * B}};
      for (const auto& c : cp) CU_TRY(e, cudaMemcpyAsync(c.dst, c.src, c.n * sizeof(float), cudaMemcpyDeviceToDevice, s));
      CU_TRY(e, cudaGraphLaunch(g.exec, s));
      CU_TRY(e, cudaMemcpyAsync(out_pose, sop, (size_t)12 * B * sizeof(float), cudaMemcpyDeviceToDevice, s));
      CU_TRY(e, cudaMemcpyAsync(out_scale, sos, (size_t)3 * B * sizeof(float), cudaMemcpyDeviceToDevice, s));
      e->launches = g.launches;
      done = true;
    } else {
      g.seen = 1;  // this step runs kernel by kernel (and sets the kernels' attributes); the next one of this key is captured
    }
  }
  if (!done) cerr = run_chain(s, make_in(x_pm, tfd_pm, obj_kps, pose, scale, K, gt_pose, gt_scale, out_pose, out_scale), e->launches);
  if (cerr == cudaSuccess) cerr = cudaMemcpyAsync(out_losses, w.losses, 6 * sizeof(float), cudaMemcpyDefault, s);
  if (cerr != cudaSuccess) {
    cudaGetLastError();
    return fail(e, CATRE_ERR_CUDA, "catre_train_step: %s", cudaGetErrorString(cerr));
  }
  return CATRE_OK;
}

int catre_train_grad(catre_engine* e, const char* name, float* dst, void* stream) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!name || !dst) return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_grad: null argument");
  const int i = weight_index(name);
  if (i < 0) return fail(e, CATRE_ERR_UNKNOWN_WEIGHT, "unknown weight '%s'", name);
  if (!e->tws_mem) return fail(e, CATRE_ERR_NOT_PACKED, "catre_train_grad before any catre_train_step");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  CU_TRY(e, cudaMemcpyAsync(dst, e->tws.G[i], catre_train::weight_numel(i, e->N) * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  return CATRE_OK;
}

int catre_train_grad_layout(catre_engine* e, int64_t* offsets, int64_t* total_floats) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!offsets || !total_floats) return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_grad_layout: null argument");
  catre_train::TrainWs probe;
  catre_train::ws_layout(probe, 1, e->N, reinterpret_cast<char*>(4096));  // dummy base: only pointer differences are used
  for (int i = 0; i < kNumWeights; ++i) offsets[i] = (int64_t)(probe.G[i] - probe.G[0]);
  *total_floats = (int64_t)probe.grad_floats;
  return CATRE_OK;
}

int catre_train_grads_flat(catre_engine* e, float* dst, float scale, void* stream) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  if (!dst) return fail(e, CATRE_ERR_INVALID_ARG, "catre_train_grads_flat: null argument");
  if (!e->tws_mem) return fail(e, CATRE_ERR_NOT_PACKED, "catre_train_grads_flat before any catre_train_step");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)e->tws.grad_floats;
  if (scale == 1.0f) {
    CU_TRY(e, cudaMemcpyAsync(dst, e->tws.G[0], (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    catre_train::KScaleCopy k{e->tws.G[0], dst, scale, n};
    catre_train::tk_run<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(k);
    CU_TRY(e, cudaGetLastError());
  }
  return CATRE_OK;
}

int catre_ranger_step(const int64_t* table, const float* lr_wd, const int64_t* elem_start, const int64_t* row_start,
                      int32_t n_tensors, int64_t total_elems, int64_t total_rows, float* rowmean, const catre_ranger_args* a,
                      void* stream) {
  static_assert(sizeof(catre_train::RangerTensor) == 64 && sizeof(catre_train::RangerArgs) == sizeof(catre_ranger_args),
                "table row / argument layout of include/catre_b200.h");
  if (n_tensors == 0 || total_elems == 0) return CATRE_OK;
  if (!table || !lr_wd || !elem_start || !row_start || !a || n_tensors < 0 || total_elems < 0 || total_rows < 0 ||
      (total_rows > 0 && !rowmean))
    return fail(nullptr, CATRE_ERR_INVALID_ARG, "catre_ranger_step: null or negative argument");
  cudaStream_t s = (cudaStream_t)stream;
  const catre_train::RangerTensor* T = reinterpret_cast<const catre_train::RangerTensor*>(table);
  catre_train::RangerArgs ra{a->beta1, a->beta2, a->eps, a->one_minus_beta1, a->one_minus_beta2, a->step_size, a->rectified, a->alpha, a->lookahead, a->nan_to_num};
  if (total_rows > 0) {
    catre_train::KRangerRowMean k{T, reinterpret_cast<const long long*>(row_start), n_tensors, (long long)total_rows, rowmean, a->nan_to_num};
    catre_train::tk_run<<<(unsigned)((total_rows + 255) / 256), 256, 0, s>>>(k);
  }
  catre_train::KRangerUpdate k{T, reinterpret_cast<const long long*>(elem_start), reinterpret_cast<const long long*>(row_start), lr_wd,
                               rowmean, n_tensors, (long long)total_elems, ra};
  catre_train::tk_run<<<(unsigned)((total_elems + 255) / 256), 256, 0, s>>>(k);
  CU_TRY(nullptr, cudaGetLastError());
  return CATRE_OK;
}

int64_t catre_last_launch_count(const catre_engine* e) { return e ? e->launches : 0; }

int catre_debug_train_gemm(const float* A, const float* B, float* C, const float* bias, const int64_t* st, int32_t M, int32_t N,
                           int32_t K, int32_t batch, int32_t relu, int32_t accumulate, int32_t splits, float* partial,
                           int32_t kernel, int32_t rows_per_set, float* vmax, int32_t* arg, void* stream) {
  if (!A || !B || !st || M <= 0 || N <= 0 || K <= 0 || batch <= 0 || kernel < 0 || kernel > 2) return CATRE_ERR_INVALID_ARG;
  if (splits > 1 && (batch != 1 || !partial || bias || relu)) return CATRE_ERR_INVALID_ARG;
  const bool colmax = vmax != nullptr;
  if (colmax ? (!arg || !partial || kernel == 0 || splits > 1 || batch != 1 || rows_per_set <= 0 || rows_per_set % 128 != 0 || M % rows_per_set != 0)
             : !C)
    return CATRE_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  catre_train::GemmP p{};
  p.A = A; p.sam = st[0]; p.sak = st[1]; p.sab = st[2]; p.B = B; p.sbk = st[3]; p.sbn = st[4]; p.sbb = st[5];
  p.C = C; p.scm = st[6]; p.scn = st[7]; p.scb = st[8]; p.bias = bias; p.sbias_b = 0;
  p.M = M; p.N = N; p.K = K; p.relu = relu; p.accumulate = accumulate; p.splits = 1; p.k_per = K; p.partial = partial;
  p.f16 = kernel == 1;
  int bz = batch;
  if (splits > 1) { p.splits = splits; p.k_per = ((K + splits - 1) / splits + 63) & ~63; bz = splits; }
  cudaError_t err;
  if (colmax) {
    const int sets = M / rows_per_set;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(partial);
    err = cudaMemsetAsync(keys, 0, (size_t)sets * N * sizeof(unsigned long long), s);
    const catre_train::TgColMax cm{keys, rows_per_set};
    if (err == cudaSuccess) err = p.f16 ? catre_train::tk_gemm_tc_launch<true>(p, 1, s, cm) : catre_train::tk_gemm_tc_launch<false>(p, 1, s, cm);
    if (err == cudaSuccess) {
      const long long n = (long long)sets * N;
      catre_train::tk_run<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(catre_train::KColMaxDecode{keys, vmax, arg, n});
      err = cudaPeekAtLastError();
    }
  } else if (kernel == 0) {
    catre_train::tk_gemm_tiled<<<dim3((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)bz), 256, 0, s>>>(p);
    err = cudaPeekAtLastError();
  } else {
    err = p.f16 ? catre_train::tk_gemm_tc_launch<true>(p, bz, s) : catre_train::tk_gemm_tc_launch<false>(p, bz, s);
  }
  if (err == cudaSuccess && splits > 1) {
    catre_train::KSplitReduce red{partial, C, p.scm, p.scn, M, N, splits, accumulate};
    catre_train::tk_run<<<(unsigned)(((long long)M * N + 255) / 256), 256, 0, s>>>(red);
    err = cudaPeekAtLastError();
  }
  if (err != cudaSuccess) { cudaGetLastError(); return CATRE_ERR_CUDA; }
  return CATRE_OK;
}

int catre_debug_read(catre_engine* e, const char* name, void* dst_host, size_t bytes) {
  if (!e || !name || !dst_host) return CATRE_ERR_INVALID_ARG;
  std::map<std::string, const void*> m = {
      {"q", e->q}, {"h64a", e->h64a}, {"h64b", e->h64b}, {"h128", e->h128}, {"h512", e->h512}, {"a0", e->a0}, {"a1", e->a1},
      {"gmax_stn", e->gmax_stn}, {"gmax_fstn", e->gmax_fstn}, {"gmax_g", e->gmax_g}, {"gmax_pf", e->gmax_pf},
      {"t3", e->t3}, {"t64", e->t64}, {"cset", e->cset},
      {"t64s_hi", e->t64s.hi}, {"t64s_lo", e->t64s.lo}, {"ts0", e->ts0},
      {"stats0", e->stats0}, {"stats1", e->stats1}, {"gn0", e->gn0}, {"gn1", e->gn1}, {"rot_partial", e->rot_partial}};
  m["x64_hi"] = e->x64.hi; m["x64_lo"] = e->x64.lo; m["f64_hi"] = e->f64.hi; m["f64_lo"] = e->f64.lo;
  m["a128_hi"] = e->a128.hi; m["a128_lo"] = e->a128.lo; m["pf_hi"] = e->pf16.hi; m["pf_lo"] = e->pf16.lo;
  m["a512_hi"] = e->a512.hi; m["a512_lo"] = e->a512.lo;
  auto it = m.find(name);
  if (it == m.end() || it->second == nullptr) return fail(e, CATRE_ERR_INVALID_ARG, "no debug buffer '%s'", name);
  CU_TRY(e, cudaDeviceSynchronize());
  CU_TRY(e, cudaMemcpy(dst_host, it->second, bytes, cudaMemcpyDeviceToHost));
  return CATRE_OK;
}

int catre_profile_enable(catre_engine* e, int32_t on) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  e->prof_on = on != 0;
  return CATRE_OK;
}

static int profile_drain(catre_engine* e) {
  for (auto& ev : e->ev_pending) {
    CU_TRY(e, cudaEventSynchronize(ev.b));
    float ms = 0.f;
    CU_TRY(e, cudaEventElapsedTime(&ms, ev.a, ev.b));
    e->prof_ms[ev.grp] += ms;
    e->prof_n[ev.grp] += 1;
    e->ev_pool.push_back(ev.a);
    e->ev_pool.push_back(ev.b);
  }
  e->ev_pending.clear();
  return 0;
}

int catre_profile_reset(catre_engine* e) {
  if (!e) return CATRE_ERR_INVALID_ARG;
  int rc = profile_drain(e);
  for (int i = 0; i < G_NUM; ++i) { e->prof_ms[i] = 0; e->prof_n[i] = 0; }
  return rc;
}

int catre_profile_get(catre_engine* e, int32_t which, double* total_ms, int64_t* launches) {
  if (!e || which < 0 || which >= G_NUM || !total_ms || !launches) return CATRE_ERR_INVALID_ARG;
  int rc = profile_drain(e);
  if (rc) return rc;
  *total_ms = e->prof_ms[which];
  *launches = e->prof_n[which];
  return CATRE_OK;
}

}  // extern "C"

// Fused Ranger step (RAdam + Lookahead + gradient centralisation) over all parameter tensors of the model in two
// launches -- the optimiser the shipped config trains with (configs/.../aug05_..._120e.py:49; reference implementation:
// lib/torch_utils/solver/ranger.py:102-200, where one step costs ~10 small launches per tensor, ~700 for the model) and
// the NaN guard the training loop applies to the gradients before it (core/catre/engine/engine.py:349-352).
// Thread-independent functor kernels like train_kernels.cuh, so tests/emu runs the same source on the CPU.
#pragma once
#include "train_kernels.cuh"

namespace catre_train {

// one row of the tensor table (int64 [n, 8]): device addresses of the parameter, its gradient and the optimiser state
// (exp_avg, exp_avg_sq, slow_buffer), the element count and, for tensors with more than gc_min_dims dimensions, the
// length of one output row (numel / shape[0]) over which the gradient is centralised (0 = no centralisation).
struct RangerTensor {
  long long p, g, m, v, slow, numel, row_len, reserved;
};
struct RangerArgs {
  float beta1, beta2, eps;
  float omb1, omb2;  // 1 - beta1, 1 - beta2 evaluated in double on the host and rounded once, as torch does with its Python scalars
  float step_size;   // RAdam step size of this step (host: ranger.py:160-178)
  int rectified;     // N_sma > threshold: adaptive update, else plain momentum (ranger.py:184-188)
  float alpha;       // lookahead interpolation
  int lookahead;     // step % k == 0 (ranger.py:194-200)
  int nan_to_num;    // nan -> 0, +inf -> 1e5, -inf -> -1e5 on the gradient first (engine.py:351)
};

TK_HD int ranger_find(const long long* start, int n, long long x) {  // largest t with start[t] <= x
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (start[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}
TK_HD float ranger_fix(float g, int on) {
  if (!on) return g;
  if (g != g) return 0.0f;
  if (g > 3.4028234e38f) return 1e5f;
  if (g < -3.4028234e38f) return -1e5f;
  return g;
}

// mean of every centralised gradient row (ranger.py:147-148).  grid (ceil(total_rows / nt))
struct KRangerRowMean {
  const RangerTensor* T; const long long* row_start; int n; long long total_rows; float* rowmean; int nan_to_num;
  TK_HD void operator()(const Idx& i) const {
    const long long r = (long long)i.bx * i.nt + i.tx;
    if (r >= total_rows) return;
    const int t = ranger_find(row_start, n, r);
    const RangerTensor& e = T[t];
    const float* g = reinterpret_cast<const float*>(e.g) + (r - row_start[t]) * e.row_len;
    double s = 0.0;
    for (long long j = 0; j < e.row_len; ++j) s += ranger_fix(g[j], nan_to_num);
    rowmean[r] = (float)(s / (double)e.row_len);
  }
};

// the update of every element of every tensor (ranger.py:150-200).  lr_wd [n, 2] = (group lr, group weight decay).
// grid (ceil(total_elems / nt))
struct KRangerUpdate {
  const RangerTensor* T; const long long *elem_start, *row_start; const float *lr_wd, *rowmean; int n; long long total;
  RangerArgs a;
  TK_HD void operator()(const Idx& i) const {
    const long long x = (long long)i.bx * i.nt + i.tx;
    if (x >= total) return;
    const int t = ranger_find(elem_start, n, x);
    const RangerTensor& e = T[t];
    const long long k = x - elem_start[t];
    float g = ranger_fix(reinterpret_cast<const float*>(e.g)[k], a.nan_to_num);
    if (e.row_len > 0) g -= rowmean[row_start[t] + k / e.row_len];
    float* pm = reinterpret_cast<float*>(e.m) + k;
    float* pv = reinterpret_cast<float*>(e.v) + k;
    float* pp = reinterpret_cast<float*>(e.p) + k;
    const float v = *pv * a.beta2 + a.omb2 * g * g;
    const float m = *pm * a.beta1 + a.omb1 * g;
    *pv = v; *pm = m;
    const float lr = lr_wd[2 * t], wd = lr_wd[2 * t + 1];
    float p = *pp;
    if (wd != 0.0f) p += p * (-wd * lr);
    if (a.rectified) p += (-a.step_size * lr) * (m / (sqrtf(v) + a.eps));
    else p += (-a.step_size * lr) * m;
    if (a.lookahead) {
      float* ps = reinterpret_cast<float*>(e.slow) + k;
      const float s = *ps + a.alpha * (p - *ps);
      *ps = s;
      p = s;
    }
    *pp = p;
  }
};

}  // namespace catre_train

// Training step of the CATRE refiner: kernels of the forward-with-saved-activations, the
// loss and the backward chain.  All fp32 FMA on CUDA cores, row-major point-major [rows, channels] buffers; the
// stage-by-stage algebra is the one of oracle/train_oracle.py::manual_forward / manual_backward (verified there
// against autograd).
//
// Every kernel except the shared-memory GEMM is a functor whose operator() is the work of ONE thread with no
// inter-thread communication (Idx carries the block / thread coordinates).  On the GPU the generic
// `tk_run<K><<<grid, block>>>(k)` wrapper calls it with blockIdx / threadIdx; with CATRE_HOST_EMU defined the same
// source compiles with a plain C++ compiler and tests/emu runs the grid as nested loops, so indexing and the host
// orchestration (train_chain.cuh) are checked on the CPU against the oracle.  The emulation build is test
// infrastructure; the product library never contains it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef CATRE_HOST_EMU
#define TK_HD inline
#else
#define TK_HD __host__ __device__ __forceinline__
#endif

namespace catre_train {

struct Idx {
  int bx, by, bz, tx;  // block coordinates, thread in block
  int nt;              // threads per block
};

TK_HD float tk_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
TK_HD float tk_gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * expf(-0.5f * x * x) * 0.39894228040143268f;
}
TK_HD float tk_i2f(int v) { float f; memcpy(&f, &v, sizeof(f)); return f; }
TK_HD int tk_f2i(float f) { int v; memcpy(&v, &f, sizeof(v)); return v; }
// In-order sum of n values v[0], v[stride], ...: the loads of 32 terms are issued before their additions, so a thread waits
// for one memory round trip per 32 terms instead of one per term; the additions stay in index order (same bits as a plain loop).
template <class T>
TK_HD T tk_sum_inorder(const T* v, long long n, long long stride) {
  T acc = (T)0;
  long long k = 0;
  for (; k + 32 <= n; k += 32) {
    T t[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) t[u] = v[(k + u) * stride];
#pragma unroll
    for (int u = 0; u < 32; ++u) acc += t[u];
  }
  for (; k < n; ++k) acc += v[k * stride];
  return acc;
}
// Sum of n doubles v[0], v[stride], ... in four interleaved partial sums (term k goes to sum k mod 4), combined as
// (s0 + s1) + (s2 + s3): a fixed order, so still deterministic, with four independent addition chains instead of one --
// the merges of 256 chunk partials were bound by the latency of their dependent fp64 additions.
TK_HD double tk_sum4(const double* v, long long n, long long stride) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  long long k = 0;
  for (; k + 16 <= n; k += 16) {
    double t[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) t[u] = v[(k + u) * stride];
#pragma unroll
    for (int u = 0; u < 16; u += 4) { a0 += t[u]; a1 += t[u + 1]; a2 += t[u + 2]; a3 += t[u + 3]; }
  }
  for (; k < n; ++k) {  // the tail keeps the k mod 4 assignment (k is a multiple of 16 here)
    const double x = v[k * stride];
    switch ((int)(k & 3)) { case 0: a0 += x; break; case 1: a1 += x; break; case 2: a2 += x; break; default: a3 += x; }
  }
  return (a0 + a1) + (a2 + a3);
}
struct alignas(16) TkI4 { int x, y, z, w; };
struct alignas(16) TkF4 { float x, y, z, w; };
TK_HD float tk_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// ---- re-posed points (core/catre/engine/batching.py:127-140 / batch_test.py:78-97): set 2b = pcl_b - t_b,
//      set 2b+1 = R_b (s_b * kps_b).  grid (ceil(N / nt), B)
struct KUpdatePoints {
  const float *pcl, *kps, *pose, *scale;
  float* q;
  int N;
  TK_HD void operator()(const Idx& i) const {
    const int n = i.bx * i.nt + i.tx, b = i.by;
    if (n >= N) return;
    const float* P = pose + (size_t)b * 12;
    const float* s = scale + (size_t)b * 3;
    const float* x = pcl + ((size_t)b * N + n) * 3;
    const float* k = kps + ((size_t)b * N + n) * 3;
    float* qo = q + ((size_t)(2 * b) * N + n) * 3;
    float* qk = q + ((size_t)(2 * b + 1) * N + n) * 3;
    const float k0 = k[0] * s[0], k1 = k[1] * s[1], k2 = k[2] * s[2];
    for (int r = 0; r < 3; ++r) {
      qo[r] = x[r] - P[r * 4 + 3];
      qk[r] = P[r * 4 + 0] * k0 + P[r * 4 + 1] * k1 + P[r * 4 + 2] * k2;
    }
  }
};

// the same two sets when the caller has already re-posed the points (the reference's forward takes x = pcl - t and
// tfd_kps = R (s * kps), CATRE_disR_shared.py:40-56): q[2b] = x_pm[b], q[2b+1] = tfd_pm[b].  grid (ceil(3 N / nt), B)
struct KInterleave {
  const float *x_pm, *tfd_pm; float* q; int N;
  TK_HD void operator()(const Idx& i) const {
    const int e = i.bx * i.nt + i.tx, b = i.by;
    if (e >= 3 * N) return;
    q[(size_t)(2 * b) * N * 3 + e] = x_pm[(size_t)b * N * 3 + e];
    q[(size_t)(2 * b + 1) * N * 3 + e] = tfd_pm[(size_t)b * N * 3 + e];
  }
};

// ---- generic strided, batched GEMM: C(m,n,z) = act(sum_k A(m,k,z) B(k,n,z) + bias(n,z)) [+ C].  One thread per
//      output element.  grid (ceil(M / (nt / 64)), ceil(N / 64), batch * splits); nt a multiple of 64.
//      With splits > 1 (batch must be 1) block z sums k in [z * k_per, (z+1) * k_per) into partial[z][m][n].
struct GemmP {
  const float* A; long long sam, sak, sab;
  const float* B; long long sbk, sbn, sbb;
  float* C; long long scm, scn, scb;
  const float* bias; long long sbias_b;  // bias[n + z * sbias_b] or nullptr
  int M, N, K;
  int relu, accumulate;
  int splits; int k_per; float* partial;
  int f16;  // tensor-core GEMM only (train_gemm_tc.cuh): 1 = fp16 hi/lo operands (forward), 0 = bf16 hi/lo (backward)
  // tensor-core GEMM only, weight-gradient launches (A(m, k) = dy[k, m], m fast, splits > 1): when set, the CTAs of the first
  // column tile also sum their A rows over k -- the bias gradient's share of the split -- into bias_grad_partial[split][M]
  float* bias_grad_partial;
};
struct KGemmNaive {
  GemmP p;
  TK_HD void operator()(const Idx& i) const {
    const int n = i.by * 64 + (i.tx & 63);
    const int m = i.bx * (i.nt >> 6) + (i.tx >> 6);
    if (m >= p.M || n >= p.N) return;
    int z = i.bz, k0 = 0, k1 = p.K;
    if (p.splits > 1) { k0 = z * p.k_per; k1 = k0 + p.k_per < p.K ? k0 + p.k_per : p.K; z = 0; }
    const float* a = p.A + (long long)z * p.sab + (long long)m * p.sam;
    const float* b = p.B + (long long)z * p.sbb + (long long)n * p.sbn;
    float acc = 0.0f;
    for (int k = k0; k < k1; ++k) acc = fmaf(a[(long long)k * p.sak], b[(long long)k * p.sbk], acc);
    if (p.splits > 1) { p.partial[((size_t)i.bz * p.M + m) * p.N + n] = acc; return; }
    if (p.bias) acc += p.bias[n + (long long)z * p.sbias_b];
    if (p.relu) acc = acc > 0.0f ? acc : 0.0f;
    float* c = p.C + (long long)z * p.scb + (long long)m * p.scm + (long long)n * p.scn;
    *c = p.accumulate ? *c + acc : acc;
  }
};
// fixed-order sum of the split-K partials (+ the GEMM's bias / ReLU when the split was the launcher's choice).
// grid (ceil(M * N / nt))
struct KSplitReduce {
  const float* partial; float* C; long long scm, scn; int M, N, splits, accumulate;
  const float* bias = nullptr; int relu = 0;
  const float* bg_partial = nullptr; float* bg_out = nullptr;  // bias-gradient shares [splits][M] -> bg_out[M] += their sum
  TK_HD void operator()(const Idx& i) const {
    const long long e = (long long)i.bx * i.nt + i.tx;
    if (bg_partial && e < M) {
      float a = 0.0f;
      for (int z = 0; z < splits; ++z) a += bg_partial[(size_t)z * M + e];
      bg_out[e] += a;
    }
    if (e >= (long long)M * N) return;
    const size_t mn = (size_t)M * N;
    float acc = 0.0f;
    int z = 0;
    for (; z + 4 <= splits; z += 4) {  // four loads in flight; the additions stay in slice order
      const float a0 = partial[(size_t)z * mn + e], a1 = partial[(size_t)(z + 1) * mn + e];
      const float a2 = partial[(size_t)(z + 2) * mn + e], a3 = partial[(size_t)(z + 3) * mn + e];
      acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; z < splits; ++z) acc += partial[(size_t)z * mn + e];
    if (bias) acc += bias[e % N];
    if (relu) acc = acc > 0.0f ? acc : 0.0f;
    float* c = C + (e / N) * scm + (e % N) * scn;
    *c = accumulate ? *c + acc : acc;
  }
};

// ---- column max + arg-max over the N points of each set (first index wins ties, like torch.max on the CPU).
//      z [S, N, C] -> vmax [S, C], arg [S, C].  grid (ceil(C / nt), S)
struct KColMaxArg {
  const float* z; float* vmax; int* arg; int N, C;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, s = i.by;
    if (c >= C) return;
    const float* p = z + (size_t)s * N * C + c;
    float best = p[0]; int bi = 0;
    for (int n = 1; n < N; ++n) { const float v = p[(size_t)n * C]; if (v > best) { best = v; bi = n; } }
    vmax[(size_t)s * C + c] = best; arg[(size_t)s * C + c] = bi;
  }
};

// The same in two stages for more parallelism: stage 1 scans `per` consecutive points per thread, stage 2 merges the chunks in
// order with a strict comparison, so the first index still wins ties and the result equals KColMaxArg's.
// stage 1: grid (ceil(C / nt), chunks, S) -> pv / pi [S, chunks, C];  stage 2: grid (ceil(C / nt), S)
struct KColMaxArgPart {
  const float* z; float* pv; int* pi; int N, C, per;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, ch = i.by, s = i.bz;
    if (c >= C) return;
    const int n0 = ch * per, n1 = n0 + per < N ? n0 + per : N;
    const float* p = z + (size_t)s * N * C + c;
    float best = p[(size_t)n0 * C]; int bi = n0;
    for (int n = n0 + 1; n < n1; ++n) { const float v = p[(size_t)n * C]; if (v > best) { best = v; bi = n; } }
    const size_t o = ((size_t)s * ((N + per - 1) / per) + ch) * C + c;
    pv[o] = best; pi[o] = bi;
  }
};
struct KColMaxArgMerge {
  const float* pv; const int* pi; float* vmax; int* arg; int chunks, C;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, s = i.by;
    if (c >= C) return;
    const size_t o = (size_t)s * chunks * C + c;
    float best = pv[o]; int bi = pi[o];
    for (int ch = 1; ch < chunks; ++ch) { const float v = pv[o + (size_t)ch * C]; if (v > best) { best = v; bi = pi[o + (size_t)ch * C]; } }
    vmax[(size_t)s * C + c] = best; arg[(size_t)s * C + c] = bi;
  }
};

// t[s, :] += I_k   (pointnets/pointnet.py:37-40, 72-77).  grid (ceil(k / nt), S)
struct KAddEye {
  float* t; int k;
  TK_HD void operator()(const Idx& i) const {
    const int d = i.bx * i.nt + i.tx;
    if (d < k) t[(size_t)i.by * k * k + (size_t)d * k + d] += 1.0f;
  }
};

// dst[e] = scale * src[e]  (gradients handed to the caller with the upstream factor applied).  grid (ceil(n / nt))
struct KScaleCopy {
  const float* src; float* dst; float scale; long long n;
  TK_HD void operator()(const Idx& i) const {
    const long long e = (long long)i.bx * i.nt + i.tx;
    if (e < n) dst[e] = scale * src[e];
  }
};

// dst[t][e] = src[t][e] for up to kMax tensors in one launch (the optimiser's parameters -> the engine's fp32 copies); a null
// src skips the tensor.  grid (chunks, tensors): block (bx, by) copies elements bx * per .. of tensor by.
struct KMultiCopy {
  static constexpr int kMax = 80;
  const float* src[kMax]; float* dst[kMax]; int n[kMax]; int per;
  TK_HD void operator()(const Idx& i) const {
    const float* s = src[i.by];
    if (!s) return;
    float* d = dst[i.by];
    const int lo = i.bx * per, hi = lo + per < n[i.by] ? lo + per : n[i.by];
    for (int e = lo + i.tx; e < hi; e += i.nt) d[e] = s[e];
  }
};

// d[e] = act[e] > 0 ? d[e] : 0   (ReLU backward).  grid (ceil(n / nt))
struct KReluMask {
  float* d; const float* act; long long n;
  TK_HD void operator()(const Idx& i) const {
    const long long e = (long long)i.bx * i.nt + i.tx;
    if (e < n && !(act[e] > 0.0f)) d[e] = 0.0f;
  }
};

// out[chunk, c] = sum over the rows of the chunk of d[r, c]  (bias gradients; two stages keep it deterministic).
// grid (ceil(C / nt), chunks); rows_per = rows per chunk.  With accumulate the result is added to out.
struct KColSum {
  const float* d; float* out; long long rows; int C; long long rows_per; int accumulate; long long ld;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx;
    if (c >= C) return;
    const long long r0 = (long long)i.by * rows_per;
    const long long r1 = r0 + rows_per < rows ? r0 + rows_per : rows;
    const float acc = r1 > r0 ? tk_sum_inorder(d + r0 * ld + c, r1 - r0, ld) : 0.0f;
    float* o = out + (size_t)i.by * C + c;
    *o = accumulate ? *o + acc : acc;
  }
};

// ---- sparse backward of `max over points of [relu](x W^T + b)` (oracle: max_layer_bwd).  Only the arg-max row of
//      each (set, channel) receives a gradient.
// dx[(s, n), k] = sum over the channels c (ascending) whose arg-max point is n of d[s,c] W[c,k].  A thread owns column
// k of MAXBWD_PTS consecutive points of one set and scans the set's C arg-max entries (warp-uniform reads), so every
// dx element is written exactly once (no zero fill, no read-modify-write chain) in a fixed order.
// grid (ceil(K / nt), ceil(N / MAXBWD_PTS), S)
constexpr int MAXBWD_PTS = 8;
struct KMaxBwdDx {
  const float *dmax, *relu_max, *W; const int* arg; float* dx; int N, C, K;
  TK_HD void operator()(const Idx& i) const {
    const int k = i.bx * i.nt + i.tx, n0 = i.by * MAXBWD_PTS, s = i.bz;
    if (k >= K) return;
    float acc[MAXBWD_PTS];
    for (int j = 0; j < MAXBWD_PTS; ++j) acc[j] = 0.0f;
    for (int c = 0; c < C; ++c) {
      const int j = arg[(size_t)s * C + c] - n0;
      if (j < 0 || j >= MAXBWD_PTS) continue;
      float d = dmax[(size_t)s * C + c];
      if (relu_max && !(relu_max[(size_t)s * C + c] > 0.0f)) d = 0.0f;
      const float v = d * W[(size_t)c * K + k];
      for (int t = 0; t < MAXBWD_PTS; ++t)
        if (t == j) acc[t] += v;
    }
    for (int j = 0; j < MAXBWD_PTS && n0 + j < N; ++j) dx[((size_t)s * N + n0 + j) * K + k] = acc[j];
  }
};
// The same result (same summation order: channels ascending) without the C-long scan per output element: first the
// inverse of the arg-max map, then a gather.
// Inverse map = the set's channels sorted by (arg-max point, channel): list[s, :] with key[s, :] = the arg-max point of each
// entry, and per point n the range start[s, n] .. + cnt[s, n] of its channels.  Two kernels without data-dependent stores:
// KMaxBwdRank, one thread per (set, channel): position of channel c = number of channels c' with (arg[c'], c') < (arg[c], c),
//   one scan of the set's C entries as 128-bit words (C a multiple of 4, rows 16-byte aligned; scalar otherwise).
//   grid (ceil(C / nt), S)
// KMaxBwdRange, one thread per (set, point): lower / upper bound of n in the sorted keys.  grid (ceil(N / nt), S)
struct KMaxBwdRank {
  const int* arg; int *list, *key; int C;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, s = i.by;
    if (c >= C) return;
    const int* a = arg + (size_t)s * C;
    const int n = a[c];
    int pos = 0;
    if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
      const TkI4* a4 = reinterpret_cast<const TkI4*>(a);
      const int q = c >> 2;
      for (int j = 0; j < q; ++j) {  // words entirely below c: ties count
        const TkI4 v = a4[j];
        pos += (v.x <= n) + (v.y <= n) + (v.z <= n) + (v.w <= n);
      }
      {
        const TkI4 v = a4[q];
        const int r = c & 3;
        pos += (r > 0 ? v.x <= n : false) + (r > 1 ? v.y <= n : (r < 1 ? v.y < n : false)) +
               (r > 2 ? v.z <= n : (r < 2 ? v.z < n : false)) + (r < 3 ? v.w < n : false);
      }
      for (int j = q + 1; j < C / 4; ++j) {  // words entirely above c: strict
        const TkI4 v = a4[j];
        pos += (v.x < n) + (v.y < n) + (v.z < n) + (v.w < n);
      }
    } else {
      for (int j = 0; j < c; ++j) pos += a[j] <= n;
      for (int j = c + 1; j < C; ++j) pos += a[j] < n;
    }
    list[(size_t)s * C + pos] = c;
    key[(size_t)s * C + pos] = n;
  }
};
struct KMaxBwdRange {
  const int* key; int *start, *cnt; int N, C;
  TK_HD void operator()(const Idx& i) const {
    const int n = i.bx * i.nt + i.tx, s = i.by;
    if (n >= N) return;
    const int* k = key + (size_t)s * C;
    int lo = 0, hi = C;  // first position with key >= n
    while (lo < hi) { const int m = (lo + hi) >> 1; if (k[m] < n) lo = m + 1; else hi = m; }
    const int first = lo;
    hi = C;              // first position with key > n
    while (lo < hi) { const int m = (lo + hi) >> 1; if (k[m] <= n) lo = m + 1; else hi = m; }
    start[(size_t)s * N + n] = first; cnt[(size_t)s * N + n] = lo - first;
  }
};
// dx[(s, n), k .. k+3] = sum over the point's channels (ascending) of d[s,c] W[c, k .. k+3], zeroed where the layer's input
// activation is not positive when `act` is given (the ReLU backward of the layer below, folded in: only the few rows that
// carry a gradient read `act`); every element written once, 16 bytes per thread.  K a multiple of 4, W / dx / act 16-byte
// aligned.  A block covers nt / (K / 4) consecutive points (ppb >= 1) so that it has nt threads for any K; the channels of a
// point are fetched eight at a time (index, gradient and weight row of all eight in flight, then added in order): a point
// that holds hundreds of a set's maxima is one long serial chain per thread, and its length in round trips is what the
// kernel takes.  grid (ceil(K / 4 / nt), ceil(N / ppb), S)
struct KMaxBwdGather {
  const float *dmax, *relu_max, *W; const int *start, *cnt, *list; float* dx; const float* act; int N, C, K, ppb;
  TK_HD void operator()(const Idx& i) const {
    const int kq = K >> 2;
    int k, n;
    if (ppb > 1) { k = 4 * (i.tx % kq); n = i.by * ppb + i.tx / kq; if (i.tx >= ppb * kq) return; }
    else { k = 4 * (i.bx * i.nt + i.tx); n = i.by; }
    const int s = i.bz;
    if (k >= K || n >= N) return;
    const int m = cnt[(size_t)s * N + n];
    const int* l = list + (size_t)s * C + start[(size_t)s * N + n];
    const float* dm = dmax + (size_t)s * C;
    const float* rm = relu_max ? relu_max + (size_t)s * C : nullptr;
    TkF4 acc = {0.0f, 0.0f, 0.0f, 0.0f};
    int j = 0;
    int cn[8];  // the next batch's channel indices are fetched while the current batch's gradients and weight rows are in flight
    if (m >= 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) cn[u] = l[u];
    }
    for (; j + 8 <= m; j += 8) {
      int c[8]; float d[8]; TkF4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) c[u] = cn[u];
      if (j + 16 <= m) {
#pragma unroll
        for (int u = 0; u < 8; ++u) cn[u] = l[j + 8 + u];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        d[u] = dm[c[u]];
        if (rm && !(rm[c[u]] > 0.0f)) d[u] = 0.0f;
        w[u] = *reinterpret_cast<const TkF4*>(W + (size_t)c[u] * K + k);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += d[u] * w[u].x; acc.y += d[u] * w[u].y; acc.z += d[u] * w[u].z; acc.w += d[u] * w[u].w; }
    }
    for (; j < m; ++j) {
      const int c = l[j];
      float d = dm[c];
      if (rm && !(rm[c] > 0.0f)) d = 0.0f;
      const TkF4 w = *reinterpret_cast<const TkF4*>(W + (size_t)c * K + k);
      acc.x += d * w.x; acc.y += d * w.y; acc.z += d * w.z; acc.w += d * w.w;
    }
    const size_t o = ((size_t)s * N + n) * K + k;
    if (act && m > 0) {
      const TkF4 a = *reinterpret_cast<const TkF4*>(act + o);
      if (!(a.x > 0.0f)) acc.x = 0.0f;
      if (!(a.y > 0.0f)) acc.y = 0.0f;
      if (!(a.z > 0.0f)) acc.z = 0.0f;
      if (!(a.w > 0.0f)) acc.w = 0.0f;
    }
    *reinterpret_cast<TkF4*>(dx + o) = acc;
  }
};
// dW[c, k] += sum_s d[s,c] x[(s, arg[s,c]), k];  db[c] += sum_s d[s,c].  grid (ceil(K / nt), C)
struct KMaxBwdDw {
  const float *dmax, *relu_max, *x; const int* arg; float *dW, *db; int S, N, C, K;
  TK_HD void operator()(const Idx& i) const {
    const int k = i.bx * i.nt + i.tx, c = i.by;
    if (k >= K) return;
    float acc = 0.0f, accb = 0.0f;
    for (int s = 0; s < S; ++s) {
      float d = dmax[(size_t)s * C + c];
      if (relu_max && !(relu_max[(size_t)s * C + c] > 0.0f)) d = 0.0f;
      acc = fmaf(d, x[((size_t)s * N + arg[(size_t)s * C + c]) * K + k], acc);
      accb += d;
    }
    dW[(size_t)c * K + k] += acc;
    if (k == 0) db[c] += accb;
  }
};
// dpf[(s, arg[s,c]), c] += dpfmax[s, c]  (max_n pointfeat of the ts-head input).  grid (ceil(C / nt), S)
struct KScatterMax {
  const float* dmax; const int* arg; float* dx; int N, C;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, s = i.by;
    if (c < C) dx[((size_t)s * N + arg[(size_t)s * C + c]) * C + c] += dmax[(size_t)s * C + c];
  }
};

// ---- GroupNorm (32 groups of 8 channels over the P points of an object; P = 1 for the ts head), eps 1e-5,
//      biased variance, fp64 accumulation in two deterministic stages: the points of an object are cut into `chunks`
//      pieces of `per` points, thread (b, chunk, g) sums its piece into part [B, chunks, 32, 2] (doubles), then thread
//      (b, g) adds the pieces in order.  y [B, P, 256] -> st [B, 32, 2] = (mean, rstd).
//      stage 1: grid (B, chunks), nt = 32;  stage 2: grid (B), nt = 32
struct KGnStatsPart {
  const float* y; double* part; int P, chunks, per;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx, ch = i.by, g = i.tx;
    if (g >= 32) return;
    const int n0 = ch * per, n1 = n0 + per < P ? n0 + per : P;
    const float* p = y + (size_t)b * P * 256 + g * 8;
    double s = 0.0, ss = 0.0;
    for (int n = n0; n < n1; ++n)
      for (int j = 0; j < 8; ++j) { const double v = p[(size_t)n * 256 + j]; s += v; ss += v * v; }
    double* o = part + (((size_t)b * chunks + ch) * 32 + g) * 2;
    o[0] = s; o[1] = ss;
  }
};
struct KGnStats {
  const double* part; float* st; int P, chunks;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx, g = i.tx;
    if (g >= 32) return;
    const double* o = part + ((size_t)b * chunks * 32 + g) * 2;
    const double s = tk_sum4(o, chunks, 64), ss = tk_sum4(o + 1, chunks, 64);
    const double cnt = 8.0 * P, mu = s / cnt;
    double var = ss / cnt - mu * mu;
    if (var < 0.0) var = 0.0;
    st[((size_t)b * 32 + g) * 2 + 0] = (float)mu;
    st[((size_t)b * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
  }
};
// u = gelu(gamma (y - mu) rstd + beta).  grid (ceil(B * P * 256 / nt))
// Four consecutive channels per thread (one 16-byte load and store; they share the object and the group): n a multiple of 4,
// y / u 16-byte aligned.  grid (ceil(n / 4 / nt))
struct KGnGeluFwd {
  const float *y, *st, *gamma, *beta; float* u; int P; long long n;
  TK_HD void operator()(const Idx& i) const {
    const long long e = 4 * ((long long)i.bx * i.nt + i.tx);
    if (e >= n) return;
    const int c = (int)(e & 255);
    const long long b = e / ((long long)P * 256);
    const float* s = st + ((size_t)b * 32 + (c >> 3)) * 2;
    const float mu = s[0], rstd = s[1];
    const TkF4 v = *reinterpret_cast<const TkF4*>(y + e);
    TkF4 r;
    r.x = tk_gelu(gamma[c] * ((v.x - mu) * rstd) + beta[c]);
    r.y = tk_gelu(gamma[c + 1] * ((v.y - mu) * rstd) + beta[c + 1]);
    r.z = tk_gelu(gamma[c + 2] * ((v.z - mu) * rstd) + beta[c + 2]);
    r.w = tk_gelu(gamma[c + 3] * ((v.w - mu) * rstd) + beta[c + 3]);
    *reinterpret_cast<TkF4*>(u + e) = r;
  }
};
// backward, pass 1 (oracle: _gn_gelu_backward): with xhat = (y - mu) rstd, dn = du gelu'(gamma xhat + beta),
// g = dn gamma:  m [B, 32, 2] = group means of (g, g xhat);  dgam [B, 256] = sum_p dn xhat, dbet [B, 256] = sum_p dn.
// Two deterministic stages like the statistics: part [B, chunks, 32, 18] doubles = (sum g, sum g xhat, 8 x dgam, 8 x dbet).
// stage 1: grid (B, chunks), nt = 32;  stage 2: grid (B), nt = 32
struct KGnBwdPart {
  const float *du, *y, *st, *gamma, *beta; double* part; int P, chunks, per;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx, ch = i.by, g = i.tx;
    if (g >= 32) return;
    const int n0 = ch * per, n1 = n0 + per < P ? n0 + per : P;
    const float mu = st[((size_t)b * 32 + g) * 2], rstd = st[((size_t)b * 32 + g) * 2 + 1];
    double s1 = 0.0, s2 = 0.0, dga[8], dbe[8];
    for (int j = 0; j < 8; ++j) dga[j] = dbe[j] = 0.0;
    const size_t base = (size_t)b * P * 256 + g * 8;
    for (int n = n0; n < n1; ++n)
      for (int j = 0; j < 8; ++j) {
        const size_t e = base + (size_t)n * 256 + j;
        const float xh = (y[e] - mu) * rstd, ga = gamma[g * 8 + j];
        const float dn = du[e] * tk_gelu_grad(ga * xh + beta[g * 8 + j]);
        s1 += (double)(dn * ga); s2 += (double)(dn * ga) * xh;
        dga[j] += (double)dn * xh; dbe[j] += dn;
      }
    double* o = part + (((size_t)b * chunks + ch) * 32 + g) * 18;
    o[0] = s1; o[1] = s2;
    for (int j = 0; j < 8; ++j) { o[2 + j] = dga[j]; o[10 + j] = dbe[j]; }
  }
};
// stage 2: one thread per (object, group, entry of the 18): grid (B), nt = 32 * 18 (consecutive threads read consecutive doubles)
struct KGnBwdSums {
  const double* part; float *m, *dgam, *dbet; int P, chunks;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx, g = i.tx / 18, j = i.tx % 18;
    if (g >= 32) return;
    const double a = tk_sum4(part + (size_t)b * chunks * 576 + g * 18 + j, chunks, 576);
    if (j < 2) m[((size_t)b * 32 + g) * 2 + j] = (float)(a / (8.0 * P));
    else if (j < 10) dgam[(size_t)b * 256 + g * 8 + (j - 2)] = (float)a;
    else dbet[(size_t)b * 256 + g * 8 + (j - 10)] = (float)a;
  }
};
// pass 2: dy = rstd (g - m1 - xhat m2), written over du.  grid (ceil(B * P * 256 / nt))
struct KGnBwdApply {
  float* du; const float *y, *st, *gamma, *beta, *m; int P; long long n;
  TK_HD float one(float d, float yv, float mu, float rstd, float m1, float m2, int c) const {
    const float xh = (yv - mu) * rstd;
    const float g = d * tk_gelu_grad(gamma[c] * xh + beta[c]) * gamma[c];
    return rstd * (g - m1 - xh * m2);
  }
  TK_HD void operator()(const Idx& i) const {  // four consecutive channels per thread, like KGnGeluFwd
    const long long e = 4 * ((long long)i.bx * i.nt + i.tx);
    if (e >= n) return;
    const int c = (int)(e & 255);
    const long long b = e / ((long long)P * 256);
    const size_t gi = ((size_t)b * 32 + (c >> 3)) * 2;
    const float mu = st[gi], rstd = st[gi + 1], m1 = m[gi], m2 = m[gi + 1];
    const TkF4 yv = *reinterpret_cast<const TkF4*>(y + e);
    TkF4 d = *reinterpret_cast<const TkF4*>(du + e);
    d.x = one(d.x, yv.x, mu, rstd, m1, m2, c);
    d.y = one(d.y, yv.y, mu, rstd, m1, m2, c + 1);
    d.z = one(d.z, yv.z, mu, rstd, m1, m2, c + 2);
    d.w = one(d.w, yv.w, mu, rstd, m1, m2, c + 3);
    *reinterpret_cast<TkF4*>(du + e) = d;
  }
};

// ---- rotation-head tail (conv_out_per_rot_head.py:136-139 by linearity): wsum[b, c] = sum_p wp[p] u1[b, p, c].
//      grid (B), nt = 256
struct KRotWsum {
  const float *u1, *wp; float* wsum; int P;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx, c = i.tx;
    if (c >= 256) return;
    const float* u = u1 + (size_t)b * P * 256 + c;
    float acc = 0.0f;
    for (int p = 0; p < P; ++p) acc = fmaf(wp[p], u[(size_t)p * 256], acc);
    wsum[(size_t)b * 256 + c] = acc;
  }
};
// Two stages for more parallelism: part[b, ch, c] = sum over the chunk's points (grid (1, chunks, B), nt = 256), then KColSum
// over the chunks of an object (rows_per = chunks).
struct KRotWsumPart {
  const float *u1, *wp; float* part; int P, per;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.tx, ch = i.by, b = i.bz;
    if (c >= 256) return;
    const int p0 = ch * per, p1 = p0 + per < P ? p0 + per : P;
    const float* u = u1 + (size_t)b * P * 256 + c;
    float acc = 0.0f;
    for (int p = p0; p < p1; ++p) acc = fmaf(wp[p], u[(size_t)p * 256], acc);
    part[((size_t)b * ((P + per - 1) / per) + ch) * 256 + c] = acc;
  }
};
// r6[b, 3h + j] = Wn[j, :] . wsum[b, :] + bn[j] sum_p wp[p] + bp; swp[h] = sum_p wp[p], computed once per step by a chunked
// column sum (a serial sum over the P weights in every thread was 30-50 us of pure load latency).  grid (B), nt >= 3
struct KRotOut {
  const float *wsum, *wn, *bn, *swp, *bp; float* r6; int P, h;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx, j = i.tx;
    if (j >= 3) return;
    float acc = 0.0f;
    for (int c = 0; c < 256; ++c) acc = fmaf(wn[j * 256 + c], wsum[(size_t)b * 256 + c], acc);
    const float sw = swp[h];
    r6[(size_t)b * 6 + 3 * h + j] = acc + bn[j] * sw + bp[0];
  }
};
// small gradients of the tail (oracle: manual_backward, "conv_p / neck"): e[b, c] = sum_j Wn[j, c] dr[b, j];
// dWn[j, c] += sum_b dr[b, j] wsum[b, c]; dbn[j] += sum_b dr[b, j] sum_p wp; dbp += sum_{b, j} dr[b, j].
// grid (1), nt = 256 (thread = channel c)
struct KRotTailBwd {
  const float *d_r6, *wsum, *wn, *swp; float *e, *dwn, *dbn, *dbp; int B, P, h;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.tx;
    if (c >= 256) return;
    float gw[3] = {0.0f, 0.0f, 0.0f};
    for (int b = 0; b < B; ++b) {
      const float* dr = d_r6 + (size_t)b * 6 + 3 * h;
      e[(size_t)b * 256 + c] = wn[c] * dr[0] + wn[256 + c] * dr[1] + wn[512 + c] * dr[2];
      for (int j = 0; j < 3; ++j) gw[j] = fmaf(dr[j], wsum[(size_t)b * 256 + c], gw[j]);
    }
    for (int j = 0; j < 3; ++j) dwn[j * 256 + c] += gw[j];
    if (c < 3) {
      const float sw = swp[h];
      float sd = 0.0f;
      for (int b = 0; b < B; ++b) sd += d_r6[(size_t)b * 6 + 3 * h + c];
      dbn[c] += sd * sw;
    }
    if (c == 3) {
      float sd = 0.0f;
      for (int b = 0; b < B; ++b) for (int j = 0; j < 3; ++j) sd += d_r6[(size_t)b * 6 + 3 * h + j];
      dbp[0] += sd;
    }
  }
};
// du1[b, p, c] = wp[p] e[b, c]  (rank one).  grid (ceil(B * P * 256 / nt))
struct KRotDu1 {
  const float *wp, *e; float* du; int P; long long n;
  TK_HD void operator()(const Idx& i) const {
    const long long x = (long long)i.bx * i.nt + i.tx;
    if (x >= n) return;
    const int c = (int)(x & 255);
    const long long bp = x >> 8;
    du[x] = wp[bp % P] * e[(size_t)(bp / P) * 256 + c];
  }
};
// dwp[p] += sum_b (u1[b, p, :] . e[b, :] + dr[b, :] . bn).  grid (ceil(P / nt))
struct KRotDwp {
  const float *u1, *e, *d_r6, *bn; float* dwp; int B, P, h;
  TK_HD void operator()(const Idx& i) const {
    const int p = i.bx * i.nt + i.tx;
    if (p >= P) return;
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) {
      const float* u = u1 + ((size_t)b * P + p) * 256;
      const float* eb = e + (size_t)b * 256;
      float a = 0.0f;
      for (int c = 0; c < 256; ++c) a = fmaf(u[c], eb[c], a);
      const float* dr = d_r6 + (size_t)b * 6 + 3 * h;
      acc += a + dr[0] * bn[0] + dr[1] * bn[1] + dr[2] * bn[2];
    }
    dwp[p] += acc;
  }
};

// The same in two stages: part[b, p, q] = u1[b, p, 32 q .. 32 q + 31] . e[b, same] (one thread per 128-byte line of u1; grid
// (ceil(8 P / nt), B)), then dwp[p] += sum_b (sum_q part[b, p, q] + dr[b, :] . bn) (grid (ceil(P / nt))).
struct KRotDwpPart {
  const float *u1, *e; float* part; int P;
  TK_HD void operator()(const Idx& i) const {
    const int x = i.bx * i.nt + i.tx, b = i.by;
    if (x >= 8 * P) return;
    const int p = x >> 3, q = x & 7;
    const float* u = u1 + ((size_t)b * P + p) * 256 + q * 32;
    const float* eb = e + (size_t)b * 256 + q * 32;
    float a = 0.0f;
    for (int c = 0; c < 32; ++c) a = fmaf(u[c], eb[c], a);
    part[((size_t)b * P + p) * 8 + q] = a;
  }
};
struct KRotDwpSum {
  const float *part, *d_r6, *bn; float* dwp; int B, P, h;
  TK_HD void operator()(const Idx& i) const {
    const int p = i.bx * i.nt + i.tx;
    if (p >= P) return;
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) {
      const float* pp = part + ((size_t)b * P + p) * 8;
      float a = 0.0f;
      for (int q = 0; q < 8; ++q) a += pp[q];
      const float* dr = d_r6 + (size_t)b * 6 + 3 * h;
      acc += a + dr[0] * bn[0] + dr[1] * bn[1] + dr[2] * bn[2];
    }
    dwp[p] += acc;
  }
};

// per-set 3 x 3 products over the set's points, stage 1: part[s, ch, 3 i + j] = sum over the chunk's points n of
// a[s, n, i] b[s, n, j]  (dT3 = q^T dq').  grid (1, chunks, S), nt >= 9; stage 2 is KColSum with rows_per = chunks.
struct KSet3x3Part {
  const float *a, *b; float* part; int N, per;
  TK_HD void operator()(const Idx& i) const {
    const int ij = i.tx, ch = i.by, s = i.bz;
    if (ij >= 9) return;
    const int n0 = ch * per, n1 = n0 + per < N ? n0 + per : N;
    const float* pa = a + (size_t)s * N * 3 + ij / 3;
    const float* pb = b + (size_t)s * N * 3 + ij % 3;
    float acc = 0.0f;
    for (int n = n0; n < n1; ++n) acc = fmaf(pa[(size_t)n * 3], pb[(size_t)n * 3], acc);
    part[((size_t)s * ((N + per - 1) / per) + ch) * 9 + ij] = acc;
  }
};

// ---- ts-head input gather / gradient scatter (CATRE_disR_shared.py:66-82): ts_in[b] = [g(2b) | pfmax(2b) | s_b]
//      grid (ceil(1091 / nt), B)
struct KTsGather {
  const float *g, *pfmax, *scale; float* ts_in;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, b = i.by;
    if (c >= 1091) return;
    float v;
    if (c < 1024) v = g[(size_t)(2 * b) * 1024 + c];
    else if (c < 1088) v = pfmax[(size_t)(2 * b) * 64 + (c - 1024)];
    else v = scale[(size_t)b * 3 + (c - 1088)];
    ts_in[(size_t)b * 1091 + c] = v;
  }
};
// dg[2b, :] += din[b, :1024], dpfmax[2b, :] += din[b, 1024:1088] (both zeroed at the start of the backward; dg already holds the
// rotation heads' share; the prior sets' rows get nothing; the initial scale is detached).  grid (ceil(1088 / nt), B)
struct KTsScatter {
  const float* din; float *dg, *dpfmax;
  TK_HD void operator()(const Idx& i) const {
    const int c = i.bx * i.nt + i.tx, b = i.by;
    if (c >= 1088) return;
    const float v = din[(size_t)b * 1091 + c];
    if (c < 1024) dg[(size_t)(2 * b) * 1024 + c] += v;
    else dpfmax[(size_t)(2 * b) * 64 + (c - 1024)] += v;
  }
};

// ---- rot6d Gram-Schmidt + pose update (core/utils/rot_reps.py:34-55, pose_scale_from_delta_init.py:47-95).
//      grid (ceil(B / nt))
struct KPoseFwd {
  const float *r6, *dts, *pose_in, *scale_in, *K; float *pose_out, *scale_out; int B;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx * i.nt + i.tx;
    if (b >= B) return;
    const float* r = r6 + (size_t)b * 6;
    const float* P = pose_in + (size_t)b * 12;
    const float* d = dts + (size_t)b * 6;
    float na = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]); na = na > 1e-12f ? na : 1e-12f;
    const float x[3] = {r[0] / na, r[1] / na, r[2] / na};
    float c[3] = {x[1] * r[5] - x[2] * r[4], x[2] * r[3] - x[0] * r[5], x[0] * r[4] - x[1] * r[3]};
    float nc = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]); nc = nc > 1e-12f ? nc : 1e-12f;
    const float z[3] = {c[0] / nc, c[1] / nc, c[2] / nc};
    const float y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
    float* O = pose_out + (size_t)b * 12;
    for (int r_ = 0; r_ < 3; ++r_)
      for (int cc = 0; cc < 3; ++cc)  // R' = dR . R, dR = [x y z] as columns
        O[r_ * 4 + cc] = x[r_] * P[0 * 4 + cc] + y[r_] * P[1 * 4 + cc] + z[r_] * P[2 * 4 + cc];
    const float fx = K[(size_t)b * 9 + 0], fy = K[(size_t)b * 9 + 4];
    const float tz = P[2 * 4 + 3], zn = d[2] * tz;
    O[0 * 4 + 3] = zn * (d[0] / fx + P[0 * 4 + 3] / tz);
    O[1 * 4 + 3] = zn * (d[1] / fy + P[1 * 4 + 3] / tz);
    O[2 * 4 + 3] = zn;
    for (int j = 0; j < 3; ++j) scale_out[(size_t)b * 3 + j] = scale_in[(size_t)b * 3 + j] + d[3 + j];
  }
};

// ---- losses of the shipped LOSS_CFG and their gradients w.r.t. the predicted (R, t, s)
//      (CATRE_disR_shared.py:168-288, core/catre/losses/pm_loss.py:110-130, rot_loss.py:45-58,
//      core/utils/pose_utils.py:472-528; oracle: catre_loss / loss_backward).  One thread per object.
//      lossp [B, 6] = this object's share of (PM_R, rot, yaxis_rot, trans_xy, trans_z, scale);
//      dpose [B, 15] = d/d(R' row-major 9, t' 3, s' 3).  grid (ceil(B / nt))
// Three stages: (1) KLossSel, one thread per object: the ground-truth rotation the point-matching term compares with (for a
// symmetric object the closest of its symmetric copies) -> gs [B, 9]; (2) KLossPm, one thread per (object, chunk of points): the
// chunk's share of the point-matching loss and of its gradients -> pm [B, chunks, 13] doubles (|diff| sum, dR 9, ds 3);
// (3) KLoss, one thread per object: sums the chunks in order and adds the rotation / translation / scale terms.
// grids: (ceil(B / nt)), (chunks, B) with nt = 1 .. any (thread tx > 0 idle) -- launched as (ceil(chunks / nt), B), (ceil(B / nt))
// (1) runs in two launches so that the search over the symmetric copies (313 rotations with the shipped loader step) is not one
// serial loop per object: KLossSelPart, one thread per (object, chunk of rotations): the chunk's first best clamped cosine and its
// index -> part [B, chunks, 2]; KLossSel, one thread per object: merges the chunks in order (strict improvement only, starting
// from the unrotated ground truth), i.e. the first maximum in rotation order -- the serial loop's choice -- and rebuilds that copy.
TK_HD float tk_sym_cos(const float* R, const float* G, const float* S, float* C) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[r * 3 + c] = G[r * 3] * S[c] + G[r * 3 + 1] * S[3 + c] + G[r * 3 + 2] * S[6 + c];
  float t2 = 0.0f;
  for (int e = 0; e < 9; ++e) t2 += R[e] * C[e];
  return fminf(1.0f, fmaxf(-1.0f, 0.5f * ((t2 <= 3.0f ? t2 : 3.0f) - 1.0f)));
}
// grid (ceil(chunks / nt), B)
struct KLossSelPart {
  const float *pose, *gt_pose, *sym_rots; const unsigned char* is_sym; float* part; int B, n_rots, chunks, per;
  TK_HD void operator()(const Idx& i) const {
    const int ch = i.bx * i.nt + i.tx, b = i.by;
    if (ch >= chunks || is_sym[b] == 0) return;
    const float* Pp = pose + (size_t)b * 12;
    const float* Gp = gt_pose + (size_t)b * 12;
    float R[9], G[9], C[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r * 3 + c] = Pp[r * 4 + c]; G[r * 3 + c] = Gp[r * 4 + c]; }
    float best = -2.0f;  // below every clamped cosine
    int bk = -1;
    const int k0 = ch * per, k1 = k0 + per < n_rots ? k0 + per : n_rots;
    for (int k = k0; k < k1; ++k) {
      const float cs = tk_sym_cos(R, G, sym_rots + (size_t)k * 9, C);
      if (cs > best) { best = cs; bk = k; }
    }
    part[((size_t)b * chunks + ch) * 2] = best;
    part[((size_t)b * chunks + ch) * 2 + 1] = tk_i2f(bk);
  }
};
// grid (ceil(B / nt))
struct KLossSel {
  const float *pose, *gt_pose, *sym_rots; const unsigned char* is_sym; const float* part; float* gs; int B, n_rots, chunks;
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx * i.nt + i.tx;
    if (b >= B) return;
    const float* Pp = pose + (size_t)b * 12;
    const float* Gp = gt_pose + (size_t)b * 12;
    float R[9], G[9], Gs[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r * 3 + c] = Pp[r * 4 + c]; G[r * 3 + c] = Gp[r * 4 + c]; Gs[r * 3 + c] = G[r * 3 + c]; }
    if (is_sym[b] != 0) {  // closest symmetric ground truth: largest clamped cosine of the rotation error, strict improvement only
      float tr = 0.0f;
      for (int e = 0; e < 9; ++e) tr += R[e] * G[e];
      float best = fminf(1.0f, fmaxf(-1.0f, 0.5f * ((tr <= 3.0f ? tr : 3.0f) - 1.0f)));
      int bk = -1;
      for (int ch = 0; ch < chunks; ++ch) {
        const float cs = part[((size_t)b * chunks + ch) * 2];
        const int k = tk_f2i(part[((size_t)b * chunks + ch) * 2 + 1]);
        if (k >= 0 && cs > best) { best = cs; bk = k; }
      }
      if (bk >= 0) tk_sym_cos(R, G, sym_rots + (size_t)bk * 9, Gs);
    }
    for (int e = 0; e < 9; ++e) gs[(size_t)b * 9 + e] = Gs[e];
  }
};
struct KLossPm {
  const float *pose, *scale, *gt_scale, *kps, *gs; double* pm; int B, N, chunks, per; float w_pm;
  TK_HD void operator()(const Idx& i) const {
    const int ch = i.bx * i.nt + i.tx, b = i.by;
    if (ch >= chunks) return;
    const float* Pp = pose + (size_t)b * 12;
    float R[9], Gs[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r * 3 + c] = Pp[r * 4 + c]; Gs[r * 3 + c] = gs[(size_t)b * 9 + r * 3 + c]; }
    const float* s = scale + (size_t)b * 3;
    const float* sg = gt_scale + (size_t)b * 3;
    float dR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ds[3] = {0, 0, 0};
    double pmv = 0.0;
    const float inv_bn = w_pm / ((float)B * (float)N);
    const int n0 = ch * per, n1 = n0 + per < N ? n0 + per : N;
    for (int n = n0; n < n1; ++n) {
      const float* k = kps + ((size_t)b * N + n) * 3;
      const float sk[3] = {k[0] * s[0], k[1] * s[1], k[2] * s[2]};
      const float gk[3] = {k[0] * sg[0], k[1] * sg[1], k[2] * sg[2]};
      float de[3];
      for (int r = 0; r < 3; ++r) {
        const float diff = (R[r * 3] * sk[0] + R[r * 3 + 1] * sk[1] + R[r * 3 + 2] * sk[2]) -
                           (Gs[r * 3] * gk[0] + Gs[r * 3 + 1] * gk[1] + Gs[r * 3 + 2] * gk[2]);
        pmv += fabsf(diff);
        de[r] = tk_sign(diff) * inv_bn;
        for (int c = 0; c < 3; ++c) dR[r * 3 + c] += de[r] * sk[c];
      }
      for (int c = 0; c < 3; ++c) ds[c] += (R[c] * de[0] + R[3 + c] * de[1] + R[6 + c] * de[2]) * k[c];
    }
    double* o = pm + ((size_t)b * chunks + ch) * 13;
    o[0] = pmv;
    for (int e = 0; e < 9; ++e) o[1 + e] = dR[e];
    for (int c = 0; c < 3; ++c) o[10 + c] = ds[c];
  }
};
struct KLoss {
  const float *pose, *scale, *gt_pose, *gt_scale; const double* pmp; const unsigned char* is_sym;
  float *lossp, *dpose; int B, N, chunks, n_sym, n_nosym;
  float w_pm, w_rot, w_trans, w_scale;  // LOSS_CFG.PM_LW, ROT_LW, TRANS_LW, SCALE_LW (all 1 in the shipped config)
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx * i.nt + i.tx;
    if (b >= B) return;
    const float* Pp = pose + (size_t)b * 12;
    const float* Gp = gt_pose + (size_t)b * 12;
    float R[9], G[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r * 3 + c] = Pp[r * 4 + c]; G[r * 3 + c] = Gp[r * 4 + c]; }
    const bool sym = is_sym[b] != 0;
    const float* s = scale + (size_t)b * 3;
    const float* sg = gt_scale + (size_t)b * 3;
    double acc[13];
    for (int e = 0; e < 13; ++e) acc[e] = 0.0;
    for (int ch = 0; ch < chunks; ++ch)
      for (int e = 0; e < 13; ++e) acc[e] += pmp[((size_t)b * chunks + ch) * 13 + e];
    float dR[9], ds[3];
    for (int e = 0; e < 9; ++e) dR[e] = (float)acc[1 + e];
    for (int c = 0; c < 3; ++c) ds[c] = (float)acc[10 + c];
    const double pm = acc[0];
    float* L = lossp + (size_t)b * 6;
    L[0] = w_pm * (float)(pm / ((double)B * N));  // PM_LW * 3 * mean over B*N*3
    L[1] = L[2] = 0.0f;
    if (sym) {
      float a = 0.0f;
      for (int r = 0; r < 3; ++r) { const float d = R[r * 3 + 1] - G[r * 3 + 1]; a += fabsf(d); dR[r * 3 + 1] += w_rot * tk_sign(d) / (3.0f * n_sym); }
      L[2] = w_rot * a / (3.0f * n_sym);
    } else {
      float tr = 0.0f;
      for (int e = 0; e < 9; ++e) { tr += R[e] * G[e]; dR[e] += -w_rot * G[e] / (4.0f * n_nosym); }
      L[1] = w_rot * (3.0f - tr) * 0.25f / n_nosym;
    }
    float* D = dpose + (size_t)b * 15;
    for (int e = 0; e < 9; ++e) D[e] = dR[e];
    const float dx = Pp[3] - Gp[3], dy = Pp[7] - Gp[7], dz = Pp[11] - Gp[11];
    L[3] = w_trans * (fabsf(dx) + fabsf(dy)) / (2.0f * B);
    L[4] = w_trans * fabsf(dz) / (float)B;
    D[9] = w_trans * tk_sign(dx) / (2.0f * B); D[10] = w_trans * tk_sign(dy) / (2.0f * B); D[11] = w_trans * tk_sign(dz) / (float)B;
    float ls = 0.0f;
    for (int c = 0; c < 3; ++c) { const float d = s[c] - sg[c]; ls += fabsf(d); D[12 + c] = ds[c] + w_scale * tk_sign(d) / (3.0f * B); }
    L[5] = w_scale * ls / (3.0f * B);
  }
};
// losses[j] = sum_b lossp[b, j] in object order.  grid (1), nt >= 6
struct KLossSum {
  const float* lossp; float* losses; int B;
  TK_HD void operator()(const Idx& i) const {
    if (i.tx >= 6) return;
    float a = 0.0f;
    for (int b = 0; b < B; ++b) a += lossp[(size_t)b * 6 + i.tx];
    losses[i.tx] = a;
  }
};

// ---- backward of the pose update and the Gram-Schmidt (oracle: pose_backward): dpose [B, 15] -> d_r6 [B, 6],
//      d_dts [B, 6] = (d Delta_t, d Delta_s).  grid (ceil(B / nt))
struct KPoseBwd {
  const float *dpose, *r6, *dts, *pose_in, *K; float *d_r6, *d_dts; int B;
  TK_HD static void cross(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
  }
  TK_HD void operator()(const Idx& i) const {
    const int b = i.bx * i.nt + i.tx;
    if (b >= B) return;
    const float* D = dpose + (size_t)b * 15;
    const float* P = pose_in + (size_t)b * 12;
    const float* d = dts + (size_t)b * 6;
    const float* r = r6 + (size_t)b * 6;
    // d(dR) = dR' . R^T  (R' = dR . R); columns gx, gy, gz of d(dR)
    float g[9];
    for (int rr = 0; rr < 3; ++rr)
      for (int cc = 0; cc < 3; ++cc) g[rr * 3 + cc] = D[rr * 3] * P[cc * 4] + D[rr * 3 + 1] * P[cc * 4 + 1] + D[rr * 3 + 2] * P[cc * 4 + 2];
    float gx[3] = {g[0], g[3], g[6]}, gy[3] = {g[1], g[4], g[7]}, gz[3] = {g[2], g[5], g[8]};
    float na = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]); na = na > 1e-12f ? na : 1e-12f;
    const float x[3] = {r[0] / na, r[1] / na, r[2] / na};
    const float bv[3] = {r[3], r[4], r[5]};
    float c[3]; cross(x, bv, c);
    float nc = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]); nc = nc > 1e-12f ? nc : 1e-12f;
    const float z[3] = {c[0] / nc, c[1] / nc, c[2] / nc};
    float t[3];
    cross(x, gy, t); for (int j = 0; j < 3; ++j) gz[j] += t[j];   // y = z x x
    cross(gy, z, t); for (int j = 0; j < 3; ++j) gx[j] += t[j];
    const float zg = z[0] * gz[0] + z[1] * gz[1] + z[2] * gz[2];
    float gc[3]; for (int j = 0; j < 3; ++j) gc[j] = (gz[j] - z[j] * zg) / nc;
    cross(bv, gc, t); for (int j = 0; j < 3; ++j) gx[j] += t[j];  // c = x x b
    float gb[3]; cross(gc, x, gb);
    const float xg = x[0] * gx[0] + x[1] * gx[1] + x[2] * gx[2];
    float* o = d_r6 + (size_t)b * 6;
    for (int j = 0; j < 3; ++j) { o[j] = (gx[j] - x[j] * xg) / na; o[3 + j] = gb[j]; }
    const float fx = K[(size_t)b * 9], fy = K[(size_t)b * 9 + 4];
    const float tx = P[3], ty = P[7], tz = P[11], zn = d[2] * tz;
    const float dzt = D[11] + D[9] * (d[0] / fx + tx / tz) + D[10] * (d[1] / fy + ty / tz);
    float* q = d_dts + (size_t)b * 6;
    q[0] = D[9] * zn / fx; q[1] = D[10] * zn / fy; q[2] = dzt * tz;
    q[3] = D[12]; q[4] = D[13]; q[5] = D[14];
  }
};

}  // namespace catre_train

#ifndef CATRE_HOST_EMU
namespace catre_train {
template <class KF>
__global__ void tk_run(KF k) {
  k(Idx{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x, (int)blockDim.x});
}

}  // namespace catre_train
#include "train_gemm_tiled.cuh"
#endif

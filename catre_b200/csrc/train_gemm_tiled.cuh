// Shared-memory tiled GEMM of the training step (see train_kernels.cuh: GemmP / KGemmNaive).  Kept in its own header so
// tests/emu/gemm_tiled_emu.cpp can compile this very kernel for the CPU (one OS thread per CUDA thread, a barrier for
// __syncthreads) and compare it with KGemmNaive over the stride / batch / split-K combinations the chain uses.
#pragma once
namespace catre_train {
// Shared-memory tiled version of KGemmNaive (same parameters and results up to summation order): 64 x 64 output
// tile, 16-deep k slabs, 256 threads with a 4 x 4 register tile each.  grid (ceil(M/64), ceil(N/64), batch * splits)
__global__ void __launch_bounds__(256) tk_gemm_tiled(GemmP p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  int z = blockIdx.z, k0 = 0, k1 = p.K;
  if (p.splits > 1) { k0 = z * p.k_per; k1 = min(k0 + p.k_per, p.K); z = 0; }
  const float* A = p.A + (long long)z * p.sab;
  const float* Bm = p.B + (long long)z * p.sbb;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const bool a_kfast = p.sak == 1, b_nfast = p.sbn == 1;
  // the next slab's 4 + 4 operand values per thread are loaded into registers underneath the FMAs of the current slab: the
  // chain's small launches (one or a few CTAs, 16 slabs each) are bound by the latency of this load, not by arithmetic
  float ra[4], rb[4];
  auto load_slab = [&](int kb) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int idx = tid + l * 256;
      int m, k;
      if (a_kfast) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      ra[l] = (m0 + m < p.M && kb + k < k1) ? A[(long long)(m0 + m) * p.sam + (long long)(kb + k) * p.sak] : 0.0f;
      int n, k2;
      if (b_nfast) { k2 = idx >> 6; n = idx & 63; } else { k2 = idx & 15; n = idx >> 4; }
      rb[l] = (n0 + n < p.N && kb + k2 < k1) ? Bm[(long long)(kb + k2) * p.sbk + (long long)(n0 + n) * p.sbn] : 0.0f;
    }
  };
  if (k0 < k1) load_slab(k0);
  for (int kb = k0; kb < k1; kb += 16) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int idx = tid + l * 256;
      int m, k;
      if (a_kfast) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      As[k][m] = ra[l];
      int n, k2;
      if (b_nfast) { k2 = idx >> 6; n = idx & 63; } else { k2 = idx & 15; n = idx >> 4; }
      Bs[k2][n] = rb[l];
    }
    __syncthreads();
    if (kb + 16 < k1) load_slab(kb + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.splits > 1) { p.partial[((size_t)blockIdx.z * p.M + m) * p.N + n] = v; continue; }
      if (p.bias) v += p.bias[n + (long long)z * p.sbias_b];
      if (p.relu) v = fmaxf(v, 0.0f);
      float* c = p.C + (long long)z * p.scb + (long long)m * p.scm + (long long)n * p.scn;
      *c = p.accumulate ? *c + v : v;
    }
  }
}
}  // namespace catre_train

namespace catre_train {
// Second version (opt-in, CATRE_TRAIN_GEMM=v2): 128 x BN output tile (BN = 128 or 64), 8-deep k slabs, 256 threads with an
// 8 x (BN/16) register tile each, the next slab prefetched into registers underneath the FMAs of the current one.  Same
// parameters and results (up to summation order) as tk_gemm_tiled.  grid (ceil(M/128), ceil(N/BN), batch * splits)
template <int BN>
__global__ void __launch_bounds__(256) tk_gemm_tiled2(GemmP p) {
  constexpr int BM = 128, BK = 8, TN = BN / 16, NB = BN / 32;  // NB = B elements each thread stages per slab
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int z = blockIdx.z, k0 = 0, k1 = p.K;
  if (p.splits > 1) { k0 = z * p.k_per; k1 = min(k0 + p.k_per, p.K); z = 0; }
  const float* A = p.A + (long long)z * p.sab;
  const float* Bm = p.B + (long long)z * p.sbb;
  const bool a_kfast = p.sak == 1, b_nfast = p.sbn == 1;
  // staging coordinates of this thread inside a slab (fixed over the k loop)
  int am[4], ak[4], bn[NB], bk[NB];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int idx = tid + l * 256;
    if (a_kfast) { am[l] = idx >> 3; ak[l] = idx & 7; } else { am[l] = idx & (BM - 1); ak[l] = idx >> 7; }
  }
#pragma unroll
  for (int l = 0; l < NB; ++l) {
    const int idx = tid + l * 256;
    if (b_nfast) { bk[l] = idx / BN; bn[l] = idx % BN; } else { bk[l] = idx & 7; bn[l] = idx >> 3; }
  }
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
  float ra[4], rb[NB];
#pragma unroll
  for (int l = 0; l < 4; ++l)
    ra[l] = (m0 + am[l] < p.M && k0 + ak[l] < k1) ? A[(long long)(m0 + am[l]) * p.sam + (long long)(k0 + ak[l]) * p.sak] : 0.0f;
#pragma unroll
  for (int l = 0; l < NB; ++l)
    rb[l] = (n0 + bn[l] < p.N && k0 + bk[l] < k1) ? Bm[(long long)(k0 + bk[l]) * p.sbk + (long long)(n0 + bn[l]) * p.sbn] : 0.0f;
  for (int kb = k0; kb < k1; kb += BK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) As[ak[l]][am[l]] = ra[l];
#pragma unroll
    for (int l = 0; l < NB; ++l) Bs[bk[l]][bn[l]] = rb[l];
    __syncthreads();
    const int kn = kb + BK;
    if (kn < k1) {  // prefetch the next slab
#pragma unroll
      for (int l = 0; l < 4; ++l)
        ra[l] = (m0 + am[l] < p.M && kn + ak[l] < k1) ? A[(long long)(m0 + am[l]) * p.sam + (long long)(kn + ak[l]) * p.sak] : 0.0f;
#pragma unroll
      for (int l = 0; l < NB; ++l)
        rb[l] = (n0 + bn[l] < p.N && kn + bk[l] < k1) ? Bm[(long long)(kn + bk[l]) * p.sbk + (long long)(n0 + bn[l]) * p.sbn] : 0.0f;
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[TN];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; a[4 + i] = As[kk][64 + ty * 4 + i]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
      if (TN == 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) b[(TN == 8 ? 4 : 0) + j] = Bs[kk][(BN / 2) + tx * 4 + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : (BN / 2) + tx * 4 + (j - 4));
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.splits > 1) { p.partial[((size_t)blockIdx.z * p.M + m) * p.N + n] = v; continue; }
      if (p.bias) v += p.bias[n + (long long)z * p.sbias_b];
      if (p.relu) v = fmaxf(v, 0.0f);
      float* c = p.C + (long long)z * p.scb + (long long)m * p.scm + (long long)n * p.scn;
      *c = p.accumulate ? *c + v : v;
    }
  }
}
}  // namespace catre_train

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
  for (int kb = k0; kb < k1; kb += 16) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int idx = tid + l * 256;
      int m, k;
      if (a_kfast) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      float v = 0.0f;
      if (m0 + m < p.M && kb + k < k1) v = A[(long long)(m0 + m) * p.sam + (long long)(kb + k) * p.sak];
      As[k][m] = v;
      int n, k2;
      if (b_nfast) { k2 = idx >> 6; n = idx & 63; } else { k2 = idx & 15; n = idx >> 4; }
      v = 0.0f;
      if (n0 + n < p.N && kb + k2 < k1) v = Bm[(long long)(kb + k2) * p.sbk + (long long)(n0 + n) * p.sbn];
      Bs[k2][n] = v;
    }
    __syncthreads();
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

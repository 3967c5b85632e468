// FP32 FMA issue-rate probe on B200: scalar FFMA (imm and 3-register forms) vs packed FFMA2 (fma.rn.f32x2),
// plus the engine's GELU polynomial in scalar and packed form.  Prints FMA/clk/SM.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

constexpr int CH = 8;  // independent chains per thread (elements for scalar, pairs for packed)

template <int MODE>
__global__ void probe(float* out, int iters, float seed) {
  float z[2 * CH], p[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) { z[i] = seed + 0.001f * (threadIdx.x + i); p[i] = 0.5f * z[i]; }
  if (MODE == 0) {  // scalar imm-form: p = p*z + const
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) p[i] = fmaf(p[i], z[i], 0.0070524085f);
    }
  } else if (MODE == 1) {  // scalar 3-register form: p = p*z + z
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) p[i] = fmaf(p[i], z[i], z[(i + 1) % (2 * CH)]);
    }
  } else if (MODE == 2 || MODE == 3) {  // packed, imm (2) or register addend (3)
    unsigned long long P[CH], Z[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { P[i] = pk(p[2 * i], p[2 * i + 1]); Z[i] = pk(z[2 * i], z[2 * i + 1]); }
    const unsigned long long C = pk(0.0070524085f, 0.0070524085f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < CH; ++i) P[i] = fma2(P[i], Z[i], MODE == 2 ? C : Z[(i + 1) % CH]);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) upk(P[i], p[2 * i], p[2 * i + 1]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) s += p[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the engine's GELU (simt_kernels.cuh gelu_fast), scalar vs packed polynomial
__device__ __forceinline__ float gelu_s(float x) {
  const float z = fminf(fabsf(x), 5.939697f);
  float p = fmaf(2.3988423e-07f, z, -5.5500227e-06f);
  p = fmaf(p, z, 5.128636e-05f); p = fmaf(p, z, -0.00021167348f); p = fmaf(p, z, -0.00010999188f);
  p = fmaf(p, z, 0.0070524085f); p = fmaf(p, z, -0.052498225f); p = fmaf(p, z, -0.45920548f);
  p = fmaf(p, z, -1.1511058f); p = fmaf(p, z, -1.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  return fmaf(-fabsf(x), e, fmaxf(x, 0.0f));
}
__device__ __forceinline__ void gelu_p(float& x0, float& x1) {
  const float z0 = fminf(fabsf(x0), 5.939697f), z1 = fminf(fabsf(x1), 5.939697f);
  const unsigned long long Z = pk(z0, z1);
  unsigned long long P = fma2(pk(2.3988423e-07f, 2.3988423e-07f), Z, pk(-5.5500227e-06f, -5.5500227e-06f));
  P = fma2(P, Z, pk(5.128636e-05f, 5.128636e-05f)); P = fma2(P, Z, pk(-0.00021167348f, -0.00021167348f));
  P = fma2(P, Z, pk(-0.00010999188f, -0.00010999188f)); P = fma2(P, Z, pk(0.0070524085f, 0.0070524085f));
  P = fma2(P, Z, pk(-0.052498225f, -0.052498225f)); P = fma2(P, Z, pk(-0.45920548f, -0.45920548f));
  P = fma2(P, Z, pk(-1.1511058f, -1.1511058f)); P = fma2(P, Z, pk(-1.0f, -1.0f));
  float p0, p1, e0, e1;
  upk(P, p0, p1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p1));
  x0 = fmaf(-fabsf(x0), e0, fmaxf(x0, 0.0f));
  x1 = fmaf(-fabsf(x1), e1, fmaxf(x1, 0.0f));
}

template <int MODE>
__global__ void gelu_probe(float* out, int iters, float seed) {
  float v[2 * CH];
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) v[i] = seed + 0.01f * (threadIdx.x % 64) - 0.3f * i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) v[i] = gelu_s(v[i]) + 0.25f;
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) { gelu_p(v[2 * i], v[2 * i + 1]); v[2 * i] += 0.25f; v[2 * i + 1] += 0.25f; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount, blocks = sms * 4, threads = 256, iters = 4096;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  const double fma_total = (double)blocks * threads * iters * 2 * CH;
  const char* names[4] = {"scalar FFMA imm", "scalar FFMA 3-reg", "packed FFMA2 imm", "packed FFMA2 3-reg"};
  for (int m = 0; m < 4; ++m) {
    float ms = 0;
    if (m == 0) ms = time_ms([&] { probe<0><<<blocks, threads>>>(out, iters, 0.3f); });
    if (m == 1) ms = time_ms([&] { probe<1><<<blocks, threads>>>(out, iters, 0.3f); });
    if (m == 2) ms = time_ms([&] { probe<2><<<blocks, threads>>>(out, iters, 0.3f); });
    if (m == 3) ms = time_ms([&] { probe<3><<<blocks, threads>>>(out, iters, 0.3f); });
    double clk = ms * 1e-3 * clk_khz * 1e3;
    printf("%-20s %8.3f ms  %7.1f FMA/clk/SM (at %d MHz nominal)\n", names[m], ms, fma_total / clk / sms, clk_khz / 1000);
  }
  for (int m = 0; m < 2; ++m) {
    float ms = m == 0 ? time_ms([&] { gelu_probe<0><<<blocks, threads>>>(out, iters, 0.7f); })
                      : time_ms([&] { gelu_probe<1><<<blocks, threads>>>(out, iters, 0.7f); });
    double clk = ms * 1e-3 * clk_khz * 1e3;
    printf("%-20s %8.3f ms  %7.2f GELU/clk/SM\n", m == 0 ? "GELU scalar" : "GELU packed poly", ms, fma_total / clk / sms);
  }
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

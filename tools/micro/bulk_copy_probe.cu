// Microbenchmark: how fast ONE CTA can stream a contiguous, L2-resident 256 KB weight slice into shared memory (the FC-chain
// kernels' weight stream), as a function of the mechanism:
//   bulk   cp.async.bulk (1-D, mbarrier complete_tx) in `chunk`-byte pieces through a ring of `slots`, issued by `lanes`
//          lanes of one producer warp (each lane a 1/lanes part of every chunk); a consumer warp frees a slot at once
//   ldg    512 threads, 16-byte ld.global.nc + st.shared, 8 loads in flight per thread
// `ctas` CTAs run at once (8 = one FC cluster, 64 = the chains of a 64-object launch), each on its own slice.
// Prints bytes/clk per CTA and microseconds for the 256 KB.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../catre_b200/csrc/tc_kernels.cuh"
using namespace catre;

constexpr int SLICE = 256 * 1024;

__global__ void __launch_bounds__(576, 1) bulk_probe(const char* __restrict__ w, int chunk, int slots, int lanes, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t base = smem_u32(smem), bar_full = base, bar_empty = base + 256, ring = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < slots; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const char* src = w + (size_t)blockIdx.x * SLICE;
  const int n = SLICE / chunk;
  const long long t0 = clock64();
  if (warp == 0 && lane < lanes) {
    const uint32_t q = chunk / lanes;
    for (int g = 0; g < n; ++g) {
      const int slot = g % slots;
      if (g >= slots) mbar_wait(bar_empty + 8 * slot, (uint32_t)(((g / slots) - 1) & 1));
      if (lane == 0) mbar_expect_tx(bar_full + 8 * slot, chunk);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ring + slot * chunk + lane * q), "l"(src + (size_t)g * chunk + (size_t)lane * q), "r"(q), "r"(bar_full + 8 * slot) : "memory");
    }
  } else if (warp == 1 && lane == 0) {
    for (int g = 0; g < n; ++g) {
      const int slot = g % slots;
      mbar_wait(bar_full + 8 * slot, (uint32_t)((g / slots) & 1));
      mbar_arrive(bar_empty + 8 * slot);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(512, 1) ldg_probe(const char* __restrict__ w, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t ring = smem_u32(smem) + 1024;
  const uint4* src = reinterpret_cast<const uint4*>(w + (size_t)blockIdx.x * SLICE);
  __syncthreads();
  const long long t0 = clock64();
  for (int i0 = 0; i0 < SLICE / 16; i0 += 512 * 8) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(src + i0 + u * 512 + threadIdx.x);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t a = ring + (uint32_t)(((i0 + u * 512 + threadIdx.x) * 16) % (96 * 1024));
      st_shared_v4(a, v[u].x, v[u].y, v[u].z, v[u].w);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

static void report(const char* what, int ctas, long long* dout, double mhz) {
  long long h[148];
  cudaMemcpy(h, dout, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < ctas; ++i) mean += h[i];
  mean /= ctas;
  printf("{\"probe\": \"bulk_copy\", \"what\": \"%s\", \"ctas\": %d, \"cycles\": %.0f, \"bytes_per_clk_per_cta\": %.1f, \"us_at_%.0fMHz\": %.2f}\n", what, ctas,
         mean, SLICE / mean, mhz, mean / mhz);
}

int main() {
  char* w; long long* dout;
  cudaMalloc(&w, (size_t)148 * SLICE);
  cudaMemset(w, 1, (size_t)148 * SLICE);
  cudaMalloc(&dout, 148 * sizeof(long long));
  cudaFuncSetAttribute(bulk_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ldg_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double mhz = clk_khz / 1000.0;
  const int cta_counts[3] = {8, 64, 148};
  for (int ci = 0; ci < 3; ++ci) {
    const int ctas = cta_counts[ci];
    const int cfgs[7][3] = {{32768, 3, 1}, {32768, 3, 4}, {16384, 6, 1}, {16384, 6, 4}, {8192, 12, 4}, {4096, 24, 4}, {32768, 6, 4}};
    for (auto& c : cfgs) {
      char what[96];
      snprintf(what, sizeof(what), "bulk chunk=%d slots=%d lanes=%d", c[0], c[1], c[2]);
      for (int rep = 0; rep < 3; ++rep) {  // the last repetition finds the slice in L2
        bulk_probe<<<ctas, 576, 200 * 1024>>>(w, c[0], c[1], c[2], dout);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", what, cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      report(what, ctas, dout, mhz);
    }
    for (int rep = 0; rep < 3; ++rep) { ldg_probe<<<ctas, 512, 200 * 1024>>>(w, dout); cudaDeviceSynchronize(); }
    report("ldg 512 threads x 8 x 16 B in flight", ctas, dout, mhz);
  }
  return 0;
}

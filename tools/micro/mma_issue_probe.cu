// Microbenchmark: sustained tcgen05.mma (kind::f16, M = 128, K = 16) rate of ONE issuing thread per SM as a function of
//   N         128 or 256 (64 or 128 tensor-pipe cycles per instruction at full rate)
//   mode 0    the product kernels' issue pattern: descriptors rebuilt per instruction from a shared-memory address
//             inside an `if (lane == 0)` region (the compiler wraps every UTCHMMA in an ELECT loop there)
//   mode 1    descriptors precomputed into registers, the loop body is only the MMAs
//   mode 2    mode 0 while 16 other warps of the CTA write 128-bit words into (other) shared memory all the time, the
//             way the fused kernels' epilogue warps write the next operand
//   distinct  operand tiles per instruction group: 1 (same 16 + N/8 KB over and over) or a 3-slot ring
// Prints tensor-pipe cycles per instruction.  Tells apart "N = 128 MMAs are shared-memory-bandwidth bound (8 KB of operand
// reads per 64 cycles)" from "one thread cannot issue them fast enough".
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../catre_b200/csrc/tc_kernels.cuh"
using namespace catre;

template <int N, int MODE>
__global__ void __launch_bounds__(64 + 512, 1) probe(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar = base, tmem_slot = base + 64, tiles = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 128);
  for (int i = threadIdx.x; i < (200 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + 1024)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // ring of 3 slots: A tile 16 KB (hi) + 16 KB (lo), B tile N*128 B (hi) + N*128 B (lo): like the product's f16x3 stage
  constexpr uint32_t A_B = 16384, B_B = N * 128, SLOT = 2 * A_B + 2 * B_B;
  constexpr int NSLOT = (N == 128) ? 3 : 2;  // 192 KB of operand tiles either way
  if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc<true>(128, N);
      const long long t0 = clock64();
      if (MODE == 1) {
        uint64_t da[NSLOT][2], db[NSLOT][2];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          da[s][0] = umma_desc_sw128(tiles + s * SLOT); da[s][1] = umma_desc_sw128(tiles + s * SLOT + A_B);
          db[s][0] = umma_desc_sw128(tiles + s * SLOT + 2 * A_B); db[s][1] = umma_desc_sw128(tiles + s * SLOT + 2 * A_B + B_B);
        }
        for (int it = 0; it < 3 * iters / NSLOT; ++it) {
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_bf16(tmem_base, da[s][0] + 2 * kk, db[s][0] + 2 * kk, idesc, 1);
              umma_bf16(tmem_base, da[s][0] + 2 * kk, db[s][1] + 2 * kk, idesc, 1);
              umma_bf16(tmem_base, da[s][1] + 2 * kk, db[s][0] + 2 * kk, idesc, 1);
            }
        }
      } else {
        int slot = 0;
        for (int it = 0; it < 3 * iters; ++it) {
          const uint32_t a_hi = tiles + slot * SLOT, a_lo = a_hi + A_B, b_hi = a_hi + 2 * A_B, b_lo = b_hi + B_B;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t off = kk * 32;
            umma_bf16(tmem_base, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, 1);
            umma_bf16(tmem_base, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1);
            umma_bf16(tmem_base, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, 1);
          }
          if (++slot == NSLOT) slot = 0;
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      *stop = 1;
    }
  } else if (warp >= 2 && MODE == 2) {
    // epilogue-like shared-memory write traffic into a region the MMAs do not read
    const uint32_t dst = tiles + NSLOT * SLOT + (uint32_t)(threadIdx.x - 64) * 16;
    uint32_t n = 0;
    while (!*stop) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_shared_v4(dst + ((n + q) & 1) * 8192, n, n, n, n);
      n += 4;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int N, int MODE>
void run(const char* what, int iters, long long* dout) {
  auto k = probe<N, MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<<<148, 64 + 512, 220 * 1024>>>(iters, dout);
    cudaError_t st = cudaDeviceSynchronize();
    if (st != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(st)); return; }
  }
  long long h[148];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; long long mx = 0;
  for (int i = 0; i < 148; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
  mean /= 148;
  const double n_mma = 36.0 * iters;
  printf("{\"probe\": \"mma_issue\", \"N\": %d, \"mode\": \"%s\", \"mma\": %.0f, \"cycles_per_mma_mean\": %.1f, \"cycles_per_mma_max\": %.1f, \"full_rate\": %d}\n",
         N, what, n_mma, mean / n_mma, mx / n_mma, N / 2);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 148 * sizeof(long long));
  const int iters = 400;
  run<128, 0>("product issue pattern", iters, dout);
  run<128, 1>("precomputed descriptors", iters, dout);
  run<128, 2>("product pattern + 16 warps writing shared memory", iters, dout);
  run<256, 0>("product issue pattern", iters, dout);
  run<256, 1>("precomputed descriptors", iters, dout);
  run<256, 2>("product pattern + 16 warps writing shared memory", iters, dout);
  return 0;
}

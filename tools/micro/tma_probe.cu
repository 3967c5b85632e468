// Microbenchmark: per-SM TMA load throughput vs bytes in flight (B200).  148 CTAs stream [128 x 64] bf16 boxes
// (16 KB, 128B swizzle) of an L2-resident matrix through a ring of `depth` slots of `boxes` boxes each; the
// consumer releases a slot as soon as it is full.  Prints bytes/clk/SM.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../catre_b200/csrc/tc_kernels.cuh"
using namespace catre;

// variant 2: `nprod` producer threads (separate warps), each with its own ring of `depth` slots of one `box_rows`-row box
__global__ void __launch_bounds__(256, 1) probe2(const __grid_constant__ CUtensorMap map, int depth, int nprod, int box_rows, int iters,
                                                 int rows_total, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_bytes = box_rows * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) { mbar_init(base + 8 * i, 1); mbar_init(base + 256 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp < nprod && lane == 0) {
    const uint32_t bar_full = base + warp * 64, bar_empty = base + 256 + warp * 64;
    const uint32_t data = base + 1024 + warp * depth * box_bytes;
    int slot = 0; uint32_t ph = 0;
    int row = ((blockIdx.x * nprod + warp) * box_rows) % rows_total;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_empty + 8 * slot, ph ^ 1);
      mbar_expect_tx(bar_full + 8 * slot, box_bytes);
      tma_load_2d(data + slot * box_bytes, &map, 0, row, bar_full + 8 * slot);
      row = (row + box_rows) % rows_total;
      if (++slot == depth) { slot = 0; ph ^= 1; }
    }
  } else if (warp >= 4 && warp < 4 + nprod && lane == 0) {
    const int w = warp - 4;
    const uint32_t bar_full = base + w * 64, bar_empty = base + 256 + w * 64;
    int slot = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_full + 8 * slot, ph);
      mbar_arrive(bar_empty + 8 * slot);
      if (++slot == depth) { slot = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// variant 3: the producers are LANES of warp 0 (lane l < nprod), consumers lanes of warp 1
__global__ void __launch_bounds__(64, 1) probe3(const __grid_constant__ CUtensorMap map, int depth, int nprod, int iters, int rows_total,
                                                long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) { mbar_init(base + 8 * i, 1); mbar_init(base + 256 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (lane < nprod) {
    const uint32_t bar_full = base + lane * 64, bar_empty = base + 256 + lane * 64;
    const uint32_t data = base + 1024 + lane * depth * 16384;
    int slot = 0; uint32_t ph = 0;
    int row = ((blockIdx.x * nprod + lane) * 128) % rows_total;
    for (int it = 0; it < iters; ++it) {
      if (warp == 0) {
        mbar_wait(bar_empty + 8 * slot, ph ^ 1);
        mbar_expect_tx(bar_full + 8 * slot, 16384);
        tma_load_2d(data + slot * 16384, &map, 0, row, bar_full + 8 * slot);
        row = (row + 128) % rows_total;
      } else {
        mbar_wait(bar_full + 8 * slot, ph);
        mbar_arrive(bar_empty + 8 * slot);
      }
      if (++slot == depth) { slot = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap map, int depth, int boxes, int iters, int rows_total,
                                               long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar_full = base, bar_empty = base + 64, data = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int slot = 0; uint32_t ph = 0;
    int row = (blockIdx.x * 128) % rows_total;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_empty + 8 * slot, ph ^ 1);
      mbar_expect_tx(bar_full + 8 * slot, boxes * 16384);
      for (int b = 0; b < boxes; ++b) {
        tma_load_2d(data + (slot * boxes + b) * 16384, &map, 0, row, bar_full + 8 * slot);
        row = (row + 128) % rows_total;
      }
      if (++slot == depth) { slot = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int slot = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_full + 8 * slot, ph);
      mbar_arrive(bar_empty + 8 * slot);
      if (++slot == depth) { slot = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int rows = 4096;  // 4096 x 64 bf16 = 512 KB matrix, L2 resident, shared by all CTAs
  __nv_bfloat16* d; cudaMalloc(&d, (size_t)rows * 64 * 2); cudaMemset(d, 0, (size_t)rows * 64 * 2);
  CUtensorMap map;
  if (!tc_make_map(&map, d, rows, 64, 64, 128)) { printf("map failed\n"); return 1; }
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
  const int iters = 2000;
  {
    CUtensorMap map256;
    tc_make_map(&map256, d, rows, 64, 64, 256);
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
    for (int box_rows : {128, 256})
      for (int nprod : {1, 2, 4})
        for (int depth : {2, 3}) {
          if (nprod * depth * box_rows * 128 > 192 * 1024) continue;
          probe2<<<148, 256, nprod * depth * box_rows * 128 + 1024>>>(box_rows == 128 ? map : map256, depth, nprod, box_rows, iters, rows, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("probe2 box_rows %3d producers %d depth %d : %6.1f B/clk/SM\n", box_rows, nprod, depth,
                 (double)iters * nprod * box_rows * 128 / (double)mx);
        }
  }
  cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
  for (int nprod : {1, 2, 4})
    for (int depth : {2, 3}) {
      probe3<<<148, 64, nprod * depth * 16384 + 1024>>>(map, depth, nprod, iters, rows, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("probe3 (lanes of one warp) producers %d depth %d : %6.1f B/clk/SM\n", nprod, depth, (double)iters * nprod * 16384 / (double)mx);
    }
  for (int boxes = 1; boxes <= 2; ++boxes)
    for (int depth = 1; depth <= 8 / boxes && depth * boxes * 16 <= 192; depth += 2) {
      for (int grid : {1, 148}) {
        probe<<<grid, 64, depth * boxes * 16384 + 1024>>>(map, depth, boxes, iters, rows, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        double bpc = (double)iters * boxes * 16384 / (double)mx;
        printf("grid %3d boxes/slot %d depth %2d in-flight %3d KB : %6.1f B/clk/SM  (%.0f cycles per slot)\n", grid, boxes, depth,
               depth * boxes * 16, bpc, (double)mx / iters);
      }
    }
  return 0;
}

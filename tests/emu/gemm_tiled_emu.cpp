// CPU emulation of tk_gemm_tiled (TEST INFRASTRUCTURE ONLY): compiles catre_b200/csrc/train_gemm_tiled.cuh -- the CUDA
// kernel source itself -- with a minimal shim (one OS thread per CUDA thread of a block, a pthread barrier for
// __syncthreads, `static` for __shared__; blocks run one after another) and runs it next to KGemmNaive.
// Build: g++ -O1 -std=c++17 -shared -fPIC -pthread -DCATRE_HOST_EMU gemm_tiled_emu.cpp
#include <pthread.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../catre_b200/csrc/train_kernels.cuh"

namespace {
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
thread_local Dim3 threadIdx, blockIdx;
pthread_barrier_t g_barrier;
inline void __syncthreads() { pthread_barrier_wait(&g_barrier); }
using std::min;
}  // namespace
#define __global__
#define __shared__ static
#define __launch_bounds__(n)
#define __align__(n) __attribute__((aligned(n)))
#include "../../catre_b200/csrc/train_gemm_tiled.cuh"

using namespace catre_train;

extern "C" void emu_gemm(const GemmP* p, int bz, int tiled) {
  // tiled: 0 = KGemmNaive, 1 = tk_gemm_tiled, 2 = tk_gemm_tiled2 (BN chosen like CudaTrainOps::gemm does)
  const int bn2 = p->N <= 64 ? 64 : 128;
  const unsigned gx = tiled == 2 ? (p->M + 127) / 128 : tiled ? (p->M + 63) / 64 : (p->M + 3) / 4;
  const unsigned gy = tiled == 2 ? (p->N + bn2 - 1) / bn2 : (p->N + 63) / 64;
  if (!tiled) {
    KGemmNaive k{*p};
    for (unsigned z = 0; z < (unsigned)bz; ++z)
      for (unsigned y = 0; y < gy; ++y)
        for (unsigned x = 0; x < gx; ++x)
          for (unsigned t = 0; t < 256; ++t) k(Idx{(int)x, (int)y, (int)z, (int)t, 256});
    return;
  }
  pthread_barrier_init(&g_barrier, nullptr, 256);
  for (unsigned z = 0; z < (unsigned)bz; ++z)
    for (unsigned y = 0; y < gy; ++y)
      for (unsigned x = 0; x < gx; ++x) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < 256; ++t)
          th.emplace_back([=]() {
            threadIdx.x = t; blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
            if (tiled == 2) { if (bn2 == 64) tk_gemm_tiled2<64>(*p); else tk_gemm_tiled2<128>(*p); }
            else tk_gemm_tiled(*p);
          });
        for (auto& h : th) h.join();
      }
  pthread_barrier_destroy(&g_barrier);
}
extern "C" int emu_gemm_param_bytes() { return (int)sizeof(GemmP); }

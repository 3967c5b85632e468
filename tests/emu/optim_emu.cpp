// CPU emulation of the fused Ranger step (TEST INFRASTRUCTURE ONLY): catre_b200/csrc/optim_kernels.cuh compiled with
// CATRE_HOST_EMU, kernels run as loops.  Same argument list as the C ABI's catre_ranger_step, with host pointers, so
// tests can also drive catre_b200.optim.FusedRanger through it.  Build: g++ -O2 -shared -fPIC -DCATRE_HOST_EMU optim_emu.cpp
#include <stdint.h>

#include "../../catre_b200/csrc/optim_kernels.cuh"

using namespace catre_train;

template <class KF>
static void run(const KF& k, long long n_threads) {
  const int nt = 256;
  for (long long b = 0; b < (n_threads + nt - 1) / nt; ++b)
    for (int t = 0; t < nt; ++t) k(Idx{(int)b, 0, 0, t, nt});
}

extern "C" int catre_ranger_step(const int64_t* table, const float* lr_wd, const int64_t* elem_start, const int64_t* row_start,
                                 int32_t n_tensors, int64_t total_elems, int64_t total_rows, float* rowmean, const RangerArgs* a,
                                 void* stream) {
  (void)stream;
  static_assert(sizeof(RangerTensor) == 64, "table row = 8 x int64");
  const RangerTensor* T = reinterpret_cast<const RangerTensor*>(table);
  if (total_rows > 0)
    run(KRangerRowMean{T, reinterpret_cast<const long long*>(row_start), n_tensors, total_rows, rowmean, a->nan_to_num}, total_rows);
  run(KRangerUpdate{T, reinterpret_cast<const long long*>(elem_start), reinterpret_cast<const long long*>(row_start), lr_wd, rowmean,
                    n_tensors, total_elems, *a}, total_elems);
  return 0;
}

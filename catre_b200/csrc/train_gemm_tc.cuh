// Tensor-core GEMM of the training step (sm_100a): the same strided, batched, split-K interface as tk_gemm_tiled (GemmP,
// train_kernels.cuh), computed with tcgen05.mma kind::f16 on 16-bit hi/lo split operands -- three products
// hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM, like the inference chain (tc_kernels.cuh).
//
// The training chain's operands are fp32 tensors with arbitrary strides (activations [rows, C], weights [C, K] or their
// transposes, per-set matrices), so there is no pre-split copy and no TMA: every thread of the CTA loads 16-byte chunks
// (8 consecutive k of one row) straight from global memory, splits them into the hi and the residual 16-bit value and writes
// both into the 128B-swizzled K-major tiles the UMMA descriptors expect (chunk' = chunk ^ (row & 7), the layout TMA's
// SWIZZLE_128B produces).  One elected thread issues the MMAs of a 64-deep k slab; their completion (tcgen05.commit) frees the
// slab's stage of a 3-stage ring, so the MMAs of slab s run underneath the conversions of slab s+1; the global loads of slabs
// s+1 and s+2 are in flight (two register buffers) while slab s is converted.  The issuing thread sits in a 17th warp of its
// own: the 16 producer warps hand a stage over through a "full" mbarrier (one arrival per warp after its cross-proxy fence) and
// never wait for the MMA issue, and there is no block-wide barrier in the k loop.
//
// Orientation: the GEMM's output COLUMNS n sit on the 128 TMEM lanes (UMMA M side) and its ROWS m on the TMEM columns (UMMA N
// side): D[n][m] = sum_k B(k, n) A(m, k).  Every output of the chain is n-fast (scn == 1), so with thread = column n a warp's
// store of one accumulator register is one contiguous 128-byte line: coalesced without a shared-memory transpose; the bias is
// one value per thread.
//
// Operand type: F16 = true for the forward GEMMs (activations and weights are O(10): fp16 hi + fp16 residual carries 22
// significand bits, what the inference chain uses), false for the backward GEMMs (bf16 hi + bf16 residual, 16 bits: gradients
// span too many decades for fp16's 5-bit exponent).
#pragma once
#include "tc_kernels.cuh"
#include "train_kernels.cuh"

namespace catre_train {

constexpr int TG_TM = 128;                           // GEMM rows per tile (TMEM columns)
constexpr int TG_TN = 128;                           // GEMM columns per tile (TMEM lanes)
constexpr int TG_BK = 64;                            // k slab: one 128-byte swizzle atom of 16-bit values
constexpr int TG_STAGES = 3;
constexpr int TG_TILE_BYTES = 128 * 128;             // 128 rows x 64 x 2 B
constexpr int TG_STAGE_BYTES = 4 * TG_TILE_BYTES;    // [n hi][n lo][m hi][m lo]
constexpr int TG_SMEM = 1024 + TG_STAGES * TG_STAGE_BYTES;
constexpr int TG_THREADS = 512;                      // producer / epilogue threads (16 warps)
constexpr int TG_CTA_THREADS = TG_THREADS + 32;      // + one warp whose elected lane issues the MMAs
constexpr int TG_CHUNKS = 128 * 8 / TG_THREADS;      // 16-byte chunks per thread and operand tile

// chunk i of thread tid inside a 128-row x 8-chunk operand tile.  k-fast operands: the 8 chunks of a row on 8 consecutive
// lanes (each lane reads 32 contiguous bytes = one sector); row-fast operands: 32 consecutive rows on a warp's lanes (every
// one of the 8 loads of a chunk is a contiguous 128-byte line).  Both are conflict-free on the swizzled store.
template <bool KFAST>
__device__ __forceinline__ void tg_chunk(int tid, int i, int& r, int& c) {
  const int q = tid + TG_THREADS * i;
  if (KFAST) { c = q & 7; r = q >> 3; } else { r = q & 127; c = q >> 7; }
}

// v[0..7] = X(row, k .. k+7), zero outside [0, rows) x [.., k1)
__device__ __forceinline__ void tg_load8(const float* base, long long s_row, long long s_k, int row, int rows, int k, int k1,
                                         bool vec, float* v) {
  if (row < rows && k + 8 <= k1) {
    const float* ptr = base + (long long)row * s_row + (long long)k * s_k;
    if (vec) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ptr)), b = __ldg(reinterpret_cast<const float4*>(ptr) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(ptr + (long long)j * s_k);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = (row < rows && k + j < k1) ? __ldg(base + (long long)row * s_row + (long long)(k + j) * s_k) : 0.0f;
  }
}

template <bool F16>
__device__ __forceinline__ void tg_store8(uint32_t tile_hi, int r, int c, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) catre::split16x2<F16>(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  const uint32_t a = tile_hi + (uint32_t)(r * 128) + (uint32_t)((c ^ (r & 7)) << 4);
  catre::st_shared_v4(a, hi[0], hi[1], hi[2], hi[3]);
  catre::st_shared_v4(a + TG_TILE_BYTES, lo[0], lo[1], lo[2], lo[3]);
}

// order-preserving map of a float onto an unsigned integer (and back): larger float <=> larger key
__device__ __forceinline__ uint32_t tg_f2key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float tg_key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// Column max + arg-max fused into the epilogue (the three 1024-channel layers whose output is only max-pooled over the points
// of a set: the [rows, 1024] pre-pool activations are never stored).  keys [sets, N] of 64 bits, zeroed before the launch:
// high word = ordered key of act(D + bias), low word = ~(row inside the set), so that an atomicMax keeps the largest value and,
// among equal values, the smallest row -- torch.max's / KColMaxArg's first-index rule.  rows_per_set must be a multiple of 128.
struct TgColMax {
  unsigned long long* keys;
  int rows_per_set;
};
// keys -> vmax [S, C], arg [S, C].  grid (ceil(S * C / nt))
struct KColMaxDecode {
  const unsigned long long* keys; float* vmax; int* arg; long long n;
  __device__ void operator()(const Idx& i) const {
    const long long e = (long long)i.bx * i.nt + i.tx;
    if (e >= n) return;
    const unsigned long long k = keys[e];
    vmax[e] = tg_key2f((uint32_t)(k >> 32));
    arg[e] = (int)(~(uint32_t)k);
  }
};

// grid (ceil(M / 128), ceil(N / 128), batch * splits), 544 threads, TG_SMEM bytes of dynamic shared memory
template <bool F16>
__global__ void __launch_bounds__(TG_CTA_THREADS, 1) tk_gemm_tc(GemmP p, TgColMax cm) {
  using namespace catre;
  extern __shared__ __align__(1024) uint8_t tg_smem[];
  const uint32_t smem_base = smem_u32(tg_smem);
  if (smem_base & 1023u) __trap();  // the swizzled tiles need a 1 KB aligned base
  const uint32_t bar_free = smem_base, bar_full = smem_base + 32, bar_done = smem_base + 64, tmem_slot = smem_base + 128;
  const bool producer = threadIdx.x < TG_THREADS;
  const uint32_t ring = smem_base + 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TG_TM, n0 = blockIdx.y * TG_TN;
  int z = blockIdx.z, k0 = 0, k1 = p.K;
  if (p.splits > 1) { k0 = z * p.k_per; k1 = min(k0 + p.k_per, p.K); z = 0; }
  const float* A = p.A + (long long)z * p.sab;
  const float* Bm = p.B + (long long)z * p.sbb;
  const int nslab = k1 > k0 ? (k1 - k0 + TG_BK - 1) / TG_BK : 0;  // a trailing split can be empty: its partial is zero

  // operand geometry: "m" tile = rows m0.. of A (element (m, k) at A[m sam + k sak]); "n" tile = columns n0.. of B
  // (element (n, k) at B[k sbk + n sbn])
  const bool a_kfast = p.sak == 1, b_kfast = p.sbk == 1;
  const bool a_vec = a_kfast && (p.sam % 4 == 0) && (p.sab % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && (k0 % 4 == 0);
  const bool b_vec = b_kfast && (p.sbn % 4 == 0) && (p.sbb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && (k0 % 4 == 0);
  const float* Am = A + (long long)m0 * p.sam;
  const float* Bn = Bm + (long long)n0 * p.sbn;
  const int m_rows = p.M - m0, n_rows = p.N - n0;  // valid rows of the two tiles (may exceed 128)

  // bias gradient folded into a weight-gradient launch (GemmP::bias_grad_partial): A is dy^T with the channel m fast, so thread
  // tid holds k values of channel m0 + (tid & 127) in every slab (tg_chunk<false>); the 4 threads of a channel are summed in a
  // fixed order through shared memory after the k loop.  Only the first column tile does it.
  const bool do_bg = p.bias_grad_partial != nullptr && blockIdx.y == 0 && !a_kfast && p.splits > 1;
  float bg = 0.0f;
  uint32_t tmem_base = 0;  // set after the allocation below (the MMA / epilogue code only runs after it)
  // two register buffers: the loads of slabs s+1 and s+2 are in flight while slab s is converted
  float va0[TG_CHUNKS][8], vb0[TG_CHUNKS][8], va1[TG_CHUNKS][8], vb1[TG_CHUNKS][8];
  auto load_slab = [&](int s, float (&va)[TG_CHUNKS][8], float (&vb)[TG_CHUNKS][8]) {
    const int kb = k0 + s * TG_BK;
#pragma unroll
    for (int i = 0; i < TG_CHUNKS; ++i) {
      int r, c;
      if (a_kfast) tg_chunk<true>(tid, i, r, c); else tg_chunk<false>(tid, i, r, c);
      tg_load8(Am, p.sam, p.sak, r, m_rows, kb + c * 8, k1, a_vec, va[i]);
      if (b_kfast) tg_chunk<true>(tid, i, r, c); else tg_chunk<false>(tid, i, r, c);
      tg_load8(Bn, p.sbn, p.sbk, r, n_rows, kb + c * 8, k1, b_vec, vb[i]);
    }
  };
  auto store_slab = [&](uint32_t stage_base, const float (&va)[TG_CHUNKS][8], const float (&vb)[TG_CHUNKS][8]) {
#pragma unroll
    for (int i = 0; i < TG_CHUNKS; ++i) {
      int r, c;
      if (b_kfast) tg_chunk<true>(tid, i, r, c); else tg_chunk<false>(tid, i, r, c);
      tg_store8<F16>(stage_base, r, c, vb[i]);
      if (a_kfast) tg_chunk<true>(tid, i, r, c); else tg_chunk<false>(tid, i, r, c);
      tg_store8<F16>(stage_base + 2 * TG_TILE_BYTES, r, c, va[i]);
    }
  };
  constexpr uint32_t idesc = umma_idesc<F16>(TG_TN, TG_TM);
  // producer side of one slab: wait until the stage's previous occupant (slab s - STAGES) has been read by its MMAs, convert and
  // store, make the generic-proxy writes visible to the tensor core's async-proxy reads, hand the stage over (one arrival per
  // warp), then start the loads of slab s + 2 into the registers just emptied
  auto step = [&](int s, float (&va)[TG_CHUNKS][8], float (&vb)[TG_CHUNKS][8]) {
    const int stage = s % TG_STAGES;
    const uint32_t sb = ring + (uint32_t)stage * TG_STAGE_BYTES;
    if (s >= TG_STAGES) mbar_wait(bar_free + 8 * stage, (uint32_t)((s / TG_STAGES) - 1) & 1);
    if (do_bg) {  // this thread's 2 x 8 k values of its A row (channel m0 + (tid & 127)): the bias gradient's share
#pragma unroll
      for (int i = 0; i < TG_CHUNKS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) bg += va[i][j];
    }
    store_slab(sb, va, vb);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_full + 8 * stage);
    if (s + 2 < nslab) load_slab(s + 2, va, vb);
  };
  // the global loads of the first two slabs are issued BEFORE the barrier / tensor-memory set-up: they depend on neither
  if (producer) {
    if (nslab > 0) load_slab(0, va0, vb0);
    if (nslab > 1) load_slab(1, va1, vb1);
  }
  if (tid == 0) {
    for (int i = 0; i < TG_STAGES; ++i) { mbar_init(bar_free + 8 * i, 1); mbar_init(bar_full + 8 * i, TG_THREADS / 32); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_base = *reinterpret_cast<volatile uint32_t*>(tg_smem + 128);
  if (producer) {
    for (int s = 0; s < nslab; s += 2) {
      step(s, va0, vb0);
      if (s + 1 < nslab) step(s + 1, va1, vb1);
    }
  } else if (lane == 0) {  // MMA issue: one elected thread
    for (int s = 0; s < nslab; ++s) {
      const int stage = s % TG_STAGES;
      const uint32_t sb = ring + (uint32_t)stage * TG_STAGE_BYTES;
      mbar_wait(bar_full + 8 * stage, (uint32_t)(s / TG_STAGES) & 1);
      tc_fence_after();
      const uint32_t n_hi = sb, n_lo = sb + TG_TILE_BYTES, m_hi = sb + 2 * TG_TILE_BYTES, m_lo = sb + 3 * TG_TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < TG_BK / 16; ++kk) {
        const uint32_t off = kk * 32;  // 16 values = 32 bytes along k inside the swizzle atom
        umma_bf16(tmem_base, umma_desc_sw128(n_hi + off), umma_desc_sw128(m_hi + off), idesc, (s | kk) != 0);
        umma_bf16(tmem_base, umma_desc_sw128(n_hi + off), umma_desc_sw128(m_lo + off), idesc, 1);
        umma_bf16(tmem_base, umma_desc_sw128(n_lo + off), umma_desc_sw128(m_hi + off), idesc, 1);
      }
      umma_commit(bar_free + 8 * stage);
      if (s == nslab - 1) umma_commit(bar_done);
    }
  }

  // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (column n = n0 + that lane), columns 32 (w / 4) .. +31 (rows m)
  if (producer && nslab > 0) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
  }
  __syncwarp();
  const int quad = warp & 3, part = warp >> 2;
  const int n = n0 + quad * 32 + lane;
  const bool n_ok = n < p.N;
  const float bias = (p.bias && n_ok && p.splits == 1) ? p.bias[n + (long long)z * p.sbias_b] : 0.0f;
  const int mc = m0 + part * 32;
  if (producer && mc < p.M) {  // warp-uniform
    float v[32];
    if (nslab > 0) {
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * 32), v);
      tmem_ld_wait32(v);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.0f;
    }
    if (n_ok) {
      if (cm.keys) {
        float best = -INFINITY; int bj = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = v[j] + bias;
          if (p.relu) x = fmaxf(x, 0.0f);
          if (mc + j < p.M && (x > best || j == 0)) { best = x; bj = j; }
        }
        const int set = mc / cm.rows_per_set, row = mc + bj - set * cm.rows_per_set;
        atomicMax(cm.keys + (long long)set * p.N + n, ((unsigned long long)tg_f2key(best) << 32) | (uint32_t)(~(uint32_t)row));
      } else if (p.splits > 1) {
        float* dst = p.partial + ((size_t)blockIdx.z * p.M + mc) * p.N + n;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (mc + j < p.M) dst[(size_t)j * p.N] = v[j];
      } else {
        float* dst = p.C + (long long)z * p.scb + (long long)mc * p.scm + (long long)n * p.scn;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += bias;
          if (p.relu) v[j] = fmaxf(v[j], 0.0f);
        }
        if (p.accumulate) {  // all loads first: a load-add-store chain per row would serialise 32 memory round trips
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = (mc + j < p.M) ? dst[(long long)j * p.scm] : 0.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += old[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (mc + j < p.M) dst[(long long)j * p.scm] = v[j];
      }
    }
  }
  if (producer && do_bg) {  // the ring is free: every producer has passed bar_done (or the split is empty and nothing used it)
    float* red = reinterpret_cast<float*>(tg_smem + 1024);
    red[tid] = bg;
    named_bar_sync(1, TG_THREADS);
    if (tid < 128 && m0 + tid < p.M)
      p.bias_grad_partial[(size_t)blockIdx.z * p.M + m0 + tid] = ((red[tid] + red[tid + 128]) + red[tid + 256]) + red[tid + 384];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

template <bool F16>
inline cudaError_t tk_gemm_tc_launch(const GemmP& p, int bz, cudaStream_t s, TgColMax cm = TgColMax{nullptr, 1}) {
  static catre::DeviceOnce configured;  // function attributes belong to the device: set them once per device, not once per process
  if (configured.needed()) {
    cudaError_t st = cudaFuncSetAttribute(tk_gemm_tc<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM);
    if (st != cudaSuccess) return st;
    configured.done();
  }
  tk_gemm_tc<F16><<<dim3((unsigned)((p.M + TG_TM - 1) / TG_TM), (unsigned)((p.N + TG_TN - 1) / TG_TN), (unsigned)bz), TG_CTA_THREADS, TG_SMEM, s>>>(p, cm);
  return cudaPeekAtLastError();
}

}  // namespace catre_train

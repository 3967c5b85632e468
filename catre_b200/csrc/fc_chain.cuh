// Fused small-M fully-connected chains of the T-Nets and the heads' global-feature layers (sm_100a, fp32 FMA).
//
//   stn   : max-pool keys [S,1024] -> 512 relu -> 256 relu -> 9 (+I3)      (pointnets/pointnet.py:29-40)
//   fstn  : max-pool keys [S,1024] -> 512 relu -> 256 relu -> 4096 (+I64)  (pointnets/pointnet.py:66-77)
//   cset  : g_set [S,1024] -> 512   = W0[:, :1024] . g_set + b0            (rot layer-0 split)
//   ts0   : g_obs [B,1024] -> 256   = ts_head.linears.0 over the global feature (heads/fc_trans_size_head.py:65)
//
// One launch runs a whole chain: a thread-block CLUSTER of 16 CTAs (8 where the device cannot co-schedule 16) owns a
// block of 16 rows (sets) for ALL layers.  Per layer every CTA computes C/16 of the output columns over the FULL K for
// the 16 rows, then broadcasts its activated outputs into the next layer's input buffer of all CTAs of the cluster over
// distributed shared memory
// (st.shared::cluster) -- the [S,512] / [S,256] intermediates never touch global memory.  Weights are packed at
// catre_pack() as [ranks][K][C/ranks] fp32, so a CTA's slice of a layer is ONE contiguous stream that the TMA engine
// copies in 32 KB chunks (cp.async.bulk issued by a producer warp into a ring of 3 slots with full / empty mbarriers, one
// "empty" arrival per compute warp: no block-wide barrier in the K loop); the prefetch runs across layer boundaries
// because weights do not depend on data.  The layer exchange is synchronised with mbarriers that remote warps arrive on
// (release.cluster) after their DSMEM stores.  Inputs are stored transposed in shared memory ([k][16 rows]) so the 8 rows a thread
// owns are two broadcast LDS.128 per k; a thread's register tile is 8 rows x 2 columns; K is additionally
// split over the threads of a CTA and summed in FIXED order through shared memory: every output's arithmetic depends
// only on its own row, never on how many rows the launch holds or on which CTA runs it -- results are bit-identical
// for any batch size and there are no atomics.
//
// This replaces round 1's 7 launches per iteration (split-K cluster FC kernels below 128 objects, tcgen05 FC
// instances above) by 3: stn chain, fstn chain, cset + ts0.  fp32 at every batch size also removes the f16x3 FC
// chain's contribution to the K = 8 tail.
#pragma once
#include "simt_kernels.cuh"

namespace catre {

constexpr int FCC_MAX_RANKS = 16;   // CTAs per cluster = column slices per layer: 16 (non-portable size) when the device can
                                    // co-schedule such clusters, else 8; chosen once per engine (fcc_pick_ranks)
constexpr int FCC_ROWS = 16;        // rows (sets) per cluster
constexpr int FCC_THREADS = 512;    // 16 compute warps: 4 per scheduler hide the shared-memory latency of the FMA loop
constexpr int FCC_CTA_THREADS = FCC_THREADS + 32;  // + one producer warp (weight stream)
constexpr int FCC_STAGES = 3;
constexpr int FCC_CHUNK_BYTES = 32768;
constexpr int FCC_PSTRIDE = 17;     // row stride of the partial-sum buffer (conflict-free both ways)
constexpr int FCC_MAX_LAYERS = 3;
// shared memory (byte offsets): A0 [1024][16] | A1 [512][16] | A2 [256][16] | ring 3 x 32 KB | barriers.  The per-slice
// partial sums of a layer's epilogue (512 columns x 17 floats = 34 KB) alias A0: layer 0 has finished reading it by then
// and the later layers never touch it.
constexpr int FCC_A0 = 1024 * FCC_ROWS * 4, FCC_A1 = 512 * FCC_ROWS * 4, FCC_A2 = 256 * FCC_ROWS * 4;
constexpr int FCC_RING_OFF = FCC_A0 + FCC_A1 + FCC_A2;
constexpr int FCC_BAR_OFF = FCC_RING_OFF + FCC_STAGES * FCC_CHUNK_BYTES;
constexpr int FCC_SMEM = FCC_BAR_OFF + 128;
static_assert(512 * FCC_PSTRIDE * 4 <= FCC_A0, "the partial sums must fit the A0 region they alias");

struct FccLayer {
  const float* wp;    // packed [ranks][K][NC] fp32 (columns >= C are zero)
  const float* bias;  // [C]
  int K, C, NC;       // NC = columns per CTA, ranks * NC >= C; an even divisor-friendly width (fcc_layer_ok)
  int relu;
};
struct FccProblem {
  const int* keys;    // ordered-int max-pool keys; row r starts at keys + r * lda
  long long lda;
  int rows;
  int n_layers;
  FccLayer L[FCC_MAX_LAYERS];
  float* out32;       // last layer, fp32 [rows][C] (or null)
  unsigned short* out_hi;  // last layer as 16-bit hi / residual pair [rows][C] (or null): tensor-core operand
  unsigned short* out_lo;
  int out_f16;
};
struct FccBatch {
  FccProblem p[2];
  int clusters0;  // clusters [0, clusters0) work on p[0], the rest on p[1]
};

// ---- explicit shared-state-space accesses (32-bit shared addresses): the buffers are reached through computed
//      offsets, and generic loads here cost a long-scoreboard round trip per k (measured: profiles/r02_ncu_fc_chain.txt)
__device__ __forceinline__ uint32_t fcc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 fcc_lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 fcc_lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float fcc_lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void fcc_sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void fcc_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fcc_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();  // a pipeline bug must fail the launch, not hang the GPU
  }
}
// cluster-scope acquire wait on a local mbarrier that remote CTAs arrive on (release.cluster) after their DSMEM stores
__device__ __forceinline__ void fcc_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();
  }
}
// one cluster-scope release fence, then relaxed remote arrivals: a release per arrival costs a MEMBAR each (8 per warp and
// exchange; 15 % of the kernel's stall samples in profiles/r02_ncu_fc_chain.txt)
__device__ __forceinline__ void fcc_fence_release_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void fcc_arrive_remote(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ uint32_t fcc_mapa(uint32_t local_addr, int rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ void fcc_st_cluster(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

// k-chunk geometry of a layer: chunk = KC k's x NC columns x 4 B <= 16 KB
__host__ __device__ inline int fcc_kc(int K, int NC) { int kc = FCC_CHUNK_BYTES / (NC * 4); return kc < K ? kc : K; }
__host__ __device__ inline int fcc_chunks(int K, int NC) { return K / fcc_kc(K, NC); }

// thread 0: start the bulk copy of global chunk g (layer, chunk-of-layer) into ring slot g % STAGES
__device__ __forceinline__ void fcc_issue(const FccProblem& P, int rank, int g, uint32_t ring, uint32_t bar_full) {
  int l = 0, ch = g;
  for (; l < FCC_MAX_LAYERS - 1; ++l) {
    const int n = fcc_chunks(P.L[l].K, P.L[l].NC);
    if (ch < n) break;
    ch -= n;
  }
  const FccLayer& Ly = P.L[l];
  const int kc = fcc_kc(Ly.K, Ly.NC);
  const float* src = Ly.wp + ((size_t)rank * Ly.K + (size_t)ch * kc) * Ly.NC;
  const int slot = g % FCC_STAGES;
  fcc_bulk_load(ring + slot * FCC_CHUNK_BYTES, src, (uint32_t)(kc * Ly.NC * 4), bar_full + 8 * slot);
}

__device__ __forceinline__ void fcc_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(FCC_THREADS) : "memory"); }

template <int RANKS>
__global__ void __launch_bounds__(FCC_CTA_THREADS, 1) fc_chain_kernel(const __grid_constant__ FccBatch bp) {
  extern __shared__ __align__(128) uint8_t fcc_smem[];
  const uint32_t sbase = fcc_smem_u32(fcc_smem);
  const uint32_t ring = sbase + FCC_RING_OFF, part = sbase, bar_full = sbase + FCC_BAR_OFF, bar_empty = bar_full + 8 * FCC_STAGES;
  const uint32_t bar_xchg = bar_empty + 8 * FCC_STAGES;  // 2 single-use barriers: the next layer's input is complete
  const int tid = threadIdx.x;
  int cl = (int)(blockIdx.x / RANKS);  // a cluster spans RANKS consecutive blocks of grid.x
  const bool second = cl >= bp.clusters0;
  const FccProblem& P = second ? bp.p[1] : bp.p[0];
  if (second) cl -= bp.clusters0;
  const int rank = (int)(blockIdx.x % RANKS), row0 = cl * FCC_ROWS;
  const int n_layers = P.n_layers;

  if (tid == 0) {
    for (int i = 0; i < FCC_STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full + 8 * i) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_empty + 8 * i), "n"(FCC_THREADS / 32) : "memory");
    }
    for (int i = 0; i < 2; ++i)  // one arrival per compute warp of every CTA of the cluster
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_xchg + 8 * i), "n"(RANKS * (FCC_THREADS / 32)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // every CTA of the cluster must be running, with its barriers initialised, before a peer writes into its shared
  // memory or arrives on its barriers.  This is the only hardware cluster barrier: the layer exchanges below use
  // mbarriers, so the producer warp (which may be blocked on a ring slot) never has to take part in them.
  if (n_layers > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  int total_chunks = 0;
  for (int l = 0; l < n_layers; ++l) total_chunks += fcc_chunks(P.L[l].K, P.L[l].NC);

  if (tid >= FCC_THREADS) {
    // ===================== producer warp: the weight stream of all layers, independent of upstream kernels ============
    if (tid == FCC_THREADS) {
      for (int g = 0; g < total_chunks; ++g) {
        const int slot = g % FCC_STAGES;
        if (g >= FCC_STAGES) fcc_wait(bar_empty + 8 * slot, (uint32_t)(((g / FCC_STAGES) - 1) & 1));
        fcc_issue(P, rank, g, ring, bar_full);
      }
    }
    return;
  }

  pdl_wait();  // the max-pool keys come from the previous kernel

  // ---- layer-0 input: keys -> floats, transposed to [k][16 rows]; thread = (row, 4-k group), rows beyond the batch are 0.
  //      All loads of a thread are issued before the first store (one L2 round trip, not K0 / 128 of them).
  {
    const int K0 = P.L[0].K;  // multiple of 128, at most 1024
    const int r = tid & 15, kq = tid >> 4;  // 16 rows x 32 quads per pass = 128 k
    const bool live = row0 + r < P.rows;
    const int* src = P.keys + (long long)(row0 + r) * P.lda;
    int4 kv[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      kv[it] = make_int4(0, 0, 0, 0);  // key of +0.0f
      if (live && it * 128 < K0) kv[it] = *reinterpret_cast<const int4*>(src + it * 128 + kq * 4);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (it * 128 < K0) {
        const uint32_t a = sbase + (uint32_t)((((it * 128 + kq * 4) * FCC_ROWS) + r) * 4);
        fcc_sts32(a, key2f(kv[it].x)); fcc_sts32(a + FCC_ROWS * 4, key2f(kv[it].y));
        fcc_sts32(a + 2 * FCC_ROWS * 4, key2f(kv[it].z)); fcc_sts32(a + 3 * FCC_ROWS * 4, key2f(kv[it].w));
      }
    }
  }
  fcc_bar_consumers();

  int g = 0;  // global chunk counter (slot = g % STAGES, parity = (g / STAGES) & 1)
  uint32_t a_in = sbase;  // this layer's input [K][16]
  for (int l = 0; l < n_layers; ++l) {
    const FccLayer& Ly = P.L[l];
    const bool last = (l + 1 == n_layers);
    const int NC = Ly.NC, C = Ly.C;
    const int CG = NC >> 1;                          // column pairs
    const int KS = (FCC_THREADS / 2) / CG;           // k slices inside the CTA; KS * NC == 512
    const int cg_i = tid % CG, rh = (tid / CG) & 1, ks = tid / (2 * CG);
    const int col0 = cg_i * 2;
    const int kc = fcc_kc(Ly.K, NC), per = kc / KS;  // k's per slice per chunk (fcc_layer_ok: kc % KS == 0)
    const int n_ch = Ly.K / kc;
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.f;
    for (int ch = 0; ch < n_ch; ++ch, ++g) {
      const int slot = g % FCC_STAGES;
      fcc_wait(bar_full + 8 * slot, (uint32_t)((g / FCC_STAGES) & 1));
      uint32_t a_addr = a_in + (uint32_t)(((ch * kc + ks * per) * FCC_ROWS + rh * 8) * 4);
      uint32_t w_addr = ring + (uint32_t)(slot * FCC_CHUNK_BYTES + ((ks * per) * NC + col0) * 4);
#pragma unroll 4
      for (int k = 0; k < per; ++k) {
        const float4 a0 = fcc_lds128(a_addr), a1 = fcc_lds128(a_addr + 16);
        const float2 w = fcc_lds64(w_addr);
        a_addr += FCC_ROWS * 4;
        w_addr += (uint32_t)NC * 4;
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(a[i], w.x, acc[i][0]);
          acc[i][1] = fmaf(a[i], w.y, acc[i][1]);
        }
      }
      __syncwarp();  // this warp is done with the slot: one arrival per warp frees it for the producer (no block-wide barrier)
      if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_empty + 8 * slot) : "memory");
    }
    fcc_bar_consumers();  // everyone has finished reading this layer's input (the partials below alias A0)
    // ---- per-slice partial sums: part[(ks * NC + col) * 17 + row]
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) fcc_sts32(part + (uint32_t)((((ks * NC + col0 + j) * FCC_PSTRIDE) + rh * 8 + i) * 4), acc[i][j]);
    fcc_bar_consumers();
    // ---- layer epilogue: fixed-order sum of the KS k-slices, bias, activation
    const int n_out = FCC_ROWS * NC;
    if (!last) {
      // row-fastest: a warp writes 2 columns x 16 rows = two 64-byte runs of the next layer's [k][16] input, to every CTA
      const uint32_t nxt = sbase + (uint32_t)(l == 0 ? FCC_A0 : FCC_A0 + FCC_A1);
      uint32_t remote[RANKS];
#pragma unroll
      for (int d = 0; d < RANKS; ++d) remote[d] = fcc_mapa(nxt, d);
      for (int idx = tid; idx < n_out; idx += FCC_THREADS) {
        const int r = idx & 15, col = idx >> 4, cglob = rank * NC + col;
        if (cglob >= C) continue;
        float s = fcc_lds32(part + (uint32_t)((col * FCC_PSTRIDE + r) * 4));
        for (int z = 1; z < KS; ++z) s += fcc_lds32(part + (uint32_t)((((z * NC + col) * FCC_PSTRIDE) + r) * 4));
        s += __ldg(Ly.bias + cglob);
        if (Ly.relu) s = fmaxf(s, 0.f);
        const uint32_t off = (uint32_t)((cglob * FCC_ROWS + r) * 4);
#pragma unroll
        for (int d = 0; d < RANKS; ++d) fcc_st_cluster(remote[d] + off, s);
      }
      // release this warp's stores to every CTA (one remote arrival per warp and destination), then acquire
      // everyone's: the next layer's input is complete in this CTA once all RANKS x 16 warps have arrived
      __syncwarp();
      if ((tid & 31) == 0) {
        fcc_fence_release_cluster();
#pragma unroll
        for (int d = 0; d < RANKS; ++d) fcc_arrive_remote(fcc_mapa(bar_xchg + 8 * l, d));
      }
      fcc_wait_cluster(bar_xchg + 8 * l, 0);
      a_in = nxt;
    } else {
      // column-fastest: coalesced global stores of the chain's result
      for (int idx = tid; idx < n_out; idx += FCC_THREADS) {
        const int col = idx % NC, r = idx / NC, cglob = rank * NC + col, grow = row0 + r;
        if (cglob >= C || grow >= P.rows) continue;
        float s = fcc_lds32(part + (uint32_t)((col * FCC_PSTRIDE + r) * 4));
        for (int z = 1; z < KS; ++z) s += fcc_lds32(part + (uint32_t)((((z * NC + col) * FCC_PSTRIDE) + r) * 4));
        s += __ldg(Ly.bias + cglob);
        if (Ly.relu) s = fmaxf(s, 0.f);
        const size_t o = (size_t)grow * C + cglob;
        if (P.out32) P.out32[o] = s;
        if (P.out_hi) split16(s, P.out_f16 != 0, P.out_hi[o], P.out_lo[o]);
      }
    }
  }
}

// =================================================================================================================
// The same layers as plain tiled GEMMs, for LARGE row counts (from 256 sets on): one launch per layer, 32 x 64 output
// tiles, every SM busy, each weight element reused for 32 rows instead of 16.  The cluster chain above wins at small
// row counts (one launch, no global intermediates); at S = 512 its 32 clusters run as two waves of 16 (one cluster needs
// 8 free SMs inside one GPC) at ~50 % FMA issue, ~4x slower than this.  BIT-IDENTICAL to the chain by construction: every
// output is accumulated in the chain kernel's own order -- k-slice z = 0 .. KS-1, inside a slice the chunks in ascending
// order and `per` consecutive k's of each chunk, fmaf by fmaf from 0, then the slices added in order, then the bias --
// so the host may pick either path per launch without changing a single bit (tests: test_result_is_independent_of_launch_size).
// =================================================================================================================
constexpr int FCT_BM = 32, FCT_BN = 64, FCT_THREADS = 128, FCT_BK = 16;

struct FctP {
  const int* keys; long long lda_keys;  // layer-0 input: ordered-int keys (or null)
  const float* a32; int lda32;          // later layers: fp32 [rows][K]
  const float* w; int ldw;              // ORIGINAL layout [C][K] row-major
  const float* bias;
  int rows, C, K;
  int KS, per, kc;                      // the chain's k-slicing of this layer (fcc_kc / KS * NC == 512)
  int relu;
  float* out32; int ldo;
  unsigned short* out_hi; unsigned short* out_lo; int out_f16;
};

__global__ void __launch_bounds__(FCT_THREADS) fc_tiled_kernel(const FctP p) {
  __shared__ __align__(16) float As[2][FCT_BK][FCT_BM + 4];
  __shared__ __align__(16) float Ws[2][FCT_BK][FCT_BN + 4];
  pdl_wait();
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // 16 column quads x 8 row quads
  const int r0 = blockIdx.y * FCT_BM, c0 = blockIdx.x * FCT_BN;
  const int n_ch = p.K / p.kc, sub = p.per / FCT_BK;          // per is a multiple of 16 here (fct_ok)
  const int steps_per_slice = n_ch * sub, n_steps = p.KS * steps_per_slice;
  // step s -> first k of its 16-wide tile, in the chain's order
  auto k_of = [&](int s) {
    const int z = s / steps_per_slice, rem = s % steps_per_slice, ch = rem / sub, q = rem % sub;
    return ch * p.kc + z * p.per + q * FCT_BK;
  };
  // global -> registers: A tile 32 rows x 16 k = 128 float4 (one per thread), W tile 64 cols x 16 k = 256 float4 (two)
  float4 ra, rw[2];
  auto load = [&](int s) {
    const int k0 = k_of(s);
    {
      const int row = tid >> 2, kq = (tid & 3) * 4;
      ra = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + row < p.rows) {
        if (p.keys) {
          const int4 kv = *reinterpret_cast<const int4*>(p.keys + (long long)(r0 + row) * p.lda_keys + k0 + kq);
          ra = make_float4(key2f(kv.x), key2f(kv.y), key2f(kv.z), key2f(kv.w));
        } else {
          ra = *reinterpret_cast<const float4*>(p.a32 + (long long)(r0 + row) * p.lda32 + k0 + kq);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int e = tid + it * FCT_THREADS, col = e >> 2, kq = (e & 3) * 4;
      rw[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + col < p.C) rw[it] = *reinterpret_cast<const float4*>(p.w + (long long)(c0 + col) * p.ldw + k0 + kq);
    }
  };
  auto stash = [&](int buf) {
    {
      const int row = tid >> 2, kq = (tid & 3) * 4;
      As[buf][kq + 0][row] = ra.x; As[buf][kq + 1][row] = ra.y; As[buf][kq + 2][row] = ra.z; As[buf][kq + 3][row] = ra.w;
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int e = tid + it * FCT_THREADS, col = e >> 2, kq = (e & 3) * 4;
      Ws[buf][kq + 0][col] = rw[it].x; Ws[buf][kq + 1][col] = rw[it].y; Ws[buf][kq + 2][col] = rw[it].z; Ws[buf][kq + 3][col] = rw[it].w;
    }
  };
  float acc[4][4], tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }
  load(0);
  stash(0);
  __syncthreads();
  for (int s = 0; s < n_steps; ++s) {
    const int buf = s & 1;
    if (s + 1 < n_steps) load(s + 1);
#pragma unroll
    for (int kk = 0; kk < FCT_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if ((s + 1) % steps_per_slice == 0) {  // end of k-slice z: the chain adds the slices in order, the first one as is
      const bool first = (s + 1 == steps_per_slice);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] = first ? acc[i][j] : __fadd_rn(tot[i][j], acc[i][j]); acc[i][j] = 0.f; }
    }
    if (s + 1 < n_steps) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int grow = r0 + ty * 4 + i;
    if (grow >= p.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = c0 + tx * 4 + j;
      if (gc >= p.C) continue;
      float y = __fadd_rn(tot[i][j], __ldg(p.bias + gc));
      if (p.relu) y = fmaxf(y, 0.f);
      if (p.out32) p.out32[(size_t)grow * p.ldo + gc] = y;
      if (p.out_hi) split16(y, p.out_f16 != 0, p.out_hi[(size_t)grow * p.C + gc], p.out_lo[(size_t)grow * p.C + gc]);
    }
  }
}

// the chain's order for a layer whose slices hold ONE k each (stn.fc3: 9 outputs, K = 256, KS = 256, per = 1):
// out = ((a0 w0 + a1 w1) + a2 w2) + ... + bias with every product rounded on its own (fmaf(a, w, 0) in the chain)
__global__ void __launch_bounds__(128) fc_small_kernel(const float* __restrict__ a32, int lda32, const float* __restrict__ w, int ldw,
                                                       const float* __restrict__ bias, int rows, int C, int K, float* __restrict__ out32) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int r = i / C, c = i % C;
  const float* a = a32 + (size_t)r * lda32;
  const float* ww = w + (size_t)c * ldw;
  float s = __fmul_rn(a[0], ww[0]);
  for (int k = 1; k < K; ++k) s = __fadd_rn(s, __fmul_rn(a[k], ww[k]));
  out32[(size_t)r * C + c] = __fadd_rn(s, bias[c]);
}

// ---- host side -----------------------------------------------------------------------------------------------
// [C][K] row-major -> [ranks][K][NC] (zero-padded columns)
inline void fcc_pack(const float* w, int C, int K, int NC, int ranks, float* out) {
  for (int r = 0; r < ranks; ++r)
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < NC; ++j) {
        const int c = r * NC + j;
        out[((size_t)r * K + k) * NC + j] = (c < C) ? w[(size_t)c * K + k] : 0.0f;
      }
}
// columns per CTA for a layer of C outputs on `ranks` CTAs (at least one column pair)
inline int fcc_nc(int C, int ranks) { int nc = (C + ranks - 1) / ranks; nc += nc & 1; return nc < 2 ? 2 : nc; }
inline bool fcc_layer_ok(const FccLayer& L, int ranks) {
  if (L.NC * ranks < L.C || L.K > 1024 || L.K % 128) return false;
  if (L.NC % 2) return false;
  const int CG = L.NC / 2;
  if (CG > FCC_THREADS / 2 || (FCC_THREADS / 2) % CG) return false;
  const int KS = (FCC_THREADS / 2) / CG, kc = fcc_kc(L.K, L.NC);
  return KS * L.NC == 512 && L.K % kc == 0 && kc % KS == 0 && (kc * L.NC * 4) % 16 == 0;
}
template <int RANKS>
inline cudaError_t fcc_launch_t(const FccBatch& b, int clusters_total, cudaStream_t s) {
  auto kern = fc_chain_kernel<RANKS>;
  static DeviceOnce configured;  // function attributes belong to the device: set them once per device, not once per process
  if (configured.needed()) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FCC_SMEM);
    if (st == cudaSuccess && RANKS > 8) st = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (st != cudaSuccess) return st;
    configured.done();
  }
  if (clusters_total < 1) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters_total * RANKS)); cfg.blockDim = dim3(FCC_CTA_THREADS);
  cfg.dynamicSmemBytes = FCC_SMEM; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = RANKS; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, b);
}
inline cudaError_t fcc_launch(const FccBatch& b, int clusters_total, int ranks, cudaStream_t s) {
  return ranks == 16 ? fcc_launch_t<16>(b, clusters_total, s) : fcc_launch_t<8>(b, clusters_total, s);
}
// 16 when the device can keep at least 4 clusters of 16 CTAs (215 KB of shared memory each) resident, else 8
inline int fcc_pick_ranks() {
  auto kern = fc_chain_kernel<16>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FCC_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 8;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16 * 8); cfg.blockDim = dim3(FCC_CTA_THREADS); cfg.dynamicSmemBytes = FCC_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 8; }
  return n >= 4 ? 16 : 8;
}

// the chain's k-slicing of a layer of NC columns per CTA (what fc_tiled_kernel must reproduce)
inline void fct_geometry(int K, int NC, int* KS, int* per, int* kc) {
  *kc = fcc_kc(K, NC);
  *KS = (FCC_THREADS / 2) / (NC / 2);
  *per = *kc / *KS;
}
inline bool fct_ok(const FctP& p) { return p.per % FCT_BK == 0 && p.K % p.kc == 0 && p.K % 4 == 0 && p.KS * p.per == p.kc; }
inline cudaError_t fct_launch(const FctP& p, cudaStream_t s) {
  if (p.rows < 1) return cudaSuccess;
  dim3 grid((unsigned)((p.C + FCT_BN - 1) / FCT_BN), (unsigned)((p.rows + FCT_BM - 1) / FCT_BM));
  return launch_pdl(fc_tiled_kernel, grid, dim3(FCT_THREADS), (size_t)0, s, p);
}

}  // namespace catre

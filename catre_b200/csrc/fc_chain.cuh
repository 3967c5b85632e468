// Fused small-M fully-connected chains of the T-Nets and the heads' global-feature layers (sm_100a, fp32 FMA).
//
//   stn   : max-pool keys [S,1024] -> 512 relu -> 256 relu -> 9 (+I3)      (pointnets/pointnet.py:29-40)
//   fstn  : max-pool keys [S,1024] -> 512 relu -> 256 relu -> 4096 (+I64)  (pointnets/pointnet.py:66-77)
//   cset  : g_set [S,1024] -> 512   = W0[:, :1024] . g_set + b0            (rot layer-0 split)
//   ts0   : g_obs [B,1024] -> 256   = ts_head.linears.0 over the global feature (heads/fc_trans_size_head.py:65)
//
// One launch runs a whole chain: a thread-block CLUSTER of 8 CTAs owns a block of 16 rows (sets) for ALL layers.
// Per layer every CTA computes C/8 of the output columns over the FULL K for the 16 rows, then broadcasts its
// activated outputs into the next layer's input buffer of all 8 CTAs over distributed shared memory
// (st.shared::cluster) -- the [S,512] / [S,256] intermediates never touch global memory.  Weights are packed at
// catre_pack() as [8 ranks][K][C/8] fp32, so a CTA's slice of a layer is ONE contiguous stream that the TMA engine
// copies in 16 KB chunks (cp.async.bulk + mbarrier ring of 4 slots); the prefetch runs across layer boundaries because
// weights do not depend on data.  Inputs are stored transposed in shared memory ([k][16 rows]) so the 8 rows a thread
// owns are two broadcast LDS.128 per k; a thread's register tile is 8 rows x 2 (or 4) columns; K is additionally
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

constexpr int FCC_RANKS = 8;        // CTAs per cluster = column slices per layer
constexpr int FCC_ROWS = 16;        // rows (sets) per cluster
constexpr int FCC_THREADS = 256;
constexpr int FCC_STAGES = 4;
constexpr int FCC_CHUNK_BYTES = 16384;
constexpr int FCC_PSTRIDE = 17;     // row stride of the partial-sum buffer (conflict-free both ways)
constexpr int FCC_MAX_LAYERS = 3;
// shared memory: A0 [1024][16] | A1 [512][16] | A2 [256][16] | ring 4 x 16 KB | partials | barriers
constexpr int FCC_A0 = 1024 * FCC_ROWS * 4, FCC_A1 = 512 * FCC_ROWS * 4, FCC_A2 = 256 * FCC_ROWS * 4;
constexpr int FCC_PART = 512 * FCC_PSTRIDE * 4 + 256;  // KS * NC <= 512 partial columns
constexpr int FCC_SMEM = FCC_A0 + FCC_A1 + FCC_A2 + FCC_STAGES * FCC_CHUNK_BYTES + FCC_PART + 64;

struct FccLayer {
  const float* wp;    // packed [8][K][NC] fp32 (columns >= C are zero)
  const float* bias;  // [C]
  int K, C, NC;       // NC = columns per CTA, 8 * NC >= C; NC in {2, 32, 64, 512}
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

__device__ __forceinline__ uint32_t fcc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// store one float into the same shared-memory offset of CTA `rank` of the cluster
__device__ __forceinline__ void fcc_st_remote(uint32_t local_addr, int rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// k-chunk geometry of a layer: chunk = KC k's x NC columns x 4 B <= 16 KB
__host__ __device__ inline int fcc_kc(int K, int NC) { int kc = FCC_CHUNK_BYTES / (NC * 4); return kc < K ? kc : K; }

template <int CT>
__device__ __forceinline__ void fcc_accumulate(const float* __restrict__ As, const float* __restrict__ Wc, int NC, int k_glob0,
                                               int k_lo, int k_hi, int rh, int col0, float (&acc)[8][CT]) {
#pragma unroll 4
  for (int k = k_lo; k < k_hi; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(As + (size_t)(k_glob0 + k) * FCC_ROWS + rh * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(As + (size_t)(k_glob0 + k) * FCC_ROWS + rh * 8 + 4);
    float w[CT];
    if (CT == 4) *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(Wc + (size_t)k * NC + col0);
    else *reinterpret_cast<float2*>(w) = *reinterpret_cast<const float2*>(Wc + (size_t)k * NC + col0);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

struct FccCtx {
  const FccProblem* P;
  int nch[FCC_MAX_LAYERS];
  int total_chunks;
  int rank;
  uint32_t ring_u32, bar0;
  const uint8_t* ring;
  float* part;
};

// thread 0: start the bulk copy of global chunk g (layer, chunk-of-layer) into slot g % STAGES
__device__ __forceinline__ void fcc_issue(const FccCtx& c, int g) {
  int l = 0, ch = g;
  while (ch >= c.nch[l]) { ch -= c.nch[l]; ++l; }
  const FccLayer& Ly = c.P->L[l];
  const int kc = fcc_kc(Ly.K, Ly.NC);
  const float* src = Ly.wp + ((size_t)c.rank * Ly.K + (size_t)ch * kc) * Ly.NC;
  const int slot = g % FCC_STAGES;
  fcc_bulk_load(c.ring_u32 + slot * FCC_CHUNK_BYTES, src, (uint32_t)(kc * Ly.NC * 4), c.bar0 + 8 * slot);
}

// the K loop of one layer for this CTA's NC columns and the cluster's 16 rows; leaves the per-slice partial sums in
// c.part[(ks * NC + col) * 17 + row]
template <int CT>
__device__ __forceinline__ void fcc_layer(const FccCtx& c, int l, const float* __restrict__ As, int& g) {
  const FccLayer& Ly = c.P->L[l];
  const int tid = threadIdx.x, NC = Ly.NC;
  const int CG = NC / CT;                          // column groups
  const int KS = FCC_THREADS / (2 * CG);           // k slices inside the CTA (1, 4, 8 or 128)
  const int cg_i = tid % CG, rh = (tid / CG) & 1, ks = tid / (2 * CG);
  const int col0 = cg_i * CT;
  const int kc = fcc_kc(Ly.K, NC), per = kc / KS;  // k's per slice per chunk (fcc_layer_ok: kc % KS == 0)
  float acc[8][CT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;
  for (int ch = 0; ch < c.nch[l]; ++ch, ++g) {
    const int slot = g % FCC_STAGES;
    fcc_wait(c.bar0 + 8 * slot, (uint32_t)((g / FCC_STAGES) & 1));
    const float* Wc = reinterpret_cast<const float*>(c.ring + slot * FCC_CHUNK_BYTES);
    fcc_accumulate<CT>(As, Wc, NC, ch * kc, ks * per, (ks + 1) * per, rh, col0, acc);
    __syncthreads();  // every thread is done with this slot
    if (tid == 0 && g + FCC_STAGES < c.total_chunks) fcc_issue(c, g + FCC_STAGES);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) c.part[((size_t)ks * NC + col0 + j) * FCC_PSTRIDE + rh * 8 + i] = acc[i][j];
}

__global__ void __cluster_dims__(FCC_RANKS, 1, 1) __launch_bounds__(FCC_THREADS, 1)
fc_chain_kernel(const __grid_constant__ FccBatch bp) {
  extern __shared__ __align__(128) uint8_t fcc_smem[];
  float* A[3] = {reinterpret_cast<float*>(fcc_smem), reinterpret_cast<float*>(fcc_smem + FCC_A0),
                 reinterpret_cast<float*>(fcc_smem + FCC_A0 + FCC_A1)};
  uint8_t* ring = fcc_smem + FCC_A0 + FCC_A1 + FCC_A2;
  const int tid = threadIdx.x;
  int cl = (int)(blockIdx.x / FCC_RANKS);  // a cluster spans 8 consecutive blocks of grid.x
  FccCtx c;
  c.P = (cl < bp.clusters0) ? &bp.p[0] : &bp.p[1];
  if (cl >= bp.clusters0) cl -= bp.clusters0;
  const FccProblem& P = *c.P;
  c.rank = (int)(blockIdx.x % FCC_RANKS);
  c.ring = ring;
  c.ring_u32 = fcc_smem_u32(ring);
  c.part = reinterpret_cast<float*>(ring + FCC_STAGES * FCC_CHUNK_BYTES);
  c.bar0 = fcc_smem_u32(ring + FCC_STAGES * FCC_CHUNK_BYTES + FCC_PART);
  const int rank = c.rank, row0 = cl * FCC_ROWS;

  if (tid == 0) {
    for (int i = 0; i < FCC_STAGES; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.bar0 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // every CTA of the cluster must be running before its shared memory is written remotely: arrive now, wait
  // right before the first exchange
  if (P.n_layers > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  // ---- weight stream: all of it is independent of upstream kernels, so it starts before the dependency wait
  c.total_chunks = 0;
  for (int l = 0; l < FCC_MAX_LAYERS; ++l) {
    c.nch[l] = (l < P.n_layers) ? P.L[l].K / fcc_kc(P.L[l].K, P.L[l].NC) : 0;
    c.total_chunks += c.nch[l];
  }
  if (tid == 0)
    for (int g = 0; g < FCC_STAGES && g < c.total_chunks; ++g) fcc_issue(c, g);

  pdl_wait();  // the max-pool keys come from the previous kernel

  // ---- layer-0 input: keys -> floats, transposed to [k][16 rows]; lane = (row, 4-k group), rows beyond the batch are 0
  {
    const int K0 = P.L[0].K;
    const int r = tid & 15, kq = tid >> 4;  // 16 rows x 16 quads per pass = 64 k
    const bool live = row0 + r < P.rows;
    const int* src = P.keys + (long long)(row0 + r) * P.lda;
    for (int k0 = 0; k0 < K0; k0 += 64) {
      const int k = k0 + kq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const int4 kv = *reinterpret_cast<const int4*>(src + k);
        v = make_float4(key2f(kv.x), key2f(kv.y), key2f(kv.z), key2f(kv.w));
      }
      A[0][(k + 0) * FCC_ROWS + r] = v.x; A[0][(k + 1) * FCC_ROWS + r] = v.y;
      A[0][(k + 2) * FCC_ROWS + r] = v.z; A[0][(k + 3) * FCC_ROWS + r] = v.w;
    }
  }
  __syncthreads();

  int g = 0;  // global chunk counter (slot = g % STAGES, parity = (g / STAGES) & 1)
  for (int l = 0; l < P.n_layers; ++l) {
    const FccLayer& Ly = P.L[l];
    const bool last = (l + 1 == P.n_layers);
    const int NC = Ly.NC;
    if (NC >= 128) fcc_layer<4>(c, l, A[l], g); else fcc_layer<2>(c, l, A[l], g);
    __syncthreads();
    // ---- layer epilogue: fixed-order sum of the KS k-slices, bias, activation
    const int KS = FCC_THREADS / (2 * (NC / ((NC >= 128) ? 4 : 2)));
    const int n_out = FCC_ROWS * NC;
    const float* part = c.part;
    if (!last) {
      if (l == 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // the start-up arrive above
      // row-fastest: a warp writes 2 columns x 16 rows = two 64-byte runs of the next layer's [k][16] input, to all 8 CTAs
      const uint32_t nxt = fcc_smem_u32(A[l + 1]);
      for (int idx = tid; idx < n_out; idx += FCC_THREADS) {
        const int r = idx & 15, col = idx >> 4, cglob = rank * NC + col;
        if (cglob >= Ly.C) continue;
        float s = part[(size_t)col * FCC_PSTRIDE + r];
        for (int z = 1; z < KS; ++z) s += part[((size_t)z * NC + col) * FCC_PSTRIDE + r];
        s += Ly.bias[cglob];
        if (Ly.relu) s = fmaxf(s, 0.f);
        const uint32_t off = nxt + (uint32_t)((cglob * FCC_ROWS + r) * 4);
#pragma unroll
        for (int d = 0; d < FCC_RANKS; ++d) fcc_st_remote(off, d, s);
      }
      // release my stores / acquire everyone's: the next layer's input is complete in every CTA of the cluster
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      // column-fastest: coalesced global stores of the chain's result
      for (int idx = tid; idx < n_out; idx += FCC_THREADS) {
        const int col = idx % NC, r = idx / NC, cglob = rank * NC + col, grow = row0 + r;
        if (cglob >= Ly.C || grow >= P.rows) continue;
        float s = part[(size_t)col * FCC_PSTRIDE + r];
        for (int z = 1; z < KS; ++z) s += part[((size_t)z * NC + col) * FCC_PSTRIDE + r];
        s += Ly.bias[cglob];
        if (Ly.relu) s = fmaxf(s, 0.f);
        const size_t o = (size_t)grow * Ly.C + cglob;
        if (P.out32) P.out32[o] = s;
        if (P.out_hi) split16(s, P.out_f16 != 0, P.out_hi[o], P.out_lo[o]);
      }
    }
    __syncthreads();  // the partial buffer is reused by the next layer
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
// [C][K] row-major -> [8][K][NC] (zero-padded columns)
inline void fcc_pack(const float* w, int C, int K, int NC, float* out) {
  for (int r = 0; r < FCC_RANKS; ++r)
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < NC; ++j) {
        const int c = r * NC + j;
        out[((size_t)r * K + k) * NC + j] = (c < C) ? w[(size_t)c * K + k] : 0.0f;
      }
}
inline bool fcc_layer_ok(const FccLayer& L) {
  if (L.NC * FCC_RANKS < L.C || L.K > 1024 || L.K % 4) return false;
  const int CT = (L.NC >= 128) ? 4 : 2;
  if (L.NC % CT) return false;
  const int CG = L.NC / CT;
  if (FCC_THREADS % (2 * CG)) return false;
  const int KS = FCC_THREADS / (2 * CG), kc = fcc_kc(L.K, L.NC);
  return KS * L.NC <= 512 && L.K % kc == 0 && kc % KS == 0 && (kc * L.NC * 4) % 16 == 0;
}
inline cudaError_t fcc_launch(const FccBatch& b, int clusters_total, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t st = cudaFuncSetAttribute(fc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FCC_SMEM);
    if (st != cudaSuccess) return st;
    configured = true;
  }
  if (clusters_total < 1) return cudaSuccess;
  return launch_pdl(fc_chain_kernel, dim3((unsigned)(clusters_total * FCC_RANKS)), dim3(FCC_THREADS), (size_t)FCC_SMEM, s, b);
}

}  // namespace catre

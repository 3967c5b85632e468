// Tensor-core path (sm_100a): tcgen05.mma kind::f16 with 16-bit hi/lo split operands ("f16x3": fp16 hi + fp16
// residual, three products hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM; single-product mode: bf16), operands staged HBM/L2 -> shared memory by
// TMA (cp.async.bulk.tensor, 128B swizzle), mbarrier producer/consumer pipeline, warp-specialised:
//   warp 0 : TMA producer          warp 1 : TMEM allocator + MMA issuer (one elected thread)
//   warps 2-5 : epilogue (tcgen05.ld -> registers -> fused bias/ReLU + column-max | GroupNorm partials |
//               bf16 hi/lo re-split for the next layer)
// One persistent CTA per SM loops over output tiles; the fp32 accumulator is double-buffered in TMEM so
// the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Every wide point-wise layer of the CATRE encoder/rot-head is a GEMM  D = W . A^T  over K = C_in with
// both operands K-major; two orientations are used:
//   CH_ON_LANES : M side (128 TMEM lanes) = output channels, N side (<=256 TMEM columns) = points.
//                 A column max over points / GroupNorm sums over points are per-thread serial reductions.
//   PT_ON_LANES : M side = points, N side = output channels.  Each thread owns one point's contiguous
//                 channels, so the re-split bf16 activations are written with 16-byte vector stores.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; cuTensorMapEncodeTiled is fetched through the runtime)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "simt_kernels.cuh"

namespace catre {

// epilogue warps: EW / 4 per TMEM lane quadrant, each takes BN / (EW / 4) of the columns.  The
// channel-on-lanes epilogues (MMA-bound layers) use 8; the point-on-lanes epilogues, which are pure
// CUDA-core work (bias/ReLU or GroupNorm+GELU, bf16 split, stores), use 16 when the tile is wide enough.
template <int ORIENT, int BN>
struct TcEpi {
  static constexpr int EW = (ORIENT == 1 && BN >= 128) ? 16 : 8;
  static constexpr int THREADS = 64 + 32 * EW;  // + TMA warp + MMA warp
};
constexpr int TC_BK = 64;  // K slab = one 128-byte swizzle atom of bf16
enum { CH_ON_LANES = 0, PT_ON_LANES = 1 };
// epilogues:            orientation   what leaves the kernel
//   EPI_MAX             CH_ON_LANES   column max over the tile's points of act(D + bias) -> atomicMax keys
//   EPI_STATS           CH_ON_LANES   GroupNorm partial sums of D + rowvec[set] only (nothing stored)
//   EPI_SPLIT           PT_ON_LANES   act(D + bias) -> bf16 hi/lo [R, C] (operand of the next layer), TMA-stored
//   EPI_SPLIT_MAX       PT_ON_LANES   D -> bf16 hi/lo, and column max over the tile's points -> atomicMax keys
//   EPI_SPLIT_STREAM    PT_ON_LANES   as EPI_SPLIT (+ optional fp32 copy) but with the weights streamed, not resident:
//                                     the small-M FC layers of the T-Nets, whose K is up to 1024
enum { EPI_MAX = 0, EPI_SPLIT = 2, EPI_STATS = 3, EPI_SPLIT_MAX = 5, EPI_SPLIT_STREAM = 6 };

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for outstanding tcgen05.ld; the 32 destination registers pass through the statement so the
// compiler cannot schedule their consumers above the wait
__device__ __forceinline__ void tmem_ld_wait32(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// visit NCHUNK consecutive 32-column chunks of this warp's TMEM lanes; the load of chunk c+1 is in
// flight while f(c, values) runs (two register buffers)
template <int NCHUNK, bool PIPE, typename F>
__device__ __forceinline__ void tmem_foreach32(uint32_t taddr, F&& f) {
  if (!PIPE) {  // many resident warps hide the load latency; keep the register footprint small
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      float v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait32(v);
      f(c, v);
    }
    return;
  }
  float va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    if (c & 1) {
      tmem_ld_wait32(vb);
      if (c + 1 < NCHUNK) tmem_ld32(taddr + (c + 1) * 32, va);
      f(c, vb);
    } else {
      tmem_ld_wait32(va);
      if (c + 1 < NCHUNK) tmem_ld32(taddr + (c + 1) * 32, vb);
      f(c, va);
    }
  }
}

// shared-memory matrix descriptor: K-major, 128B swizzle, rows 128 B apart, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout[61,64)=2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The 16-bit operand type of a tensor-core mode.  3-product (fp32-parity) mode: fp16 hi + fp16 residual, i.e.
// 22 significand bits per operand (the dropped lo*lo term is 2^-22 relative) -- 4x closer to fp32 than a
// bf16 hi/lo split (16 bits) at the same MMA cost; the path's activations and weights are O(10) at most, far
// inside fp16's range, and the conversions saturate instead of overflowing.  Single-product mode: bf16
// (BASELINE.json config 3).
template <int NPROD>
struct TcOperand { static constexpr bool F16 = (NPROD == 3); };

// instruction descriptor kind::f16: D fp32 (bit 4), A/B format at bits 7 / 10 (0 = f16, 1 = bf16), K-major both,
// N>>3 at 17, M>>4 at 24
template <bool F16>
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// two floats -> packed 16-bit hi pair and packed 16-bit residual pair (element 0 in the low half)
template <bool F16>
__device__ __forceinline__ void split16x2(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
  if (F16) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(x1), "f"(x0));
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
    const float r0 = x0 - h.x;
    const float r1 = x1 - h.y;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(r1), "f"(r0));
  } else {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi2 << 16);
    const float r1 = x1 - __uint_as_float(hi2 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(r1), "f"(r0));
  }
}
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------
// The GEMM kernel
// ------------------------------------------------------------------------------------------------
struct TcGemmP {
  int K;                 // reduction length, multiple of 64
  int m_tiles, n_tiles;  // tiles on the M side (128 rows each) / N side (BN rows each)
  int rows_per_set;      // n_obs: observed points per object (a tile never straddles two sets; see set_of_row)
  int rows_per_obj;      // P = n_obs + n_prior
  int nb_per_set;        // PT_ON_LANES: the NB operand is per set (rows set*BN .. +BN): feature transform
  int res_stages;        // resident-weight layers: activation stages in the ring (set by tc_launch)
  int out_bufs;          // point-on-lanes layers: output staging buffers (1 or 2, set by tc_launch)
  // epilogue
  const float* bias;     // per output channel (or null)
  int relu;
  int* gmax; int C;                             // EPI_MAX / EPI_SPLIT_MAX: keys [S, C]
  const float* rowvec; int ldrv;                // EPI_STATS: per-set additive vector [S, ldrv]
  float* stats; int stats_ld, stats_goff;       // EPI_STATS: GroupNorm partials [R/(BN/2), stats_ld, 2]
  float* out32; int ldo32; int rows32;          // EPI_SPLIT_STREAM: optional fp32 copy [rows32, ldo32]
  // object-group launches (the group's output lives in a small buffer that stays in L2 until the next layer has read it):
  int mi_in0;            // PT_ON_LANES: first 128-row tile of the group in the INPUT (M-side loads use mi + mi_in0; stores use mi)
  long long set_row0;    // CH_ON_LANES / EPI_MAX: global row of the group's first point (the set of N-tile ni is taken at ni * BN + set_row0)
};

template <int ORIENT, int BN, int NPROD>
struct TcCfg {
  static constexpr int ARR = (NPROD == 3) ? 2 : 1;  // hi (+ lo) arrays per operand
  static constexpr int MA_BYTES = 128 * 128;        // 128 rows x 64 bf16
  static constexpr int NB_BYTES = BN * 128;
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int BIAS_BYTES = (ORIENT == 1) ? 2048 : 4096;  // <= 512 (points on lanes) / 1024 channels
  // point-on-lanes layers stage their bf16 hi/lo output tile in shared memory for the TMA store engine:
  // BN/64 boxes of [128 rows x 64 channels] per array
  static constexpr int OUT_BYTES = (ORIENT == 1) ? (BN / 64) * 16384 * ARR : 0;
  // streaming both operands: as many stages as fit (<= 4); the output staging is double-buffered when >= 3
  // stages remain
  static constexpr int STAGE_BYTES = (MA_BYTES + NB_BYTES) * ARR;
  static constexpr int FIXED1 = 1024 + BIAS_BYTES + OUT_BYTES, FIXED2 = FIXED1 + OUT_BYTES;
  static constexpr int OUT_BUFS = (ORIENT == 1 && (MAX_SMEM - FIXED2) / STAGE_BYTES >= 3) ? 2 : 1;
  static constexpr int FIXED = (OUT_BUFS == 2) ? FIXED2 : FIXED1;  // barriers | ... | output staging | bias
  static constexpr int STAGES = ((MAX_SMEM - FIXED) / STAGE_BYTES) > 4 ? 4 : ((MAX_SMEM - FIXED) / STAGE_BYTES);
  static constexpr int SMEM_BYTES = FIXED + STAGES * STAGE_BYTES;
  // resident-weight variant (point-on-lanes layers with a fixed weight operand): all K slabs of the CTA's
  // BN weight rows stay in shared memory for the whole kernel; only the activation tiles stream.
  static constexpr int RES_STAGE_BYTES = MA_BYTES * ARR;
  static int res_bytes(int K) { return (K / 64) * NB_BYTES * ARR; }
  static int res_out_bufs(int K) { return (MAX_SMEM - FIXED2 - res_bytes(K)) / RES_STAGE_BYTES >= 2 ? 2 : 1; }
  static int res_fixed(int K) { return res_out_bufs(K) == 2 ? FIXED2 : FIXED1; }
  static int res_stages(int K) {
    int st = (MAX_SMEM - res_fixed(K) - res_bytes(K)) / RES_STAGE_BYTES;
    return st > 4 ? 4 : st;
  }
  static int res_smem(int K) { return res_fixed(K) + res_bytes(K) + res_stages(K) * RES_STAGE_BYTES; }
};
// weights stay resident for the plain point-on-lanes layers (bias/ReLU/split epilogue)
template <int ORIENT, int EPI>
struct TcRes { static constexpr bool value = (ORIENT == 1 && EPI == 2); };
template <int ORIENT, int EPI, int BN, int NPROD>
__global__ void __launch_bounds__(TcEpi<ORIENT, BN>::THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap ma_hi, const __grid_constant__ CUtensorMap ma_lo,
               const __grid_constant__ CUtensorMap nb_hi, const __grid_constant__ CUtensorMap nb_lo,
               const __grid_constant__ CUtensorMap out_hi, const __grid_constant__ CUtensorMap out_lo, const TcGemmP p) {
  using Cfg = TcCfg<ORIENT, BN, NPROD>;
  constexpr bool RESW = TcRes<ORIENT, EPI>::value;
  const int STAGES = RESW ? p.res_stages : Cfg::STAGES;
  constexpr int STAGE_BYTES = RESW ? Cfg::RES_STAGE_BYTES : Cfg::STAGE_BYTES;
  constexpr int EW = TcEpi<ORIENT, BN>::EW, TC_THREADS = TcEpi<ORIENT, BN>::THREADS;
  constexpr int PARTS = EW / 4;        // epilogue warps per TMEM lane quadrant
  constexpr int HALF = BN / PARTS;     // columns per epilogue warp
  constexpr int NCHUNK = HALF / 32;    // 32-column chunks per epilogue warp
  constexpr bool PIPE = (EW == 8);
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be 64, 128 or 256");
  static_assert(HALF % 32 == 0, "each epilogue warp needs whole 32-column chunks");
  static_assert(ORIENT == CH_ON_LANES || HALF == 32, "point-on-lanes epilogues handle one 32-column chunk per warp");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();                 // the 128B-swizzled tiles need a 1 KB aligned base
  const uint32_t tiles_base = smem_base + 1024;    // barriers live in the first 1 KB
  // barrier block: full[4], empty[4], tmem_full[2], tmem_empty[2], resident-weights barrier, tmem base pointer
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 32;
  const uint32_t bar_tfull = smem_base + 64, bar_tempty = bar_tfull + 16;
  const uint32_t bar_res = bar_tempty + 16;
  const uint32_t tmem_slot = bar_res + 8;
  // layout after the barrier block: [resident weights (RESW)] [stage ring] [output staging (PT)] [bias]
  const uint32_t res_base = tiles_base;
  const uint32_t ring_base = RESW ? tiles_base + (uint32_t)((p.K / TC_BK) * Cfg::NB_BYTES * Cfg::ARR) : tiles_base;
  const uint32_t ostage_base = ring_base + (uint32_t)STAGES * STAGE_BYTES;
  const int out_bufs = (ORIENT == PT_ON_LANES) ? p.out_bufs : 0;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (ostage_base - smem_base) + (uint32_t)(Cfg::OUT_BYTES * out_bufs));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_slabs = p.K / TC_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&ma_hi); prefetch_tmap(&nb_hi);
    if (NPROD == 3) { prefetch_tmap(&ma_lo); prefetch_tmap(&nb_lo); }
    if (ORIENT == PT_ON_LANES) { prefetch_tmap(&out_hi); if (NPROD == 3) prefetch_tmap(&out_lo); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, EW); }
    mbar_init(bar_res, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (p.bias && EPI != EPI_SPLIT_STREAM) {  // per-channel bias of the whole layer staged once (FC layers read it directly)
    const int nbias = (ORIENT == CH_ON_LANES) ? p.m_tiles * 128 : p.n_tiles * BN;
    for (int i = threadIdx.x; i < nbias; i += TC_THREADS) s_bias[i] = p.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // everything above overlapped the previous kernel's tail; upstream data is touched only below

  // tile t of this CTA's sequence: blockIdx.x, blockIdx.x + gridDim.x, ...  With resident weights the grid is
  // a multiple of n_tiles, so t % n_tiles (the weight tile) is the same for every tile of a CTA.
  auto tile_coords = [&](int t, int& mi, int& ni) {
    if (ORIENT == CH_ON_LANES) { mi = t % p.m_tiles; ni = t / p.m_tiles; }  // channel tile fastest
    else { ni = t % p.n_tiles; mi = t / p.n_tiles; }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // A single thread can only keep ~30 B/clk/SM of TMA traffic in flight (tools/micro/tma_probe.cu: the
    // issue path serialises per thread), so the up to four boxes of a stage are issued by four lanes in
    // parallel: lane 0 = M-side hi (+ expect_tx), 1 = N-side hi, 2 = M-side lo, 3 = N-side lo.
    if (lane < 4) {
      int stage = 0; uint32_t phase = 0;
      const bool is_lo = lane >= 2, is_nb = (lane & 1) != 0;
      const bool active = (NPROD == 3 || !is_lo) && !(RESW && is_nb);
      if (RESW && blockIdx.x < total_tiles && lane < 2) {  // the CTA's weight tile, all K slabs, once (lane 0 hi, lane 1 lo)
        const int ni0 = blockIdx.x % p.n_tiles;
        if (lane == 0) mbar_expect_tx(bar_res, (uint32_t)(k_slabs * Cfg::NB_BYTES * Cfg::ARR));
        if (lane == 0 || NPROD == 3)
          for (int ks = 0; ks < k_slabs; ++ks) {
            const uint32_t rb = res_base + (uint32_t)(ks * Cfg::NB_BYTES * Cfg::ARR) + (lane ? Cfg::NB_BYTES : 0);
            tma_load_2d(rb, lane ? &nb_lo : &nb_hi, ks * TC_BK, ni0 * BN, bar_res);
          }
      }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int mi, ni; tile_coords(t, mi, ni);
        const int nb_row = (ORIENT == PT_ON_LANES && p.nb_per_set) ? set_of_row((long long)mi * 128, p.rows_per_obj, p.rows_per_set) * BN : ni * BN;
        for (int ks = 0; ks < k_slabs; ++ks) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sb = ring_base + stage * STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          if (lane == 0) mbar_expect_tx(full, STAGE_BYTES);
          if (active) {
            // stage layout: [M hi][M lo (x3)] then (streamed weights only) [N hi][N lo (x3)]
            const uint32_t dst = sb + (is_nb ? Cfg::MA_BYTES * Cfg::ARR + (is_lo ? Cfg::NB_BYTES : 0) : (is_lo ? Cfg::MA_BYTES : 0));
            const CUtensorMap* map = is_nb ? (is_lo ? &nb_lo : &nb_hi) : (is_lo ? &ma_lo : &ma_hi);
            tma_load_2d(dst, map, ks * TC_BK, is_nb ? nb_row : (mi + (ORIENT == PT_ON_LANES ? p.mi_in0 : 0)) * 128, full);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc<TcOperand<NPROD>::F16>(128, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      if (RESW && blockIdx.x < total_tiles) mbar_wait(bar_res, 0);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int ks = 0; ks < k_slabs; ++ks) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sb = ring_base + stage * STAGE_BYTES;
          const uint32_t a_hi = sb, a_lo = sb + Cfg::MA_BYTES;
          const uint32_t b_hi = RESW ? res_base + (uint32_t)(ks * Cfg::NB_BYTES * Cfg::ARR) : sb + Cfg::MA_BYTES * Cfg::ARR;
          const uint32_t b_lo = b_hi + Cfg::NB_BYTES;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; ++kk) {
            const uint32_t off = kk * 32;  // 16 bf16 = 32 bytes along K inside the swizzle atom
            umma_bf16(d_tmem, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, (ks | kk) != 0);
            if (NPROD == 3) {
              umma_bf16(d_tmem, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1);
              umma_bf16(d_tmem, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, 1);
            }
          }
          umma_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // warp w may touch TMEM lanes 32*(w%4) .. +31; the PARTS warps of a quadrant split the BN columns
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int lane_row = quad * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int mi, ni; tile_coords(t, mi, ni);
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * HALF);
      if (EPI == EPI_MAX) {
        const int ch = mi * 128 + lane_row;
        float m = -INFINITY;
        tmem_foreach32<NCHUNK, PIPE>(taddr, [&](int, const float* v) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, v[j]);
        });
        if (p.bias) m += s_bias[ch];
        if (p.relu) m = fmaxf(m, 0.f);
        const int set = set_of_row((long long)ni * BN + p.set_row0, p.rows_per_obj, p.rows_per_set);
        atomicMax(p.gmax + (long long)set * p.C + ch, f2key(m));
      } else if (EPI == EPI_STATS) {
        const int ch = mi * 128 + lane_row;
        const long long p0 = (long long)ni * BN + half * HALF;
        const int set = set_of_row(p0, p.rows_per_obj, p.rows_per_set);
        float add = p.bias ? s_bias[ch] : 0.f;
        if (p.rowvec) add += p.rowvec[(long long)set * p.ldrv + ch];
        float s = 0.f, ss = 0.f;
        tmem_foreach32<NCHUNK, PIPE>(taddr, [&](int, const float* v) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = v[j] + add;
            s += x;
            ss = fmaf(x, x, ss);
          }
        });
        s += __shfl_xor_sync(0xffffffffu, s, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        if ((lane & 7) == 0) {
          long long o = (((long long)ni * PARTS + half) * p.stats_ld + p.stats_goff + (ch >> 3)) * 2;
          p.stats[o] = s;
          p.stats[o + 1] = ss;
        }
      } else {
        // PT_ON_LANES epilogues: lane = point row, this warp's 32 channels n0 .. n0+31.  The bf16 hi/lo
        // tile is staged in shared memory as 128B-swizzled [128 rows x 64 channels] boxes and written by
        // the TMA store engine (coalesced, asynchronous).  The GW warps that share a 64-channel box form
        // a group with its own named barrier; one of its threads issues and tracks the group's stores.
        constexpr int GW = (64 / HALF) * 4;               // warps per 64-channel box
        const int box = (half * HALF) / 64;                // box index within the tile
        const bool issuer = (quad == 0) && ((half * HALF) % 64 == 0) && (lane == 0);
        const int n0 = half * HALF;                        // first channel of this warp within the tile
        const int set = set_of_row((long long)mi * 128, p.rows_per_obj, p.rows_per_set);
        float x[32];
        {
          float v[32];
          tmem_ld32(taddr, v);
          tmem_ld_wait32(v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) {
              if (EPI == EPI_SPLIT_STREAM) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ni * BN + n0 + j));  // up to 4096 channels
              else b4 = *reinterpret_cast<const float4*>(s_bias + ni * BN + n0 + j);  // broadcast LDS.128
            }
            x[j + 0] = v[j + 0] + b4.x; x[j + 1] = v[j + 1] + b4.y;
            x[j + 2] = v[j + 2] + b4.z; x[j + 3] = v[j + 3] + b4.w;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (EPI == EPI_SPLIT_STREAM && p.out32 != nullptr) {
            const int grow = mi * 128 + lane_row;
            if (grow < p.rows32) {
              float4* o4 = reinterpret_cast<float4*>(p.out32 + (long long)grow * p.ldo32 + ni * BN + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
          }
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) split16x2<TcOperand<NPROD>::F16>(x[j], x[j + 1], hi[j >> 1], lo[j >> 1]);
        // the stores that last read this staging buffer must be done with it (with two buffers: the group
        // committed two tiles ago, i.e. all but the most recent one)
        const uint32_t obuf = ostage_base + (uint32_t)((out_bufs == 2 ? acc : 0) * Cfg::OUT_BYTES);
        if (issuer) { if (out_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
        named_bar_sync(1 + box, GW * 32);
        {
          const uint32_t bhi = obuf + (uint32_t)(box * 16384 * Cfg::ARR) + (uint32_t)lane_row * 128;
          const uint32_t c0 = (uint32_t)((n0 % 64) / 8), sw = (uint32_t)(lane_row & 7);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_shared_v4(bhi + (((c0 + q) ^ sw) << 4), hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (NPROD == 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(bhi + 16384 + (((c0 + q) ^ sw) << 4), lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + box, GW * 32);
        if (issuer) {
          const uint32_t src = obuf + (uint32_t)(box * 16384 * Cfg::ARR);
          tma_store_2d(&out_hi, src, ni * BN + box * 64, mi * 128);
          if (NPROD == 3) tma_store_2d(&out_lo, src + 16384, ni * BN + box * 64, mi * 128);
          tma_store_commit();
        }
        if (EPI == EPI_SPLIT_MAX) {
          // column max over the warp's 32 rows: butterfly that halves the live columns each step;
          // afterwards x[0] of lane l is the max of channel n0 + l
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) {
            const bool upper = (lane & w) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float mine = upper ? x[j + w] : x[j];
              const float give = upper ? x[j] : x[j + w];
              x[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, give, w));
            }
          }
          atomicMax(p.gmax + (long long)set * p.C + ni * BN + n0 + lane, f2key(x[0]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (ORIENT == PT_ON_LANES && quad == 0 && ((half * HALF) % 64 == 0) && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// element-wise helpers of the tensor-core path
// ------------------------------------------------------------------------------------------------

// front layer (see front3_kernel) writing the bf16 hi/lo split of the 64 channels.  One block = 128 consecutive
// points (never straddling two sets: N is a multiple of 128); 256 threads = 32 point slots x 8 channel groups,
// four points per thread, so weights / transform / points are staged once per 128 points.
constexpr int FRONT_PTS = 128;
template <bool F16>
__global__ void __launch_bounds__(256) front3_split_kernel(const float* __restrict__ q, const float* __restrict__ t3,
                                                           const float* __restrict__ W, const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                           int R, int P, int N) {
  __shared__ float sW[64 * 3];
  __shared__ float sB[64];
  __shared__ float sT[9];
  __shared__ float sQ[FRONT_PTS * 3];
  const int r0 = blockIdx.x * FRONT_PTS;  // first point of the block
  if (threadIdx.x < 192) sW[threadIdx.x] = W[threadIdx.x];
  if (threadIdx.x >= 192) sB[threadIdx.x - 192] = bias[threadIdx.x - 192];
  pdl_wait();  // weights above are constants; the transform and the points come from upstream kernels
  if (threadIdx.x < 9) sT[threadIdx.x] = t3 ? t3[set_of_row(r0, P, N) * 9 + threadIdx.x] : ((threadIdx.x % 4 == 0) ? 1.0f : 0.0f);
  for (int i = threadIdx.x; i < FRONT_PTS * 3; i += 256) sQ[i] = (r0 * 3 + i < R * 3) ? q[(size_t)r0 * 3 + i] : 0.0f;
  __syncthreads();
  const int cg = threadIdx.x & 7;  // channel group (8 channels)
  float w[8][3], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bb[j] = sB[cg * 8 + j];
    w[j][0] = sW[(cg * 8 + j) * 3 + 0]; w[j][1] = sW[(cg * 8 + j) * 3 + 1]; w[j][2] = sW[(cg * 8 + j) * 3 + 2];
  }
#pragma unroll
  for (int it = 0; it < FRONT_PTS / 32; ++it) {
    const int pl = it * 32 + (threadIdx.x >> 3), r = r0 + pl;
    if (r >= R) break;
    const float q0 = sQ[pl * 3 + 0], q1 = sQ[pl * 3 + 1], q2 = sQ[pl * 3 + 2];
    const float x0 = q0 * sT[0] + q1 * sT[3] + q2 * sT[6];
    const float x1 = q0 * sT[1] + q1 * sT[4] + q2 * sT[7];
    const float x2 = q0 * sT[2] + q1 * sT[5] + q2 * sT[8];
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float v0 = fmaxf(bb[j] + w[j][0] * x0 + w[j][1] * x1 + w[j][2] * x2, 0.0f);
      const float v1 = fmaxf(bb[j + 1] + w[j + 1][0] * x0 + w[j + 1][1] * x1 + w[j + 1][2] * x2, 0.0f);
      split16x2<F16>(v0, v1, hi[j >> 1], lo[j >> 1]);
    }
    *reinterpret_cast<uint4*>(out_hi + (size_t)r * 64 + cg * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out_lo + (size_t)r * 64 + cg * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Head of a refinement iteration in one launch (tensor-core modes): [pose update of the previous iteration] + U1 point update +
// stn.conv1.  The three were separate latency-bound launches on the critical path between two tensor-core kernels
// (pose_update_kernel: 8 blocks; update_points_kernel; front3_split_kernel); here every block of 128 points recomputes the pose
// of its object from the heads' outputs (a few hundred flops, same code as pose_update_kernel: pose_update_warp), the block
// that holds the object's first points stores it, and the re-posed points go to q (the second front layer reads them) and
// straight into stn.conv1.
struct IterHeadP {
  TsPoseP prev;        // the previous iteration's heads -> pose (used when have_prev; its pose_out / scale_out are THIS iteration's pose)
  int have_prev;       // 0: first iteration of a call, the pose is read from pose / scale
  const float* pose;   // [B, 12] / [B, 3] current pose when !have_prev
  const float* scale;
  const float* pcl; const float* prior; const int* cls; int n_cls;
  float* q; int* gmax; long long n_keys;
  const float* W; const float* bias;  // stn.conv1 [64, 3], [64]
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int B, N, Np;
};
template <bool F16>
__global__ void __launch_bounds__(256) iter_head_kernel(const IterHeadP p) {
  __shared__ float sW[64 * 3];
  __shared__ float sB[64];
  __shared__ float sQ[FRONT_PTS * 3];
  __shared__ float sPose[12], sScale[3];
  const int P = p.N + p.Np;
  const long long r0 = (long long)blockIdx.x * FRONT_PTS;  // first point of the block (a block never straddles two sets)
  const int b = (int)(r0 / P), rin = (int)(r0 - (long long)b * P);
  if (threadIdx.x < 192) sW[threadIdx.x] = p.W[threadIdx.x];
  if (threadIdx.x >= 192) sB[threadIdx.x - 192] = p.bias[threadIdx.x - 192];
  pdl_wait();  // weights above are constants; everything below comes from upstream kernels
  if (threadIdx.x < 32) {
    if (p.have_prev) {
      float Pn[12], Sn[3];
      pose_update_warp(p.prev, b, (int)threadIdx.x, Pn, Sn);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) sPose[i] = Pn[i];
        sScale[0] = Sn[0]; sScale[1] = Sn[1]; sScale[2] = Sn[2];
        if (rin == 0) {  // one block per object publishes the pose of the finished iteration
          float* Po = p.prev.pose_out + (long long)b * 12;
          float* So = p.prev.scale_out + (long long)b * 3;
#pragma unroll
          for (int i = 0; i < 12; ++i) Po[i] = Pn[i];
          So[0] = Sn[0]; So[1] = Sn[1]; So[2] = Sn[2];
        }
      }
    } else {
      if (threadIdx.x < 12) sPose[threadIdx.x] = p.pose[(long long)b * 12 + threadIdx.x];
      if (threadIdx.x < 3) sScale[threadIdx.x] = p.scale[(long long)b * 3 + threadIdx.x];
    }
  }
  // reset this iteration's column-max keys (as update_points_kernel does)
  for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < p.n_keys; k += (long long)gridDim.x * 256) p.gmax[k] = KEY_NEG_INF;
  __syncthreads();
  if (threadIdx.x < FRONT_PTS) {  // U1 (batch_test.py:78-97): x = pcl - t ; k = R (s * kps)
    const int r = rin + (int)threadIdx.x;
    float o0, o1, o2;
    if (r < p.N) {
      const float* v = p.pcl + ((long long)b * p.N + r) * 3;
      o0 = v[0] - sPose[3]; o1 = v[1] - sPose[7]; o2 = v[2] - sPose[11];
    } else {
      int row = b;
      if (p.cls != nullptr) { row = p.cls[b]; if (row < 0 || row >= p.n_cls) row = 0; }
      const float* v = p.prior + ((long long)row * p.Np + (r - p.N)) * 3;
      const float k0 = v[0] * sScale[0], k1 = v[1] * sScale[1], k2 = v[2] * sScale[2];
      o0 = sPose[0] * k0 + sPose[1] * k1 + sPose[2] * k2;
      o1 = sPose[4] * k0 + sPose[5] * k1 + sPose[6] * k2;
      o2 = sPose[8] * k0 + sPose[9] * k1 + sPose[10] * k2;
    }
    float* o = p.q + (r0 + threadIdx.x) * 3;
    o[0] = o0; o[1] = o1; o[2] = o2;
    sQ[threadIdx.x * 3 + 0] = o0; sQ[threadIdx.x * 3 + 1] = o1; sQ[threadIdx.x * 3 + 2] = o2;
  }
  __syncthreads();
  // stn.conv1 (pointnets/pointnet.py:26), no input transform: as front3_split_kernel with T3 = I
  const int cg = threadIdx.x & 7;  // channel group (8 channels)
  float w[8][3], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bb[j] = sB[cg * 8 + j];
    w[j][0] = sW[(cg * 8 + j) * 3 + 0]; w[j][1] = sW[(cg * 8 + j) * 3 + 1]; w[j][2] = sW[(cg * 8 + j) * 3 + 2];
  }
#pragma unroll
  for (int it = 0; it < FRONT_PTS / 32; ++it) {
    const int pl = it * 32 + (threadIdx.x >> 3);
    const long long r = r0 + pl;
    const float q0 = sQ[pl * 3 + 0], q1 = sQ[pl * 3 + 1], q2 = sQ[pl * 3 + 2];
    // the identity transform of front3_split_kernel spelled out with the same operations: x_j = q0*T0j + q1*T1j + q2*T2j
    const float x0 = q0 * 1.0f + q1 * 0.0f + q2 * 0.0f;
    const float x1 = q0 * 0.0f + q1 * 1.0f + q2 * 0.0f;
    const float x2 = q0 * 0.0f + q1 * 0.0f + q2 * 1.0f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float v0 = fmaxf(bb[j] + w[j][0] * x0 + w[j][1] * x1 + w[j][2] * x2, 0.0f);
      const float v1 = fmaxf(bb[j + 1] + w[j + 1][0] * x0 + w[j + 1][1] * x1 + w[j + 1][2] * x2, 0.0f);
      split16x2<F16>(v0, v1, hi[j >> 1], lo[j >> 1]);
    }
    *reinterpret_cast<uint4*>(p.out_hi + (size_t)r * 64 + cg * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p.out_lo + (size_t)r * 64 + cg * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps and launch
// ------------------------------------------------------------------------------------------------
struct TcPair {  // hi/lo bf16 arrays [rows, ld] and their tensor maps (box = 64 x box_rows, 128B swizzle)
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  CUtensorMap map_hi, map_lo;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 [rows, kext] view with row pitch ld (elements); returns false on failure
inline bool tc_make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t kext, uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled fn = tc_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {kext, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int ORIENT, int EPI, int BN, int NPROD>
cudaError_t tc_launch(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& nb_hi, const CUtensorMap& nb_lo,
                      const CUtensorMap& out_hi, const CUtensorMap& out_lo, const TcGemmP& p, int num_sms, cudaStream_t s) {
  using Cfg = TcCfg<ORIENT, BN, NPROD>;
  constexpr bool RESW = TcRes<ORIENT, EPI>::value;
  auto kern = tc_gemm_kernel<ORIENT, EPI, BN, NPROD>;
  static DeviceOnce configured;  // function attributes belong to the device: set them once per device, not once per process
  if (configured.needed()) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          RESW ? Cfg::MAX_SMEM : Cfg::SMEM_BYTES);
    if (st != cudaSuccess) return st;
    configured.done();
  }
  TcGemmP q = p;
  q.out_bufs = Cfg::OUT_BUFS;
  int smem = Cfg::SMEM_BYTES;
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < num_sms ? tiles : num_sms;
  if (RESW) {
    q.res_stages = Cfg::res_stages(p.K);
    q.out_bufs = Cfg::res_out_bufs(p.K);
    if (q.res_stages < 2) return cudaErrorInvalidConfiguration;
    smem = Cfg::res_smem(p.K);
    grid -= grid % p.n_tiles;  // every CTA keeps one weight tile: t % n_tiles must not change along its sequence
  }
  if (grid < 1) return cudaSuccess;
  return launch_pdl(kern, dim3(grid), dim3(TcEpi<ORIENT, BN>::THREADS), (size_t)smem, s, ma_hi, ma_lo, nb_hi, nb_lo, out_hi, out_lo, q);
}

}  // namespace catre

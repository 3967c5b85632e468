// Fused rotation-head kernel (sm_100a): for a tile of 128 points and one of the two axis heads,
//   L0   D0[128 pts x 256 ch] = pf[128 x 64] . W0p_h[256 x 64]^T              (tcgen05, points on TMEM lanes)
//   epi0 u = gelu(D0 * sc + sh)   (GroupNorm folded into sc/sh, per set)      -> bf16 hi/lo written by the
//        epilogue warps straight into shared memory in the 128B-swizzled K-major operand layout
//   L1   D1[mt][128 ch x 128 pts] = W1_h[mt*128.., 256] . u^T,  mt = 0, 1      (tcgen05, channels on lanes)
//   epi1 the rest of the head, straight from the TMEM accumulator (heads/conv_out_per_rot_head.py:131-140): GroupNorm-1 +
//        GELU + neck 256 -> 3 + the learned weighted sum over the point index.  GroupNorm-1 needs statistics over ALL points
//        of the object, i.e. over the 16 items (tiles) of this (object, head), which other CTAs are working on at the same
//        time: every item publishes its partial sums in global memory and bumps a per-(object, head) counter, waits until
//        the counter shows all tiles, finalises the statistics and then reads D1 a second time.  The layer-1 output
//        (4 KB per point in fp32: 268 MB per iteration at 64 objects) is never stored, and the separate rot-tail kernel is gone.
// so the layer-0 activations (2 KB per point) never touch HBM.  The layer-1 MMAs of K-slab ks start as
// soon as epi0 has written slab ks, i.e. they run underneath the GELU work of slabs ks+1.. (one U buffer).
//
// Reference: heads/conv_out_per_rot_head.py:126-135 (layers.0 .. layers.3 of one RotHead), with the
// layer-0 split of SURVEY.md 8(a) R1.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2..17 epilogue (4 per lane quadrant).
// Shared memory: barriers (1 KB) | U: 4 K-slabs x {hi, lo} x [128 rows x 128 B] = 128 KB | ring of 3 x 32 KB slots.
// Why (profiles/r02_rot_fused_timeline.txt): D1 is single-buffered (TMEM is full), so the layer-1 MMAs of the next item wait
// for epi1.  Storing the tile (128 KB of st.global per item) held D1 for 6 k cycles of a 19.5 k period, pushed the weight
// tiles that were queued behind the stores out by another ~5 k and slowed the GELU warps; with the stores compiled out the
// drain took 1.6 k.  Deadlock-freedom of the cross-CTA wait: the grid is persistent with one CTA per SM (every CTA is
// resident or becomes resident without anyone's help), an item publishes BEFORE it waits, and what it waits for are items of
// the same object, which their CTAs reach after finishing items of earlier objects only (induction over the object index).
// TMEM: D0 = columns 0..255, D1[mt] = columns 256 + 128 mt.
#pragma once
#include "tc_kernels.cuh"

namespace catre {

constexpr int RF_EW = 16;                       // epilogue warps of enc_fused_kernel; E0 (GELU) warps of rot_fused_kernel
constexpr int RF_E1W = 8;                       // rot_fused_kernel: E1 warps (pass 1 of the tail + the exchange between CTAs)
constexpr int RF_THREADS = 64 + 32 * RF_EW;
constexpr int RF_SLOT = 32 * 1024;
constexpr int RF_SLOTS = 3;
constexpr int RF_U_BYTES = 128 * 1024;
constexpr int RF_SMEM = 1024 + 1024 + RF_U_BYTES + RF_SLOTS * RF_SLOT;  // barriers + alignment slack + U + ring
// rot_fused_kernel: 2 KB of barriers + tail scratch | U = a ring of two K-slab buffers (hi | lo, 32 KB each) | TMA ring of 5 slots.
// (Round 2a kept all four slabs of an item in U and had 3 ring slots: the layer-1 phase then waited for weight tiles -- 384 KB
// of W0 / W1 / pf per item through a ring that shallow -- for longer than its MMAs took, profiles/r02_rot_fused_timeline.txt.)
constexpr int ROT_SLOTS = 5;
constexpr int ROT_U_BYTES = 64 * 1024;
constexpr int ROT_SMEM = 2048 + 1024 + ROT_U_BYTES + ROT_SLOTS * RF_SLOT;  // = 227 KB

struct RotFusedP {
  int tiles;             // R / 128
  int rows_per_set;      // n_obs
  int rows_per_obj;      // P = n_obs + n_prior
  const float* gn_scale; // [S][512]  GroupNorm-0 scale per (set, channel)
  const float* gn_shift; // [S][512]  shift with the per-set constant folded in
  const float* bias1;    // [512]     layers.3 bias, both heads
  float* stats;          // [R/128][64 groups][2]  GroupNorm-1 partial sums of every item, exchanged between the CTAs
  int* obj_count;        // [B][2]   items of (object, head) that have published their sums; zero before the launch
  const float* gn1_gamma; const float* gn1_beta;  // [512]
  const float* neck_w;   // [2][3][256]
  const float* neck_b;   // [2][3]
  const float* wp;       // [2][P]   conv_p weights per head and point index
  float* partial;        // [B][P/128][6]  per-item contributions to the two heads' 3-vectors (summed by pose_update_kernel)
  float* a1t;            // split tail: [B][P/4][512][4] fp32 layer-1 output + bias; fused tail: debug tap only (else null)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// Schedule (software-pipelined over the CTA's work items j; half A = layer-0 channels 0..127 = U slabs 0-1,
// half B = channels 128..255 = slabs 2-3; D0a / D0b = TMEM columns 0..127 / 128..255):
//   MMA warp      L0(0,A) L0(0,B) | L1(j,A) L0(j+1,A) L1(j,B) L0(j+1,B) | ...
//   E0 warps      E0(0,s0..s3)    | E0(j+1,s0) E0(j+1,s1) E0(j+1,s2) E0(j+1,s3) | ...      (8 warps: D0 -> U)
//   E1 warps                      |            E1(j)  (whenever D1 of item j is complete) | ...      (8 warps: D1 -> a1T, stats)
// The two epilogue groups are independent instruction streams: the GELU work of item j+1 runs underneath the layer-1
// MMAs of item j AND underneath the drain of item j's D1; the only couplings are the TMEM / U barriers.
constexpr int ROT_THREADS = RF_THREADS + 32 * RF_E1W;
// FT (fused tail): true = the rot tail runs out of TMEM as described above; false = "split tail": E1 adds the bias, takes the
// GroupNorm-1 partial sums and stores the fp32 tile as a1T [B][P/4][512][4] for rot_tail_t_kernel (the round-2a design).
// Measured (64 / 256 objects, N = 1024; 256 objects, N = 2048, K = 8): split 3.33 / 12.9 / 49.6 ms per step, fused 3.45 / 13.1 /
// 52.0 ms, fused 6 % faster at 8 objects -- the standalone tail kernel runs its GELU sweep at 7.3 GELU/clk/SM (48 resident warps
// per SM, streaming), the 24 epilogue warps of this kernel at 2.5 -- so split is the default and the fused tail is the choice
// when the 4 KB per point of a1T (1.07 GB at 256 objects, 540 MB of DRAM traffic per iteration at 64) matters more.
template <int NPROD, bool FT>
__global__ void __launch_bounds__(ROT_THREADS, 1)
rot_fused_kernel(const __grid_constant__ CUtensorMap pf_hi, const __grid_constant__ CUtensorMap pf_lo,
                 const __grid_constant__ CUtensorMap w0_hi, const __grid_constant__ CUtensorMap w0_lo,
                 const __grid_constant__ CUtensorMap w1_hi, const __grid_constant__ CUtensorMap w1_lo, const RotFusedP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t u_base = (smem_base + 2048 + 1023) & ~1023u;  // barriers and the tail's scratch live in the first 2 KB
  const uint32_t ring_base = u_base + ROT_U_BYTES;
  // barriers (8 B each): full[5] empty[5] d0_full[2] d0_empty[2] d1_full d1_empty u_full[4], TMEM base slot, stats_ready,
  // pass2_done, u_empty[2]
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 40;
  const uint32_t bar_d0_full = smem_base + 80, bar_d0_empty = smem_base + 96;
  const uint32_t bar_d1_full = smem_base + 112, bar_d1_empty = smem_base + 120;
  const uint32_t bar_u_full = smem_base + 128;  // 4 barriers (one per K slab of an item)
  const uint32_t tmem_slot = smem_base + 160;
  const uint32_t bar_stats_ready = smem_base + 168, bar_pass2_done = smem_base + 176;
  const uint32_t bar_u_empty = smem_base + 184;  // 2 barriers (one per U buffer)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + 160);
  // tail scratch: conv_p weights of the item's 128 points | GroupNorm-1 (rstd, mean) of the head's 32 groups | per-warp neck
  // partials [24][3]
  float* s_wp = reinterpret_cast<float*>(smem_raw + 256);
  float2* s_grp = reinterpret_cast<float2*>(smem_raw + 768);   // [32]
  float* s_part = reinterpret_cast<float*>(smem_raw + 1024);    // [24 warps][3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ht = p.tiles * 2;  // (tile, head) work items
  const int n_items = (n_ht > (int)blockIdx.x) ? (n_ht - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto item_ht = [&](int j) { return (int)blockIdx.x + j * (int)gridDim.x; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&pf_hi); prefetch_tmap(&w0_hi); prefetch_tmap(&w1_hi);
    if (NPROD == 3) { prefetch_tmap(&pf_lo); prefetch_tmap(&w0_lo); prefetch_tmap(&w1_lo); }
    for (int i = 0; i < ROT_SLOTS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    // RF_EW warps write U (E0 group), RF_E1W warps drain D1 (E1 group)
    for (int i = 0; i < 2; ++i) { mbar_init(bar_d0_full + 8 * i, 1); mbar_init(bar_d0_empty + 8 * i, RF_EW); }
    mbar_init(bar_d1_full, 1); mbar_init(bar_d1_empty, FT ? RF_EW + RF_E1W : RF_E1W);  // D1 is released by pass 2 of the tail (all 24 warps) / by the drain
    mbar_init(bar_stats_ready, 1); mbar_init(bar_pass2_done, RF_EW + RF_E1W);
    for (int i = 0; i < 4; ++i) mbar_init(bar_u_full + 8 * i, RF_EW);
    for (int i = 0; i < 2; ++i) mbar_init(bar_u_empty + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  constexpr uint32_t SLOT_TX = (NPROD == 3) ? 32768 : 16384;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer (same order as the MMA warp consumes) =====================
    // lane 0 issues the hi box (+ expect_tx), lane 1 the lo box: one thread alone sustains only ~30 B/clk/SM
    if (lane < (NPROD == 3 ? 2 : 1) && n_items > 0) {
      int slot = 0; uint32_t phase = 0;
      auto load2 = [&](const CUtensorMap* hi, const CUtensorMap* lo, int c0, int c1) {  // one slot: hi | lo, 16 KB each
        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
        const uint32_t sb = ring_base + slot * RF_SLOT, full = bar_full + 8 * slot;
        if (lane == 0) { mbar_expect_tx(full, SLOT_TX); tma_load_2d(sb, hi, c0, c1, full); }
        else tma_load_2d(sb + 16384, lo, c0, c1, full);
        if (++slot == ROT_SLOTS) { slot = 0; phase ^= 1; }
      };
      auto load_l0 = [&](int j, int half) {  // point features of the tile + 128 rows of W0p_h
        const int ht = item_ht(j), tile = ht >> 1, h = ht & 1;
        load2(&pf_hi, &pf_lo, 0, tile * 128);
        load2(&w0_hi, &w0_lo, 0, h * 256 + half * 128);
      };
      auto load_l1 = [&](int j, int half) {  // W1_h[mt*128.., ks*64..] for the half's two K slabs
        const int h = item_ht(j) & 1;
        for (int ks = half * 2; ks < half * 2 + 2; ++ks)
          for (int mt = 0; mt < 2; ++mt) load2(&w1_hi, &w1_lo, ks * 64, h * 256 + mt * 128);
      };
      load_l0(0, 0); load_l0(0, 1);
      for (int j = 0; j < n_items; ++j) {
        load_l1(j, 0);
        if (j + 1 < n_items) load_l0(j + 1, 0);
        load_l1(j, 1);
        if (j + 1 < n_items) load_l0(j + 1, 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_items > 0) {
      constexpr uint32_t idesc = umma_idesc<TcOperand<NPROD>::F16>(128, 128);
      int slot = 0; uint32_t phase = 0;
      auto next = [&]() { if (++slot == ROT_SLOTS) { slot = 0; phase ^= 1; } };
      // L0(j, half): D0[half] = pf . W0p_h[half*128 ..]^T   (K = 64: 4 k-steps, N = 128)
      auto L0 = [&](int j, int half) {
        mbar_wait(bar_d0_empty + 8 * half, ((uint32_t)j & 1) ^ 1);  // E0(j-1) has drained this half of D0
        tc_fence_after();
        const int sx = slot; mbar_wait(bar_full + 8 * slot, phase); next();
        const int sw = slot; mbar_wait(bar_full + 8 * slot, phase); next();
        tc_fence_after();
        const uint32_t x_hi = ring_base + sx * RF_SLOT, x_lo = x_hi + 16384;
        const uint32_t w_hi = ring_base + sw * RF_SLOT, w_lo = w_hi + 16384;
        const uint32_t d0 = tmem_base + (uint32_t)(half * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t off = kk * 32;
          umma_bf16(d0, umma_desc_sw128(x_hi + off), umma_desc_sw128(w_hi + off), idesc, kk != 0);
          if (NPROD == 3) {
            umma_bf16(d0, umma_desc_sw128(x_hi + off), umma_desc_sw128(w_lo + off), idesc, 1);
            umma_bf16(d0, umma_desc_sw128(x_lo + off), umma_desc_sw128(w_hi + off), idesc, 1);
          }
        }
        umma_commit(bar_empty + 8 * sx);
        umma_commit(bar_empty + 8 * sw);
        umma_commit(bar_d0_full + 8 * half);
      };
      // L1(j, half): D1[mt] (+)= W1_h[mt][ks] . U[ks]^T for the half's two K slabs, as the slabs arrive
      auto L1 = [&](int j, int half) {
        const uint32_t par = (uint32_t)j & 1;
        if (half == 0) {  // first write of D1 for this item: E1(j-1) must have drained it
          mbar_wait(bar_d1_empty, par ^ 1);
          tc_fence_after();
        }
        for (int ks = half * 2; ks < half * 2 + 2; ++ks) {
          mbar_wait(bar_u_full + 8 * ks, par);
          tc_fence_after();
          const uint32_t u_hi = u_base + (ks & 1) * 32768, u_lo = u_hi + 16384;
          for (int mt = 0; mt < 2; ++mt) {
            mbar_wait(bar_full + 8 * slot, phase);
            tc_fence_after();
            const uint32_t w_hi = ring_base + slot * RF_SLOT, w_lo = w_hi + 16384;
            const uint32_t d1 = tmem_base + 256 + mt * 128;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t off = kk * 32;
              umma_bf16(d1, umma_desc_sw128(w_hi + off), umma_desc_sw128(u_hi + off), idesc, (ks | kk) != 0);
              if (NPROD == 3) {
                umma_bf16(d1, umma_desc_sw128(w_hi + off), umma_desc_sw128(u_lo + off), idesc, 1);
                umma_bf16(d1, umma_desc_sw128(w_lo + off), umma_desc_sw128(u_hi + off), idesc, 1);
              }
            }
            umma_commit(bar_empty + 8 * slot);
            next();
          }
          umma_commit(bar_u_empty + 8 * (ks & 1));  // E0 may refill this U buffer once these MMAs have read it
        }
        if (half == 1) umma_commit(bar_d1_full);
      };
      L0(0, 0); L0(0, 1);
      for (int j = 0; j < n_items; ++j) {
        L1(j, 0);
        if (j + 1 < n_items) L0(j + 1, 0);
        L1(j, 1);
        if (j + 1 < n_items) L0(j + 1, 1);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // Two groups that run CONCURRENTLY.
    //   warps 2..17 (E0, CUDA-core heavy): turn D0 into the layer-1 operand U (GroupNorm-0 affine + GELU + 16-bit split), and
    //     between the two halves of an item run PASS 2 of the previous item's tail out of D1 (GroupNorm-1 + GELU + conv_p
    //     weighted sum): program order  E0(j, s0) E0(j, s1) PASS2(j-1) E0(j, s2) E0(j, s3).
    //   warps 18..25 (E1): PASS 1 of the tail (GroupNorm-1 partial sums of D1), the exchange of those sums with the other CTAs
    //     of the object, the finalised per-group statistics, and the neck over the E0 warps' partial results.
    // History: round 1 ran E0 and the D1 drain on the same 16 warps in sequence (tensor pipe 38 %); round 2a made them
    // concurrent groups and stored D1 for a separate tail kernel -- those stores held D1 and blocked the TMA queue
    // (profiles/r02_rot_fused_timeline.txt); now nothing of layer 1 is stored.
    const int quad = warp & 3;
    const int lane_row = quad * 32 + lane;
    const uint32_t row_off = (uint32_t)((lane_row >> 3) * 1024 + (lane_row & 7) * 128);
    const int P = p.rows_per_obj, tiles_per_obj = P / 128;
    // PASS 2 of the tail for item j on one warp: thread = channel (TMEM lane) of m-tile mt, the NCH 16-column chunks (points)
    // from column col0 of D1[mt]:  acc = sum_p wp[p] gelu(GN1(D1 + b1)),  then this warp's share of the neck's 3 outputs
    auto pass2_cols = [&](int j, int mt, int col0, int nch, int slot) {
      const int ht = item_ht(j), h = ht & 1;
      const int cl = mt * 128 + lane_row, ch = h * 256 + cl;  // channel within the head / of both heads
      const float add = __ldg(p.bias1 + ch), gam = __ldg(p.gn1_gamma + ch), bet = __ldg(p.gn1_beta + ch);
      const float nw0 = __ldg(p.neck_w + (h * 3 + 0) * 256 + cl), nw1 = __ldg(p.neck_w + (h * 3 + 1) * 256 + cl),
                  nw2 = __ldg(p.neck_w + (h * 3 + 2) * 256 + cl);
      mbar_wait(bar_stats_ready, (uint32_t)j & 1);  // the statistics of the whole object are there (and D1(j) is complete)
      tc_fence_after();
      const float2 rm = s_grp[cl >> 3];  // (rstd, mean) of this channel's group
      const float sc = rm.x * gam, sh = bet - rm.y * sc;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(256 + mt * 128 + col0);
      float acc = 0.f;
      float xa[16], xb[16];
      tmem_ld16(taddr, xa);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < nch) {
          float* x = (c & 1) ? xb : xa;
          tmem_ld_wait16(x);
          if (c + 1 < nch) {
            tmem_ld16(taddr + (c + 1) * 16, (c & 1) ? xa : xb);
          } else {  // D1 is in registers for the last time: the next item's layer-1 MMAs may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_d1_empty);
          }
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 w = *reinterpret_cast<const float4*>(s_wp + col0 + c * 16 + q);  // same address on every lane
            acc = fmaf(w.x, gelu_fast(fmaf(x[q + 0] + add, sc, sh)), acc);
            acc = fmaf(w.y, gelu_fast(fmaf(x[q + 1] + add, sc, sh)), acc);
            acc = fmaf(w.z, gelu_fast(fmaf(x[q + 2] + add, sc, sh)), acc);
            acc = fmaf(w.w, gelu_fast(fmaf(x[q + 3] + add, sc, sh)), acc);
          }
        }
      }
      // neck (linearity: sum_p wp (neck . g_p + nb) = neck . S + nb sum_p wp)
      float r0 = nw0 * acc, r1 = nw1 * acc, r2 = nw2 * acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      }
      if (lane == 0) {
        s_part[slot * 3 + 0] = r0; s_part[slot * 3 + 1] = r1; s_part[slot * 3 + 2] = r2;
        mbar_arrive(bar_pass2_done);  // (release: the stores above are visible to the E1 threads that wait)
      }
    };
    if (warp < 2 + RF_EW) {
      const int part = (warp - 2) >> 2;  // 4 warps per TMEM lane quadrant: 16 of a slab's 64 channels each
      // PASS2(j) on this warp: E0 warps take 48 of the m-tile's 128 points each, the E1 warp of the same quadrant and m-tile
      // the last 32 (see pass2_cols)
      auto pass2 = [&](int j) { pass2_cols(j, part >> 1, (part & 1) * 48, 3, warp - 2); };
      // E0(j, s): lane = point row; slab s (layer-0 channels s*64 .. +63), this warp's 16 channels
      for (int j = 0; j < n_items; ++j) {
        const int ht = item_ht(j), tile = ht >> 1, h = ht & 1;
        const int set = set_of_row((long long)tile * 128, p.rows_per_obj, p.rows_per_set);
        const float* scp = p.gn_scale + (long long)set * 512 + h * 256;
        const float* shp = p.gn_shift + (long long)set * 512 + h * 256;
        bool tail_done = (j == 0);
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          const int half = ks >> 1;
          // PASS2(j-1): as soon as its statistics are there, and at the latest before slab 2 -- slab 2 reuses the U buffer of
          // slab 0, i.e. it waits for the layer-1 MMAs of THIS item, which wait for D1, which PASS2(j-1) releases
          if (FT && !tail_done && (ks == 2 || mbar_try_wait(bar_stats_ready, (uint32_t)(j - 1) & 1))) { pass2(j - 1); tail_done = true; }
          if ((ks & 1) == 0) {
            mbar_wait(bar_d0_full + 8 * half, (uint32_t)j & 1);
            tc_fence_after();
          }
          {
            const int ch0 = ks * 64 + part * 16;
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)ch0, v);
            float4 sc4[4], sh4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // broadcast loads (same address on every lane), overlapped with the TMEM load
              sc4[q] = __ldg(reinterpret_cast<const float4*>(scp + ch0) + q);
              sh4[q] = __ldg(reinterpret_cast<const float4*>(shp + ch0) + q);
            }
            tmem_ld_wait16(v);
            if (ks & 1) {  // this half of D0 is drained: the next item's L0 may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_d0_empty + 8 * half);
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float g0 = gelu_fast(fmaf(v[4 * q + 0], sc4[q].x, sh4[q].x));
              const float g1 = gelu_fast(fmaf(v[4 * q + 1], sc4[q].y, sh4[q].y));
              const float g2 = gelu_fast(fmaf(v[4 * q + 2], sc4[q].z, sh4[q].z));
              const float g3 = gelu_fast(fmaf(v[4 * q + 3], sc4[q].w, sh4[q].w));
              split16x2<TcOperand<NPROD>::F16>(g0, g1, hi[2 * q], lo[2 * q]);
              split16x2<TcOperand<NPROD>::F16>(g2, g3, hi[2 * q + 1], lo[2 * q + 1]);
            }
            // U is a ring of two slab buffers (slab ks -> buffer ks & 1): its previous occupant, two slabs ago, must have been
            // read by the layer-1 MMAs (use m of a buffer waits for the completion of use m - 1; the first use passes)
            mbar_wait(bar_u_empty + 8 * (ks & 1), (uint32_t)((2 * j + (ks >> 1)) & 1) ^ 1);
            // two 16-byte chunks (8 channels each) of this row, 128B-swizzled: chunk' = chunk ^ (row & 7)
            const uint32_t slab = u_base + (ks & 1) * 32768 + row_off;
            const uint32_t c0 = (uint32_t)(part * 2), sw = (uint32_t)(lane_row & 7);
            st_shared_v4(slab + (((c0 + 0) ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(slab + (((c0 + 1) ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
            if (NPROD == 3) {
              st_shared_v4(slab + 16384 + (((c0 + 0) ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
              st_shared_v4(slab + 16384 + (((c0 + 1) ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_u_full + 8 * ks);
        }
      }
      if (FT && n_items > 0) pass2(n_items - 1);
    } else {
      // E1(j): lane = output channel of m-tile mt (thread = channel, the item's 128 points are its TMEM columns).
      const int mt = (warp - 2 - RF_EW) >> 2;
      const int t1 = (int)threadIdx.x - (64 + 32 * RF_EW);  // 0 .. 255 within the E1 group
      for (int j = 0; j < n_items; ++j) {
        const int ht = item_ht(j), tile = ht >> 1, h = ht & 1;
        const long long row0 = (long long)tile * 128;
        const int b = (int)(row0 / P), tile_in_obj = (int)((row0 - (long long)b * P) >> 7);
        const int ch = h * 256 + mt * 128 + lane_row;  // channel in [0, 512)
        const float add = p.bias1[ch];
        // (the previous item's readers of s_wp / s_grp / s_part finished before pass2_done(j-1), which this group waited for)
        if (FT && t1 < 128) s_wp[t1] = __ldg(p.wp + (long long)h * P + (row0 - (long long)b * P) + t1);
        while (!mbar_try_wait(bar_d1_full, (uint32_t)j & 1)) __nanosleep(64);  // idle most of the time: poll politely
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(256 + mt * 128);
        // ---- pass 1: GroupNorm-1 partial sums of y = D1 + b1 over the item's 128 points (four independent chains per sum,
        //      16-column chunks, the tcgen05.ld of the next chunk in flight).  Split tail: this is the drain -- y goes to a1T
        //      (4 consecutive points of a channel = one 16-byte streaming store, 512 contiguous bytes per warp instruction)
        //      and D1 is released as soon as its last chunk is in registers.
        {
          float xa[16], xb[16];
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
          tmem_ld16(taddr, xa);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float* x = (c & 1) ? xb : xa;
            tmem_ld_wait16(x);
            if (c + 1 < 8) {
              tmem_ld16(taddr + (c + 1) * 16, (c & 1) ? xa : xb);
            } else if (!FT) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_d1_empty);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              x[q] += add;
              s4[q & 3] += x[q];
              q4[q & 3] = fmaf(x[q], x[q], q4[q & 3]);
            }
            if (!FT || p.a1t != nullptr) {  // fused tail: debug tap only (tests/test_stages_gpu.py)
              float4* dst = reinterpret_cast<float4*>(p.a1t) + ((row0 >> 2) + c * 4) * 512 + ch;
#pragma unroll
              for (int q = 0; q < 4; ++q) __stcs(dst + q * 512, make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]));
            }
          }
          float s = (s4[0] + s4[1]) + (s4[2] + s4[3]), ss = (q4[0] + q4[1]) + (q4[2] + q4[3]);
          s += __shfl_xor_sync(0xffffffffu, s, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
          if ((lane & 7) == 0)
            *reinterpret_cast<float2*>(p.stats + (((long long)tile * 64 + (ch >> 3)) * 2)) = make_float2(s, ss);
        }
        if (!FT) continue;  // split tail: rot_tail_t_kernel finalises the statistics and does the rest
        // ---- publish, then (first E1 warp) wait for the other tiles of this (object, head) and finalise the statistics of
        //      the head's 32 groups: lane = group; fp64 sum of the object's tiles in tile order, biased variance, eps 1e-5
        named_bar_sync(1, 32 * RF_E1W);  // every E1 thread's sums are written (and s_wp is complete)
        if (warp == 2 + RF_EW) {
          int* cnt = p.obj_count + b * 2 + h;
          if (lane == 0) {
            // release at GPU scope, cumulative over the sums the other E1 threads wrote before the barrier above
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(cnt) : "memory");
            uint32_t spins = 0;
            int seen;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
              if (seen >= tiles_per_obj) break;
              __nanosleep(32);
              if (++spins > (1u << 22)) __trap();  // a scheduling bug must fail the launch, not hang the GPU
            } while (true);
          }
          __syncwarp();
          const float2* st = reinterpret_cast<const float2*>(p.stats) + (long long)b * tiles_per_obj * 64 + h * 32 + lane;
          double s = 0.0, ss = 0.0;
          for (int t0 = 0; t0 < tiles_per_obj; t0 += 16) {  // one L2 round trip per 16 tiles
            float2 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u)  // written by other SMs: read from L2, not L1
              v[u] = (t0 + u < tiles_per_obj) ? __ldcg(st + (long long)(t0 + u) * 64) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) { s += (double)v[u].x; ss += (double)v[u].y; }
          }
          const double n = 8.0 * P, mean = s / n;
          double var = ss / n - mean * mean;
          if (var < 0.0) var = 0.0;
          s_grp[lane] = make_float2((float)(1.0 / sqrt(var + 1e-5)), (float)mean);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_stats_ready);  // release: s_grp and s_wp are visible to the E0 warps that wait
        }
        // ---- pass 2: this warp's 32 points (the E0 warps take 48 each); then three threads sum the 24 warps' partial results
        //      in warp order
        pass2_cols(j, mt, 96, 2, RF_EW + (warp - 2 - RF_EW));
        if (t1 < 3) {
          mbar_wait(bar_pass2_done, (uint32_t)j & 1);
          float wsum = 0.f;
          for (int i = 0; i < 128; ++i) wsum += s_wp[i];
          float sum = 0.f;
          for (int w = 0; w < RF_EW + RF_E1W; ++w) sum += s_part[w * 3 + t1];
          p.partial[((long long)b * tiles_per_obj + tile_in_obj) * 6 + h * 3 + t1] = fmaf(__ldg(p.neck_b + h * 3 + t1), wsum, sum);
        }
        named_bar_sync(1, 32 * RF_E1W);  // s_wp / s_part / s_grp may be rewritten for the next item
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NPROD, bool FT>
cudaError_t rot_fused_launch(const CUtensorMap& pf_hi, const CUtensorMap& pf_lo, const CUtensorMap& w0_hi,
                             const CUtensorMap& w0_lo, const CUtensorMap& w1_hi, const CUtensorMap& w1_lo,
                             const RotFusedP& p, int num_sms, cudaStream_t s) {
  auto kern = rot_fused_kernel<NPROD, FT>;
  static DeviceOnce configured;  // function attributes belong to the device: set them once per device, not once per process
  if (configured.needed()) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ROT_SMEM);
    if (st != cudaSuccess) return st;
    configured.done();
  }
  int items = p.tiles * 2;
  int grid = items < num_sms ? items : num_sms;
  if (grid < 1) return cudaSuccess;
  return launch_pdl(kern, dim3(grid), dim3(ROT_THREADS), (size_t)ROT_SMEM, s, pf_hi, pf_lo, w0_hi, w0_lo, w1_hi, w1_lo, p);
}


// =================================================================================================
// Fused T-Net trunk (sm_100a):  128-wide layer + 1024-wide layer + column max, for one tile of 256 points:
//   L0   D0[sub][128 pts x 128 ch] = X[sub][128 x 64] . W2[128 x 64]^T, sub = 0, 1   (points on TMEM lanes)
//   epi0 u = relu(D0 + b2) -> bf16 hi/lo straight into shared memory (swizzled K-major operand, 256 rows)
//   L1   D1[128 ch x 256 pts] = W3[mt*128.., 128] . u^T for the 8 channel tiles mt     (channels on lanes)
//   epi1 gmax[set][ch] = max over the tile's points of relu(D1 + b3)                   (atomicMax of keys)
// i.e. stn.conv2 + stn.conv3 + max (pointnets/pointnet.py:27-29) or fstn.conv2 + fstn.conv3 + max (:60-62).
// The CTA owns its point tile for all 8 channel tiles, so the 128-wide activations never leave the SM
// and only the 1024 x 128 weights stream (512 KB per 256 points, ~21 B/clk/SM under the layer-1 MMAs).
// TMEM: D0 = columns 0..255; D1 double-buffered in columns 256..511 (even mt) and 0..255 (odd mt: by then epi0
// has drained D0, which the u_full barriers of mt 0 guarantee).  Shared memory and warp roles as in
// rot_fused_kernel.
// =================================================================================================
struct EncFusedP {
  int tiles;             // R / 256
  int rows_per_set;      // n_obs
  int rows_per_obj;      // P = n_obs + n_prior
  const float* bias2;    // [128]
  const float* bias3;    // [1024]
  int* gmax;             // [S][1024] ordered-int keys
};

// Work units of a CTA: unit k = global unit blockIdx.x + k * gridDim.x.  Full rounds are whole items (all 8
// channel tiles of a 256-point tile); when the last round would be at most half full, its items are split
// into two half-units of 4 channel tiles each (the cheap 64->128 layer is recomputed) so the tail spreads
// over twice as many CTAs: 512 items on 148 CTAs take 3 + ~0.55 rounds instead of 4.
struct EncUnit { int item, mt0, mtn; };
__device__ __forceinline__ bool enc_unit(int k, int items, EncUnit& u) {
  const int G = (int)gridDim.x, g = (int)blockIdx.x + k * G;
  const int full = (items / G) * G, rem = items - full;
  const bool split = rem > 0 && 2 * rem <= G;
  if (g < full || !split) { u.item = g; u.mt0 = 0; u.mtn = 8; return g < items; }
  const int v = g - full;
  u.item = full + (v >> 1); u.mt0 = (v & 1) * 4; u.mtn = 4;
  return v < 2 * rem;
}

template <int NPROD>
__global__ void __launch_bounds__(RF_THREADS, 1)
enc_fused_kernel(const __grid_constant__ CUtensorMap x_hi, const __grid_constant__ CUtensorMap x_lo,
                 const __grid_constant__ CUtensorMap w2_hi, const __grid_constant__ CUtensorMap w2_lo,
                 const __grid_constant__ CUtensorMap w3_hi, const __grid_constant__ CUtensorMap w3_lo, const EncFusedP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t u_base = (smem_base + 1024 + 1023) & ~1023u;
  const uint32_t ring_base = u_base + RF_U_BYTES;
  // barriers: full[3] empty[3] d0_full u_full[2] d1_full[2] d1_empty[2], then the TMEM base slot
  const uint32_t bar_full = smem_base, bar_empty = smem_base + 24;
  const uint32_t bar_d0_full = smem_base + 48;
  const uint32_t bar_u_full = smem_base + 56;    // 2
  const uint32_t bar_d1_full = smem_base + 72;   // 2
  const uint32_t bar_d1_empty = smem_base + 88;  // 2
  const uint32_t tmem_slot = smem_base + 112;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&x_hi); prefetch_tmap(&w2_hi); prefetch_tmap(&w3_hi);
    if (NPROD == 3) { prefetch_tmap(&x_lo); prefetch_tmap(&w2_lo); prefetch_tmap(&w3_lo); }
    for (int i = 0; i < RF_SLOTS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_d0_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_u_full + 8 * i, RF_EW);
      mbar_init(bar_d1_full + 8 * i, 1);
      mbar_init(bar_d1_empty + 8 * i, RF_EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  constexpr uint32_t SLOT_TX = (NPROD == 3) ? 32768 : 16384;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane < (NPROD == 3 ? 2 : 1)) {  // lane 0: hi box (+ expect_tx), lane 1: lo box, issued in parallel
      int slot = 0; uint32_t phase = 0;
      auto load2 = [&](const CUtensorMap* hi, const CUtensorMap* lo, int c0, int c1) {  // one slot: hi | lo, 16 KB each
        mbar_wait(bar_empty + 8 * slot, phase ^ 1);
        const uint32_t sb = ring_base + slot * RF_SLOT, full = bar_full + 8 * slot;
        if (lane == 0) { mbar_expect_tx(full, SLOT_TX); tma_load_2d(sb, hi, c0, c1, full); }
        else tma_load_2d(sb + 16384, lo, c0, c1, full);
        if (++slot == RF_SLOTS) { slot = 0; phase ^= 1; }
      };
      EncUnit un;
      for (int k = 0; enc_unit(k, p.tiles, un); ++k) {
        load2(&x_hi, &x_lo, 0, un.item * 256);
        load2(&x_hi, &x_lo, 0, un.item * 256 + 128);
        load2(&w2_hi, &w2_lo, 0, 0);
        for (int mt = un.mt0; mt < un.mt0 + un.mtn; ++mt)
          for (int ks = 0; ks < 2; ++ks) load2(&w3_hi, &w3_lo, ks * 64, mt * 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc0 = umma_idesc<TcOperand<NPROD>::F16>(128, 128);
      constexpr uint32_t idesc1 = umma_idesc<TcOperand<NPROD>::F16>(128, 256);
      int slot = 0; uint32_t phase = 0;
      auto next = [&]() { if (++slot == RF_SLOTS) { slot = 0; phase ^= 1; } };
      uint32_t it_phase = 0;         // per-item barriers (d0_full, u_full)
      uint32_t use[2] = {0, 0};      // uses of each D1 buffer so far (parity of its full / empty barriers)
      EncUnit un;
      for (int k = 0; enc_unit(k, p.tiles, un); ++k) {
        // ---- L0 into columns 0..255 (= D1 buffer 1): wait until its last reader (epi1 of the previous unit's last odd tile) is done
        mbar_wait(bar_d1_empty + 8 * 1, (use[1] & 1) ^ 1);
        tc_fence_after();
        const int sx0 = slot; mbar_wait(bar_full + 8 * slot, phase); next();
        const int sx1 = slot; mbar_wait(bar_full + 8 * slot, phase); next();
        const int sw = slot;  mbar_wait(bar_full + 8 * slot, phase); next();
        tc_fence_after();
        {
          const uint32_t w_hi = ring_base + sw * RF_SLOT, w_lo = w_hi + 16384;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t x_hi_a = ring_base + (sub ? sx1 : sx0) * RF_SLOT, x_lo_a = x_hi_a + 16384;
            const uint32_t d0 = tmem_base + sub * 128;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t off = kk * 32;
              umma_bf16(d0, umma_desc_sw128(x_hi_a + off), umma_desc_sw128(w_hi + off), idesc0, kk != 0);
              if (NPROD == 3) {
                umma_bf16(d0, umma_desc_sw128(x_hi_a + off), umma_desc_sw128(w_lo + off), idesc0, 1);
                umma_bf16(d0, umma_desc_sw128(x_lo_a + off), umma_desc_sw128(w_hi + off), idesc0, 1);
              }
            }
          }
          umma_commit(bar_empty + 8 * sx0);
          umma_commit(bar_empty + 8 * sx1);
          umma_commit(bar_empty + 8 * sw);
          umma_commit(bar_d0_full);
        }
        // ---- L1: 8 channel tiles; even mt -> buffer 0 (columns 256..511), odd mt -> buffer 1 (columns 0..255)
        for (int mt = 0; mt < un.mtn; ++mt) {  // mt = index within the unit; channel tile un.mt0 + mt
          const int buf = mt & 1;
          if (mt != 1) {  // mt 1 is the first writer of buffer 1 after L0: covered by the wait above + u_full
            mbar_wait(bar_d1_empty + 8 * buf, (use[buf] & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t d1 = tmem_base + (buf ? 0u : 256u);
          for (int ks = 0; ks < 2; ++ks) {
            if (mt == 0) { mbar_wait(bar_u_full + 8 * ks, it_phase); tc_fence_after(); }
            mbar_wait(bar_full + 8 * slot, phase);
            tc_fence_after();
            const uint32_t w_hi = ring_base + slot * RF_SLOT, w_lo = w_hi + 16384;
            const uint32_t u_hi = u_base + ks * 65536, u_lo = u_hi + 32768;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t off = kk * 32;
              umma_bf16(d1, umma_desc_sw128(w_hi + off), umma_desc_sw128(u_hi + off), idesc1, (ks | kk) != 0);
              if (NPROD == 3) {
                umma_bf16(d1, umma_desc_sw128(w_hi + off), umma_desc_sw128(u_lo + off), idesc1, 1);
                umma_bf16(d1, umma_desc_sw128(w_lo + off), umma_desc_sw128(u_hi + off), idesc1, 1);
              }
            }
            umma_commit(bar_empty + 8 * slot);
            next();
          }
          umma_commit(bar_d1_full + 8 * buf);
          use[buf]++;
        }
        it_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const int lane_row = quad * 32 + lane;
    uint32_t it_phase = 0;
    uint32_t use[2] = {0, 0};
    EncUnit un;
    for (int k = 0; enc_unit(k, p.tiles, un); ++k) {
      const int set = set_of_row((long long)un.item * 256, p.rows_per_obj, p.rows_per_set);
      // ---- epi0: lane = point row of sub-tile `sub`; this warp's 16 channels of slab ks
      mbar_wait(bar_d0_full, it_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ks = 0; ks < 2; ++ks) {
        const int ch0 = ks * 64 + part * 16;
        float4 b4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias2 + ch0) + j);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * 128 + ch0), v);
          tmem_ld_wait16(v);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float g0 = fmaxf(v[4 * j + 0] + b4[j].x, 0.f), g1 = fmaxf(v[4 * j + 1] + b4[j].y, 0.f);
            const float g2 = fmaxf(v[4 * j + 2] + b4[j].z, 0.f), g3 = fmaxf(v[4 * j + 3] + b4[j].w, 0.f);
            split16x2<TcOperand<NPROD>::F16>(g0, g1, hi[2 * j], lo[2 * j]);
            split16x2<TcOperand<NPROD>::F16>(g2, g3, hi[2 * j + 1], lo[2 * j + 1]);
          }
          const int row = sub * 128 + lane_row;  // row of the 256-row operand slab
          const uint32_t slab = u_base + ks * 65536 + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
          const uint32_t c0 = (uint32_t)(part * 2), sw = (uint32_t)(row & 7);
          st_shared_v4(slab + (((c0 + 0) ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(slab + (((c0 + 1) ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
          if (NPROD == 3) {
            st_shared_v4(slab + 32768 + (((c0 + 0) ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
            st_shared_v4(slab + 32768 + (((c0 + 1) ^ sw) << 4), lo[4], lo[5], lo[6], lo[7]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_u_full + 8 * ks);
      }
      // ---- epi1: lane = output channel of tile mt; this warp's 64 of the 256 points
      for (int mt = 0; mt < un.mtn; ++mt) {
        const int buf = mt & 1;
        mbar_wait(bar_d1_full + 8 * buf, use[buf] & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (buf ? 0u : 256u) + (uint32_t)(part * 64);
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float x[32];
          tmem_ld32(taddr + c * 32, x);
          tmem_ld_wait32(x);
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, x[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_d1_empty + 8 * buf);
        use[buf]++;
        const int ch = (un.mt0 + mt) * 128 + lane_row;
        m = fmaxf(m + __ldg(p.bias3 + ch), 0.f);
        atomicMax(p.gmax + (long long)set * 1024 + ch, f2key(m));
      }
      it_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NPROD>
cudaError_t enc_fused_launch(const CUtensorMap& x_hi, const CUtensorMap& x_lo, const CUtensorMap& w2_hi, const CUtensorMap& w2_lo,
                             const CUtensorMap& w3_hi, const CUtensorMap& w3_lo, const EncFusedP& p, int num_sms, cudaStream_t s) {
  auto kern = enc_fused_kernel<NPROD>;
  static DeviceOnce configured;  // function attributes belong to the device: set them once per device, not once per process
  if (configured.needed()) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_SMEM);
    if (st != cudaSuccess) return st;
    configured.done();
  }
  int grid = p.tiles < num_sms ? p.tiles : num_sms;
  if (grid < 1) return cudaSuccess;
  return launch_pdl(kern, dim3(grid), dim3(RF_THREADS), (size_t)RF_SMEM, s, x_hi, x_lo, w2_hi, w2_lo, w3_hi, w3_lo, p);
}

}  // namespace catre

// fp32 CUDA-core kernels of the CATRE refinement path (precision mode CATRE_PREC_FP32_SIMT, and the
// narrow / per-object stages of every mode).  Semantics follow SURVEY.md appendix A; each kernel cites
// the reference lines it restates.  Written for sm_100a; no library calls.
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace catre {

// "Once per device" flag for one-time kernel configuration (cudaFuncSetAttribute is per device): a process that drives engines
// on several GPUs must configure every kernel on each of them.  Not thread-safe, like the launches it guards.
struct DeviceOnce {
  unsigned long long mask = 0;
  int dev = -1;
  bool needed() {
    if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; return true; }
    return dev < 0 || dev >= 64 || !((mask >> dev) & 1ull);
  }
  void done() { if (dev >= 0 && dev < 64) mask |= 1ull << dev; }
};


// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------

// Programmatic dependent launch: every kernel of the chain is launched with the programmatic-stream-
// serialization attribute, so its CTAs may be scheduled (and run their prologue: barrier init, TMEM allocation,
// tensor-map prefetch, weight staging) while the previous kernel drains; pdl_wait() blocks until that kernel
// has completed and its writes are visible, and must precede the first access to upstream data.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 16-bit hi / residual split of an fp32 value for the tensor-core operands (see TcOperand in tc_kernels.cuh):
// fp16 pair in the 3-product mode, bf16 pair otherwise.
__device__ __forceinline__ void split16(float x, bool f16, unsigned short& hi, unsigned short& lo) {
  if (f16) {
    const float xs = fminf(fmaxf(x, -65504.0f), 65504.0f);
    const __half h = __float2half_rn(xs);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
  }
}

// order-preserving float <-> int key, so a column max over points can use atomicMax(int)
__device__ __forceinline__ int f2key(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
constexpr int KEY_NEG_INF = (int)0x807fffff;  // f2key(-inf)

// exact GELU (nn.GELU default, lib/torch_utils/layers/layer_utils.py:61-95 "gelu")
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU for the tensor-core modes' fused epilogues: one MUFU, 12 FP32 ops, branch-free.
//   gelu(x) = max(x, 0) - |x| * (erfc(|x|/sqrt2) / 2),   erfc(z)/2 = 2^p(|x|),  p = degree-9 fit of
//   log2(erfc(z)) on z in [0, 4.2] re-expressed in |x| = z sqrt2 with the -1 (the 1/2) folded into c0; beyond
//   |x| = 5.94 erfc < 3e-9 and |x| is clamped.  Max |gelu err| vs fp64 over [-12, 12]: 2.4e-7 (the
//   erff-based fp32 GELU itself is 1.2e-6 off), no cancellation on the negative side.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fminf(fabsf(x), 5.939697f);
  float p = fmaf(2.3988423e-07f, z, -5.5500227e-06f);
  p = fmaf(p, z, 5.128636e-05f);
  p = fmaf(p, z, -0.00021167348f);
  p = fmaf(p, z, -0.00010999188f);
  p = fmaf(p, z, 0.0070524085f);
  p = fmaf(p, z, -0.052498225f);
  p = fmaf(p, z, -0.45920548f);
  p = fmaf(p, z, -1.1511058f);
  p = fmaf(p, z, -1.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));  // p in [-28, -1]: no range fix-up needed
  return fmaf(-fabsf(x), e, fmaxf(x, 0.0f));
}

// Row layout of every per-point buffer: object b owns rows [b * P, (b + 1) * P), P = n_obs + n_prior; its first n_obs rows
// are set 2b (observed points), the remaining n_prior rows set 2b + 1 (prior points).  n_obs and n_prior are multiples of
// the kernels' row tiles, so a tile never straddles two sets (NUM_PCL != NUM_KPS is allowed, as in the reference:
// conv_out_per_rot_head.py:112 only ties conv_p to the SUM).
__host__ __device__ __forceinline__ int set_of_row(long long row, int P, int n_obs) {
  const int obj = (int)(row / P);
  return 2 * obj + (((int)(row - (long long)obj * P) >= n_obs) ? 1 : 0);
}

// ----------------------------------------------------------------------------------------------
// U1: per-iteration point update (core/catre/engine/batch_test.py:78-97,
//     lib/pysixd/misc.py:1011-1026):  x = pcl - t ;  k = R (s * kps)
// q layout: set 2b = observed points of object b, set 2b+1 = prior points; [2B, N, 3] point-major.
// ----------------------------------------------------------------------------------------------
// Both point kernels also reset the iteration's column-max keys (gmax, n_keys ints) to -inf.
// cls == nullptr: prior is [B, N, 3] (one prior per object).  Otherwise prior is a table [n_cls, N, 3] and
// object b uses row cls[b] (batch["obj_kps"] = the category's mean shape, engine_utils.py:17-24); a class id
// outside [0, n_cls) reads row 0 here and pose_update_kernel overwrites that object's result with NaN.
__global__ void update_points_kernel(const float* __restrict__ pcl, const float* __restrict__ prior,
                                     const float* __restrict__ pose, const float* __restrict__ scale,
                                     float* __restrict__ q, int B, int N, int Np, int* __restrict__ gmax, long long n_keys,
                                     const int* __restrict__ cls, int n_cls) {
  pdl_wait();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * (N + Np);
  for (long long k = i; k < n_keys; k += (long long)gridDim.x * blockDim.x) gmax[k] = KEY_NEG_INF;
  if (i >= total) return;
  int b = (int)(i / (N + Np));
  int r = (int)(i % (N + Np));
  const float* P = pose + (long long)b * 12;
  float o0, o1, o2;
  if (r < N) {
    const float* p = pcl + ((long long)b * N + r) * 3;
    o0 = p[0] - P[3];
    o1 = p[1] - P[7];
    o2 = p[2] - P[11];
  } else {
    int row = b;
    if (cls != nullptr) {
      row = cls[b];
      if (row < 0 || row >= n_cls) row = 0;
    }
    const float* p = prior + ((long long)row * Np + (r - N)) * 3;
    const float* s = scale + (long long)b * 3;
    float k0 = p[0] * s[0], k1 = p[1] * s[1], k2 = p[2] * s[2];
    o0 = P[0] * k0 + P[1] * k1 + P[2] * k2;
    o1 = P[4] * k0 + P[5] * k1 + P[6] * k2;
    o2 = P[8] * k0 + P[9] * k1 + P[10] * k2;
  }
  float* o = q + i * 3;
  o[0] = o0; o[1] = o1; o[2] = o2;
}

// [K+1, B, 12] poses + [K+1, B, 3] scales -> packed [B, 15] (R|t row-major 12, then s 3) of iteration `it`: the buffer one
// all-gather moves between the GPUs (replaces the pickled-object gather of catre_custom_evaluator.py:202-203)
__global__ void pack_poses_kernel(const float* __restrict__ poses, const float* __restrict__ scales, int B, int it,
                                  float* __restrict__ packed) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 15) return;
  const int b = i / 15, c = i % 15;
  packed[i] = c < 12 ? poses[((size_t)it * B + b) * 12 + c] : scales[((size_t)it * B + b) * 3 + (c - 12)];
}

// forward_once entry: x and tfd_kps arrive already transformed; interleave them into the q layout
__global__ void gather_points_kernel(const float* __restrict__ x_pm, const float* __restrict__ kps_pm,
                                     float* __restrict__ q, int B, int N, int Np, int* __restrict__ gmax, long long n_keys) {
  pdl_wait();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * (N + Np);
  for (long long k = i; k < n_keys; k += (long long)gridDim.x * blockDim.x) gmax[k] = KEY_NEG_INF;
  if (i >= total) return;
  int b = (int)(i / (N + Np));
  int r = (int)(i % (N + Np));
  const float* p = (r < N) ? x_pm + ((long long)b * N + r) * 3 : kps_pm + ((long long)b * Np + (r - N)) * 3;
  float* o = q + i * 3;
  o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
}

// ----------------------------------------------------------------------------------------------
// E1/E2 front layer: optional per-set 3x3 input transform, then 3 -> 64 point-wise conv + ReLU
// (pointnets/pointnet.py:26 stn.conv1 ; :100-103 bmm with T3 then conv1).
//   x'_j = sum_i x_i T3[set][i][j] ;  h[c] = relu(W[c][0..2] . x' + b[c])
// One thread per (point, 4 channels); output [R, 64] fp32.
// ----------------------------------------------------------------------------------------------
__global__ void front3_kernel(const float* __restrict__ q, const float* __restrict__ t3 /*[S,9] or null*/,
                              const float* __restrict__ W /*[64,3]*/, const float* __restrict__ bias,
                              float* __restrict__ out, long long R, int P, int N) {
  pdl_wait();
  __shared__ float sW[64 * 3];
  __shared__ float sB[64];
  for (int i = threadIdx.x; i < 192; i += blockDim.x) sW[i] = W[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sB[i] = bias[i];
  __syncthreads();
  long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long r = gid >> 4;
  int cg = (int)(gid & 15);
  if (r >= R) return;
  float x0 = q[r * 3 + 0], x1 = q[r * 3 + 1], x2 = q[r * 3 + 2];
  if (t3 != nullptr) {
    const float* T = t3 + (long long)set_of_row(r, P, N) * 9;
    float y0 = x0 * T[0] + x1 * T[3] + x2 * T[6];
    float y1 = x0 * T[1] + x1 * T[4] + x2 * T[7];
    float y2 = x0 * T[2] + x1 * T[5] + x2 * T[8];
    x0 = y0; x1 = y1; x2 = y2;
  }
  float4 o;
  float* po = reinterpret_cast<float*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = cg * 4 + j;
    float v = sB[c] + sW[c * 3 + 0] * x0 + sW[c * 3 + 1] * x1 + sW[c * 3 + 2] * x2;
    po[j] = fmaxf(v, 0.0f);
  }
  *reinterpret_cast<float4*>(out + r * 64 + cg * 4) = o;
}

// ----------------------------------------------------------------------------------------------
// Generic fp32 point-wise layer:  out[r, c] = act( sum_k f(A[r, k]) * W[c, k] + bias[c] + rowvec[set(r), c] )
// 128 x BN output tile, K chunks of 16, 256 threads, 8 x (BN/16) micro-tile.
// ----------------------------------------------------------------------------------------------
enum { A_PLAIN = 0, A_KEY = 1, A_GN_GELU = 2 };

struct GemmP {
  const float* A; int lda;
  const float* W; int wcs; long long w_set_stride;  // W[c][k] at W[c * wcs + k] (+ set * w_set_stride)
  const float* bias;
  const float* rowvec; int ldrv;
  float* out; int ldo;
  int* gmax;       // [S, C] ordered-int keys (column max over the points of a set)
  float* stats;    // [R/128, stats_ld, 2] per-row-tile GroupNorm partials (sum, sum of squares)
  int stats_ld, stats_goff;  // groups per row tile in the stats buffer, first group of this launch
  const float* gn_scale; const float* gn_shift; int ldgn;  // A_GN_GELU: per (object, k)
  int R, C, K;
  int rows_per_set, rows_per_obj;
  int relu;
};

template <int BN, int AMODE>
__global__ void __launch_bounds__(256) pw_gemm_kernel(GemmP p) {
  pdl_wait();
  constexpr int BM = 128, BK = 16;
  constexpr int TN = BN / 16;           // 8 or 4 columns per thread
  constexpr int NCH = TN / 4;           // 4-column chunks per thread (2 or 1)
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  __shared__ float red[16][BN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
  // rows_per_obj > 1: per-point layers (rows_per_set = n_obs of the row layout above); otherwise one row = one set
  const int set = p.rows_per_obj > 1 ? set_of_row(r0, p.rows_per_obj, p.rows_per_set) : r0 / p.rows_per_set;
  const int obj = r0 / p.rows_per_obj;
  const float* Wb = p.W + (long long)set * p.w_set_stride;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const int k_begin = 0, k_end = p.K;

  // global -> registers for one K chunk (A: 128 rows x 16 k, W: BN channels x 16 k; float4 along k).  The
  // loads of chunk i+1 are issued before the FMAs of chunk i, so their latency hides behind the math.
  float4 areg[2], wreg[BN / 64];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int e = tid + it * 256;  // 0..511
      int row = e >> 2, kq = (e & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + row < p.R) {
        const float* src = p.A + (long long)(r0 + row) * p.lda + k0 + kq;
        if (AMODE == A_KEY) {
          int4 kv = *reinterpret_cast<const int4*>(src);
          v = make_float4(key2f(kv.x), key2f(kv.y), key2f(kv.z), key2f(kv.w));
        } else {
          v = *reinterpret_cast<const float4*>(src);
          if (AMODE == A_GN_GELU) {
            const float4 sc = *reinterpret_cast<const float4*>(p.gn_scale + (long long)obj * p.ldgn + k0 + kq);
            const float4 sh = *reinterpret_cast<const float4*>(p.gn_shift + (long long)obj * p.ldgn + k0 + kq);
            v.x = gelu_exact(v.x * sc.x + sh.x);
            v.y = gelu_exact(v.y * sc.y + sh.y);
            v.z = gelu_exact(v.z * sc.z + sh.z);
            v.w = gelu_exact(v.w * sc.w + sh.w);
          }
        }
      }
      areg[it] = v;
    }
#pragma unroll
    for (int it = 0; it < BN / 64; ++it) {
      int e = tid + it * 256;
      int col = e >> 2, kq = (e & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + col < p.C) v = *reinterpret_cast<const float4*>(Wb + (long long)(c0 + col) * p.wcs + k0 + kq);
      wreg[it] = v;
    }
  };

  load_chunk(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int e = tid + it * 256;
      int row = e >> 2, kq = (e & 3) * 4;
      As[kq + 0][row] = areg[it].x; As[kq + 1][row] = areg[it].y; As[kq + 2][row] = areg[it].z; As[kq + 3][row] = areg[it].w;
    }
#pragma unroll
    for (int it = 0; it < BN / 64; ++it) {
      int e = tid + it * 256;
      int col = e >> 2, kq = (e & 3) * 4;
      Ws[kq + 0][col] = wreg[it].x; Ws[kq + 1][col] = wreg[it].y; Ws[kq + 2][col] = wreg[it].z; Ws[kq + 3][col] = wreg[it].w;
    }
    __syncthreads();
    if (k0 + BK < k_end) load_chunk(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], w[TN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
        *reinterpret_cast<float4*>(&w[ch * 4]) = *reinterpret_cast<const float4*>(&Ws[kk][ch * (BN / 2) + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias / per-set vector / ReLU
  // thread's rows: ty*4+i (i<4), 64+ty*4+(i-4); columns: chunk ch -> ch*(BN/2) + tx*4 + j
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    int cbase = c0 + ch * (BN / 2) + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = cbase + j;
      float add = 0.f;
      if (c < p.C) {
        if (p.bias) add += p.bias[c];
        if (p.rowvec) add += p.rowvec[(long long)set * p.ldrv + c];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = acc[i][ch * 4 + j] + add;
        if (p.relu) v = fmaxf(v, 0.f);
        acc[i][ch * 4 + j] = v;
      }
    }
  }
  if (p.out) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int row = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (row >= p.R) continue;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        int cbase = c0 + ch * (BN / 2) + tx * 4;
        if (cbase + 3 < p.C && (p.ldo & 3) == 0) {
          *reinterpret_cast<float4*>(p.out + (long long)row * p.ldo + cbase) =
              make_float4(acc[i][ch * 4 + 0], acc[i][ch * 4 + 1], acc[i][ch * 4 + 2], acc[i][ch * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (cbase + j < p.C) p.out[(long long)row * p.ldo + cbase + j] = acc[i][ch * 4 + j];
        }
      }
    }
  }
  if (p.gmax) {  // column max over the tile's rows (all rows of a tile belong to one set)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float m = acc[0][ch * 4 + j];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, acc[i][ch * 4 + j]);
        red[ty][ch * (BN / 2) + tx * 4 + j] = m;
      }
    __syncthreads();
    if (tid < BN && c0 + tid < p.C) {
      float m = red[0][tid];
#pragma unroll
      for (int i = 1; i < 16; ++i) m = fmaxf(m, red[i][tid]);
      atomicMax(p.gmax + (long long)set * p.C + c0 + tid, f2key(m));
    }
    __syncthreads();
  }
  if (p.stats) {  // GroupNorm partial sums per 8-channel group over the tile's 128 rows
    float* red2 = &As[0][0];  // reuse: needs 16 * (BN/4) * 2 floats <= 16*132
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][ch * 4 + j];
          s += v;
          ss = fmaf(v, v, ss);
        }
      int chunk = ch * (BN / 8) + tx;  // 4-column chunk index within the tile
      red2[(ty * (BN / 4) + chunk) * 2 + 0] = s;
      red2[(ty * (BN / 4) + chunk) * 2 + 1] = ss;
    }
    __syncthreads();
    if (tid < BN / 8) {  // one thread per group: chunks 2g, 2g+1, all 16 ty, fixed order
      float s = 0.f, ss = 0.f;
      for (int y = 0; y < 16; ++y)
        for (int h = 0; h < 2; ++h) {
          s += red2[(y * (BN / 4) + tid * 2 + h) * 2 + 0];
          ss += red2[(y * (BN / 4) + tid * 2 + h) * 2 + 1];
        }
      long long o = ((long long)blockIdx.y * p.stats_ld + p.stats_goff + (c0 / 8 + tid)) * 2;
      p.stats[o] = s;
      p.stats[o + 1] = ss;
    }
  }
}

// GroupNorm(32 groups of 8 channels per head) statistics -> per (object, channel) affine
//   y = x * scale + shift,  scale = rstd * gamma, shift = beta - mean * rstd * gamma, biased variance,
//   eps = 1e-5 (torch.nn.GroupNorm; heads/conv_out_per_rot_head.py:100-104)
__global__ void gn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale,
                                   float* __restrict__ shift, int B, int C, int tiles_per_obj, int rows_per_obj) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int G = C / 8;
  if (i >= B * G) return;
  int b = i / G, g = i % G;
  double s = 0.0, ss = 0.0;
  for (int t = 0; t < tiles_per_obj; ++t) {
    long long o = (((long long)b * tiles_per_obj + t) * G + g) * 2;
    s += (double)stats[o];
    ss += (double)stats[o + 1];
  }
  double n = 8.0 * rows_per_obj;
  double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + 1e-5));
  float fmean = (float)mean;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c = g * 8 + j;
    float sc = rstd * gamma[c];
    scale[(long long)b * C + c] = sc;
    shift[(long long)b * C + c] = beta[c] - fmean * sc;
  }
}

// ----------------------------------------------------------------------------------------------
// Small-M fully-connected layer (T-Net FCs, rot g-feature GEMV, ts-head layer 0):
//   out[r, c] = act(sum_k f(A[r, k]) W[c, k] + bias[c]),   R = number of sets (or objects), a few hundred.
// There is too little work per output tile to hide memory latency and too few tiles to fill 148 SMs, so
// K is split over a thread-block CLUSTER of FC_KSPLIT CTAs (grid.z): each CTA reduces K / FC_KSPLIT
// for one 128 x 64 output tile (register-tiled FMA, all of its global loads issued up front in groups of
// 4 chunks), parks its partial tile in its own shared memory, and after a cluster barrier every CTA sums
// 1/FC_KSPLIT of the tile over the cluster's shared memories (DSMEM) in fixed rank order -- deterministic,
// no global scratch, no atomics -- and writes the finished values (+ optional bf16 hi/lo copy).
// ----------------------------------------------------------------------------------------------
constexpr int FC_KSPLIT = 8;

struct FcP {
  const float* A; int lda;      // [R, K] (A_KEY: ordered-int keys of a column max)
  const float* W; int ldw;      // [C, K]
  const float* bias;            // [C] or null
  float* out32;                 // [R, C] or null
  unsigned short* out_hi; unsigned short* out_lo;  // 16-bit hi/lo split of the result [R, C], or null
  int out_f16;                                     // split type: 1 = fp16 pair, 0 = bf16 pair
  int R, C, K;                  // K multiple of 16 * FC_KSPLIT
  int relu;
};

// Tile shapes: 128 x 64 outputs on 256 threads (8 x 4 per thread) when there are many rows, 64 x 32 on 128
// threads (4 x 4 per thread) when R is small, so that a layer still spreads over >= 128 CTAs.
template <int AMODE, int BM, int BN, int NT>
__global__ void __cluster_dims__(1, 1, FC_KSPLIT) __launch_bounds__(NT) fc_cluster_kernel(FcP p) {
  pdl_wait();
  namespace cg = cooperative_groups;
  constexpr int BK = 16, DEPTH = (NT == 128) ? 8 : 4;  // chunks of global loads in flight per thread
  constexpr int TM = BM * BN / NT / 4;     // rows per thread (8 or 4); 4 columns per thread
  constexpr int TXN = BN / 4;              // threads along the columns
  constexpr int A_IT = BM * BK / 4 / NT;   // float4 loads of the A chunk per thread
  constexpr int W_IT = (BN * BK / 4 + NT - 1) / NT;
  static_assert(TM == 8 || TM == 4, "unsupported tile");
  static_assert(BM * BK / 4 % NT == 0 && NT / TXN * (TM == 8 ? 8 : 4) == BM, "tile / thread mismatch");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  __shared__ __align__(16) float part[BM][BN];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();  // == blockIdx.z (cluster spans grid.z)

  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  const int r0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
  const int kz = p.K / FC_KSPLIT, k_begin = rank * kz, nsteps = kz / BK;
  auto row_of = [&](int i) { return (TM == 8) ? ((i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * 4 + i; };

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  float4 areg[DEPTH][A_IT], wreg[DEPTH][W_IT];
  auto load_chunk = [&](int slot, int k0) {
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int e = tid + it * NT, row = e >> 2, kq = (e & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + row < p.R) {
        const float* src = p.A + (long long)(r0 + row) * p.lda + k0 + kq;
        if (AMODE == A_KEY) {
          const int4 kv = *reinterpret_cast<const int4*>(src);
          v = make_float4(key2f(kv.x), key2f(kv.y), key2f(kv.z), key2f(kv.w));
        } else {
          v = *reinterpret_cast<const float4*>(src);
        }
      }
      areg[slot][it] = v;
    }
#pragma unroll
    for (int it = 0; it < W_IT; ++it) {
      const int e = tid + it * NT, col = e >> 2, kq = (e & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < BN && c0 + col < p.C) v = *reinterpret_cast<const float4*>(p.W + (long long)(c0 + col) * p.ldw + k0 + kq);
      wreg[slot][it] = v;
    }
  };

  for (int s0 = 0; s0 < nsteps; s0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
      if (s0 + d < nsteps) load_chunk(d, k_begin + (s0 + d) * BK);
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (s0 + d < nsteps) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
          const int e = tid + it * NT, row = e >> 2, kq = (e & 3) * 4;
          As[kq + 0][row] = areg[d][it].x; As[kq + 1][row] = areg[d][it].y;
          As[kq + 2][row] = areg[d][it].z; As[kq + 3][row] = areg[d][it].w;
        }
#pragma unroll
        for (int it = 0; it < W_IT; ++it) {
          const int e = tid + it * NT, col = e >> 2, kq = (e & 3) * 4;
          if (col < BN) {
            Ws[kq + 0][col] = wreg[d][it].x; Ws[kq + 1][col] = wreg[d][it].y;
            Ws[kq + 2][col] = wreg[d][it].z; Ws[kq + 3][col] = wreg[d][it].w;
          }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float a[TM], w[4];
          *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          if (TM == 8) *reinterpret_cast<float4*>(&a[TM - 4]) = *reinterpret_cast<const float4*>(&As[kk][BM / 2 + ty * 4]);
          *reinterpret_cast<float4*>(&w[0]) = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
  // park the partial tile
#pragma unroll
  for (int i = 0; i < TM; ++i)
    *reinterpret_cast<float4*>(&part[row_of(i)][tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  cluster.sync();
  // CTA `rank` finishes rows rank*BM/8 .. of the tile: (BM/8) x BN values, one float4 per thread
  constexpr int RROWS = BM / FC_KSPLIT, RV = RROWS * BN / 4;
  if (tid < RV) {
    const int row = rank * RROWS + tid / (BN / 4), col = (tid % (BN / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int z = 0; z < FC_KSPLIT; ++z) {
      const float4 t = *reinterpret_cast<const float4*>(cluster.map_shared_rank(&part[row][col], z));
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    const int gr = r0 + row, gc = c0 + col;
    if (gr < p.R) {
      float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gc + j < p.C) {
          float y = f[j] + (p.bias ? p.bias[gc + j] : 0.f);
          if (p.relu) y = fmaxf(y, 0.f);
          const long long o = (long long)gr * p.C + gc + j;
          if (p.out32) p.out32[o] = y;
          if (p.out_hi) split16(y, p.out_f16 != 0, p.out_hi[o], p.out_lo[o]);
        }
    }
  }
  cluster.sync();  // no CTA may exit while its shared memory is still being read by the others
}

// ordered-int max keys [n] -> 16-bit hi/lo split of the float values (operand of the tensor-core FC layers)
__global__ void keys_split_kernel(const int* __restrict__ keys, unsigned short* __restrict__ out_hi,
                                  unsigned short* __restrict__ out_lo, long long n4, int f16) {
  pdl_wait();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int4 k = *reinterpret_cast<const int4*>(keys + i * 4);
  const float f[4] = {key2f(k.x), key2f(k.y), key2f(k.z), key2f(k.w)};
  unsigned short h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split16(f[j], f16 != 0, h[j], l[j]);
  *reinterpret_cast<uint2*>(out_hi + i * 4) = make_uint2(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16));
  *reinterpret_cast<uint2*>(out_lo + i * 4) = make_uint2(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16));
}

// Tensor-core modes: GroupNorm statistics of rot layer 0 -> per (SET, channel) affine applied to the raw
// MMA accumulator D = W0p . pf (the per-set constant cset = W0g . g_set + b0 is folded into the shift):
//   gelu_in = (D + cset) * sc + sh  =  D * sc + (cset * sc + sh)
__global__ void gn_finalize_set_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ cset,
                                       float* __restrict__ scale, float* __restrict__ shift, int B, int C,
                                       int tiles_per_obj, int rows_per_obj, int* __restrict__ obj_count) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (obj_count != nullptr && i < 2 * B) obj_count[i] = 0;  // the fused rot kernel's per-(object, head) publication counters
  int G = C / 8;
  if (i >= B * G) return;
  int b = i / G, g = i % G;
  double s = 0.0, ss = 0.0;
  for (int t = 0; t < tiles_per_obj; ++t) {
    long long o = (((long long)b * tiles_per_obj + t) * G + g) * 2;
    s += (double)stats[o];
    ss += (double)stats[o + 1];
  }
  double n = 8.0 * rows_per_obj;
  double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + 1e-5));
  float fmean = (float)mean;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c = g * 8 + j;
    float sc = rstd * gamma[c];
    float sh = beta[c] - fmean * sc;
#pragma unroll
    for (int q = 0; q < 2; ++q) {  // the two sets (observed, prior) of object b
      long long o = ((long long)(2 * b + q)) * C + c;
      scale[o] = sc;
      shift[o] = fmaf(cset[o], sc, sh);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// R1 tail: GN + GELU of the second rot layer, neck 256 -> 3, learned weighted sum over the point index
// (heads/conv_out_per_rot_head.py:131-140).  a1 [R, 512] = [head x 256 | head y 256] raw layer-3 output.
// Block = 128 rows of one object; warp per row; partial[b][tile][6] (deterministic two-level reduce).
// ----------------------------------------------------------------------------------------------
// By linearity  sum_p wp[p] (neck . g_p + nb) = neck . (sum_p wp[p] g_p) + nb sum_p wp[p]:  each lane
// accumulates the wp-weighted GELU outputs of its 8 channels per head over the block's rows (no
// per-row shuffles); the 256 -> 3 neck is applied once per warp at the end.
__global__ void __launch_bounds__(256) rot_tail_kernel(const float* __restrict__ a1, const float* __restrict__ gn_scale,
                                                       const float* __restrict__ gn_shift,
                                                       const float* __restrict__ neck_w /*[2][3][256]*/,
                                                       const float* __restrict__ neck_b /*[2][3]*/,
                                                       const float* __restrict__ wp /*[2][P]*/, float* __restrict__ partial,
                                                       int P) {
  pdl_wait();
  const int b = blockIdx.y, tile = blockIdx.x, tiles = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float s_part[8][6];
  float sc[2][8], sh[2][8], acc[2][8];
  float wsum[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = h * 256 + lane * 8 + j;
      sc[h][j] = gn_scale[(long long)b * 512 + c];
      sh[h][j] = gn_shift[(long long)b * 512 + c];
      acc[h][j] = 0.f;
    }
#pragma unroll 2
  for (int rr = warp; rr < 128; rr += 8) {
    const int pidx = tile * 128 + rr;
    const float* row = a1 + ((long long)b * P + pidx) * 512;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 v0 = *reinterpret_cast<const float4*>(row + h * 256 + lane * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(row + h * 256 + lane * 8 + 4);
      const float u[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      const float w = __ldg(wp + (long long)h * P + pidx);
      wsum[h] += w;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[h][j] = fmaf(w, gelu_exact(fmaf(u[j], sc[h][j], sh[h][j])), acc[h][j]);
    }
  }
  float d[6];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t = fmaf(neck_w[(h * 3 + k) * 256 + lane * 8 + j], acc[h][j], t);
      d[h * 3 + k] = t;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 6; ++k) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 6; ++k) s_part[warp][k] = fmaf(neck_b[k], wsum[k / 3], d[k]);
  __syncthreads();
  if (threadIdx.x < 6) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_part[w][threadIdx.x];
    partial[((long long)b * tiles + tile) * 6 + threadIdx.x] = s;
  }
}

// Split-tail variant of the tensor-core modes (CATRE_ROT_TAIL=split, the default; see rot_fused_kernel): the fused rot kernel
// stores the rot layer-1 output in fp32 as a1T [B][P/4][512][4]: groups of 4 consecutive points,
// channel-major inside a group, so that the fused rot kernel (lane = channel, 64 points per thread) and this
// kernel (lane = channel too) both move 512 contiguous bytes per warp instruction.  Block = 32 channels
// (one per lane) of one head of one object, the 8 warps split the P points;  S_c = sum_p wp[p] gelu(a1[c][p] sc_c
// + sh_c) is accumulated per lane and the neck is applied to S_c (linearity, see above);
// partial[b][blockIdx.x][6] (other head's 3 = 0).  The GroupNorm-1 statistics are finalised here as well (fp64
// sum of the block's 4 groups x tiles_per_obj partials written by the fused rot kernel), so no separate finalize
// launch is needed.
__global__ void __launch_bounds__(256) rot_tail_t_kernel(const float* __restrict__ a1t, const float* __restrict__ stats /*[R/128][64][2]*/,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta /*[512]*/,
                                                         int tiles_per_obj,
                                                         const float* __restrict__ neck_w /*[2][3][256]*/,
                                                         const float* __restrict__ neck_b /*[2][3]*/,
                                                         const float* __restrict__ wp /*[2][P]*/, float* __restrict__ partial,
                                                         int P, int reverse) {
  extern __shared__ __align__(16) float s_wp[];  // [P]
  __shared__ float s_part[8][3];
  __shared__ float s_sc[32], s_sh[32];
  // reverse: walk the objects from the last one down -- the fused rot kernel wrote them in ascending order, so the last
  // ones are the part of a1T that is still in L2 (-7 % on this kernel at 64 objects, nothing at 256)
  const int b = reverse ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
  const int cg = blockIdx.x, h = cg >> 3;  // 16 blocks per object, 8 per head
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < P / 4; i += 256)  // weights: constants, staged before the dependency wait
    reinterpret_cast<float4*>(s_wp)[i] = __ldg(reinterpret_cast<const float4*>(wp + (long long)h * P) + i);
  pdl_wait();
  if (warp < 4) {  // GroupNorm(32 groups of 8 channels per head): group cg*4 + warp of the object's 64
    const int g = cg * 4 + warp;
    double s = 0.0, ss = 0.0;
    for (int t = lane; t < tiles_per_obj; t += 32) {
      const long long o = (((long long)b * tiles_per_obj + t) * 64 + g) * 2;
      s += (double)stats[o];
      ss += (double)stats[o + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    const double n = 8.0 * P, mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5)), fmean = (float)mean;
    if (lane < 8) {
      const int c = g * 8 + lane;
      const float sc = rstd * gamma[c];
      s_sc[warp * 8 + lane] = sc;
      s_sh[warp * 8 + lane] = beta[c] - fmean * sc;
    }
  }
  __syncthreads();
  const int c = cg * 32 + lane;  // this lane's channel in [0, 512)
  const float sc = s_sc[lane], sh = s_sh[lane];
  const int quads = P / 4, per_warp = quads / 8;  // P is a multiple of 256
  const float4* src = reinterpret_cast<const float4*>(a1t) + ((long long)b * quads + warp * per_warp) * 512 + c;
  const float4* wq = reinterpret_cast<const float4*>(s_wp) + warp * per_warp;
  float acc = 0.f;
#pragma unroll 4
  for (int i = 0; i < per_warp; ++i) {
    const float4 f = __ldcs(src + (long long)i * 512);  // read once: streaming
    const float4 w = wq[i];  // same address on every lane: shared-memory broadcast
    acc = fmaf(w.x, gelu_fast(fmaf(f.x, sc, sh)), acc);
    acc = fmaf(w.y, gelu_fast(fmaf(f.y, sc, sh)), acc);
    acc = fmaf(w.z, gelu_fast(fmaf(f.z, sc, sh)), acc);
    acc = fmaf(w.w, gelu_fast(fmaf(f.w, sc, sh)), acc);
  }
  const int cl = c & 255;
  float r0 = neck_w[(h * 3 + 0) * 256 + cl] * acc;
  float r1 = neck_w[(h * 3 + 1) * 256 + cl] * acc;
  float r2 = neck_w[(h * 3 + 2) * 256 + cl] * acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r0 += __shfl_xor_sync(0xffffffffu, r0, o);
    r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    r2 += __shfl_xor_sync(0xffffffffu, r2, o);
  }
  if (lane == 0) { s_part[warp][0] = r0; s_part[warp][1] = r1; s_part[warp][2] = r2; }
  __syncthreads();
  if (warp == 0) {
    float wsum = 0.f;
    if ((cg & 7) == 0) {  // the neck bias term  nb * sum_p wp[p]  is added once per head
      for (int i = lane; i < P; i += 32) wsum += s_wp[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if (lane < 6) {
      float s = 0.f;
      if (lane / 3 == h) {
        const int d = lane % 3;
        for (int w = 0; w < 8; ++w) s += s_part[w][d];
        s = fmaf(neck_b[h * 3 + d], wsum, s);
      }
      partial[((long long)b * gridDim.x + cg) * 6 + lane] = s;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// H1 + G1 + G2: per object -- translation/size head (heads/fc_trans_size_head.py:61-70), rot6d
// Gram-Schmidt (core/utils/rot_reps.py:34-55) and the pose update
// (models/pose_scale_from_delta_init.py:47-95: image-space, K-aware, cosypose z, additive scale,
// R' = dR R).  One block of 256 threads per object.
//   w0t [1091, 256], w1t [256, 256]: transposed at pack time so threads read coalesced.
// ----------------------------------------------------------------------------------------------
struct TsPoseP {
  const float* ts0;     // [B, 256] ts layer 0 over the 1024 global-feature inputs (+ bias), from the cluster FC
  const int* gmax_pf;   // [2B, 64] keys: max over points of pointfeat
  const float* w0t; const float* g0; const float* be0;  // w0t [1091, 256] transposed layer-0 weights
  const float* w1t; const float* b1; const float* g1; const float* be1;
  const float* wt; const float* bt; const float* ws; const float* bs;
  const float* rot_partial; int rot_tiles;  // [B, tiles, 6]
  const float* convp_bias;                  // [2]
  const float* pose_in; const float* scale_in; const float* K;
  float* pose_out; float* scale_out;
  float* dts;           // [B, 6] raw head outputs (dt 3 | ds 3): written by ts_head, read by pose_update
  const int* cls; int n_cls;  // prior-table entries: an object whose class id is outside [0, n_cls) gets a NaN pose
};

__device__ __forceinline__ float gn8_gelu(float v, float gamma, float beta) {
  // GroupNorm over the 8 consecutive channels held by 8 consecutive lanes, then GELU
  float s = v;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  float mean = s * 0.125f;
  float d = v - mean;
  float q = d * d;
  q += __shfl_xor_sync(0xffffffffu, q, 1);
  q += __shfl_xor_sync(0xffffffffu, q, 2);
  q += __shfl_xor_sync(0xffffffffu, q, 4);
  float rstd = rsqrtf(q * 0.125f + 1e-5f);
  return gelu_exact(d * rstd * gamma + beta);
}

// ts_head_kernel: H1 (independent of the rotation head, so the host runs it on a side stream underneath the
// rot kernels).  pose_update_kernel: G1 + G2, one warp per object, after both heads are done.
__global__ void __launch_bounds__(256) ts_head_kernel(TsPoseP p) {
  pdl_wait();
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ float feat2[68];  // pointfeat max (64) | init scale (3)
  __shared__ float h[256];
  __shared__ float outv[6];
  const float* sc_in = p.scale_in + (long long)b * 3;
  if (t < 64) feat2[t] = key2f(p.gmax_pf[(long long)(2 * b) * 64 + t]);
  if (t >= 64 && t < 67) feat2[t] = sc_in[t - 64];
  __syncthreads();
  {
    float a0 = p.ts0[(long long)b * 256 + t], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < 64; k += 4) {
      a0 = fmaf(p.w0t[(1024 + k + 0) * 256 + t], feat2[k + 0], a0);
      a1 = fmaf(p.w0t[(1024 + k + 1) * 256 + t], feat2[k + 1], a1);
      a2 = fmaf(p.w0t[(1024 + k + 2) * 256 + t], feat2[k + 2], a2);
      a3 = fmaf(p.w0t[(1024 + k + 3) * 256 + t], feat2[k + 3], a3);
    }
    a0 = fmaf(p.w0t[1088 * 256 + t], feat2[64], a0);
    a1 = fmaf(p.w0t[1089 * 256 + t], feat2[65], a1);
    a2 = fmaf(p.w0t[1090 * 256 + t], feat2[66], a2);
    h[t] = gn8_gelu((a0 + a1) + (a2 + a3), p.g0[t], p.be0[t]);
  }
  __syncthreads();
  float a2;
  {
    float c0 = p.b1[t], c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f, c5 = 0.f, c6 = 0.f, c7 = 0.f;
#pragma unroll 4
    for (int k = 0; k < 256; k += 8) {
      c0 = fmaf(p.w1t[(k + 0) * 256 + t], h[k + 0], c0);
      c1 = fmaf(p.w1t[(k + 1) * 256 + t], h[k + 1], c1);
      c2 = fmaf(p.w1t[(k + 2) * 256 + t], h[k + 2], c2);
      c3 = fmaf(p.w1t[(k + 3) * 256 + t], h[k + 3], c3);
      c4 = fmaf(p.w1t[(k + 4) * 256 + t], h[k + 4], c4);
      c5 = fmaf(p.w1t[(k + 5) * 256 + t], h[k + 5], c5);
      c6 = fmaf(p.w1t[(k + 6) * 256 + t], h[k + 6], c6);
      c7 = fmaf(p.w1t[(k + 7) * 256 + t], h[k + 7], c7);
    }
    a2 = gn8_gelu(((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7)), p.g1[t], p.be1[t]);
  }
  __syncthreads();
  h[t] = a2;
  __syncthreads();
  // fc_t (3) and fc_s (3): warp w < 6 computes one output
  int warp = t >> 5, lane = t & 31;
  if (warp < 6) {
    const float* w = (warp < 3) ? p.wt + warp * 256 : p.ws + (warp - 3) * 256;
    float s = 0.f;
    for (int k = lane; k < 256; k += 32) s = fmaf(w[k], h[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) outv[warp] = s + ((warp < 3) ? p.bt[warp] : p.bs[warp - 3]);
  }
  __syncthreads();
  if (t < 6) p.dts[(long long)b * 6 + t] = outv[t];
}

// G1 + G2 for object b, executed by ONE WARP: the two heads' 3-vectors from the rot partial sums (lanes stride over the tiles
// in fixed order, butterfly sum), rot6d Gram-Schmidt, pose update.  Lane 0 ends up with the new pose in Pout[12] / So[3]
// (registers or shared memory of the caller); nothing is written to global memory here.
__device__ __forceinline__ void pose_update_warp(const TsPoseP& p, int b, int lane, float* Pout, float* So) {
  float r6[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float s = 0.f;
    for (int tl = lane; tl < p.rot_tiles; tl += 32) s += p.rot_partial[((long long)b * p.rot_tiles + tl) * 6 + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    r6[j] = s + p.convp_bias[j / 3];
  }
  if (lane == 0) {
    const float* outv = p.dts + (long long)b * 6;
    const float* sc_in = p.scale_in + (long long)b * 3;
    // rot6d -> R (columns x, y, z)
    float nx = sqrtf(r6[0] * r6[0] + r6[1] * r6[1] + r6[2] * r6[2]);
    nx = fmaxf(nx, 1e-12f);
    float x0 = r6[0] / nx, x1 = r6[1] / nx, x2 = r6[2] / nx;
    float z0 = x1 * r6[5] - x2 * r6[4];
    float z1 = x2 * r6[3] - x0 * r6[5];
    float z2 = x0 * r6[4] - x1 * r6[3];
    float nz = fmaxf(sqrtf(z0 * z0 + z1 * z1 + z2 * z2), 1e-12f);
    z0 /= nz; z1 /= nz; z2 /= nz;
    float y0 = z1 * x2 - z2 * x1;
    float y1 = z2 * x0 - z0 * x2;
    float y2 = z0 * x1 - z1 * x0;
    float dR[9] = {x0, y0, z0, x1, y1, z1, x2, y2, z2};
    const float* Pin = p.pose_in + (long long)b * 12;
    float Rin[9] = {Pin[0], Pin[1], Pin[2], Pin[4], Pin[5], Pin[6], Pin[8], Pin[9], Pin[10]};
    float tin[3] = {Pin[3], Pin[7], Pin[11]};
    float Ro[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Ro[i * 3 + j] = dR[i * 3 + 0] * Rin[0 * 3 + j] + dR[i * 3 + 1] * Rin[1 * 3 + j] + dR[i * 3 + 2] * Rin[2 * 3 + j];
    const float* Kb = p.K + (long long)b * 9;
    float fx = Kb[0], fy = Kb[4];
    float zt = outv[2] * tin[2];
    float xt = zt * (outv[0] / fx + tin[0] / tin[2]);
    float yt = zt * (outv[1] / fy + tin[1] / tin[2]);
    Pout[0] = Ro[0]; Pout[1] = Ro[1]; Pout[2] = Ro[2]; Pout[3] = xt;
    Pout[4] = Ro[3]; Pout[5] = Ro[4]; Pout[6] = Ro[5]; Pout[7] = yt;
    Pout[8] = Ro[6]; Pout[9] = Ro[7]; Pout[10] = Ro[8]; Pout[11] = zt;
    So[0] = sc_in[0] + outv[3];
    So[1] = sc_in[1] + outv[4];
    So[2] = sc_in[2] + outv[5];
    const float qnan = __int_as_float(0x7fc00000);
    // IEEE propagation as in the reference.  A non-finite input pose poisons the re-posed points (x = pcl - t,
    // tfd_kps = R (s * kps): batch_test.py:85-97) and torch's relu / max / GroupNorm carry the NaN to every output
    // that depends on them, whereas fmaxf() and the ordered-int atomicMax of this chain drop NaNs.  Restore it here:
    // t or s non-finite -> everything (the ts head sees the observed features and the scale, the rot head both sets);
    // only R non-finite -> only R' (t', s' do not depend on the prior set: WITH_KPS_FEATURE=False).  One REAL275
    // initial pose (of 15,374) has t = 0, for which the reference itself returns NaN from iteration 1 on.
    bool bad_r = false, bad_ts = false;
    for (int i = 0; i < 9; ++i) bad_r |= !isfinite(Rin[i]);
    for (int i = 0; i < 3; ++i) bad_ts |= !isfinite(tin[i]) || !isfinite(sc_in[i]);
    if (bad_r || bad_ts) {
      Pout[0] = Pout[1] = Pout[2] = Pout[4] = Pout[5] = Pout[6] = Pout[8] = Pout[9] = Pout[10] = qnan;
      if (bad_ts) { Pout[3] = Pout[7] = Pout[11] = qnan; So[0] = So[1] = So[2] = qnan; }
    }
    if (p.cls != nullptr && (p.cls[b] < 0 || p.cls[b] >= p.n_cls)) {  // bad class id: make it visible, not plausible
      for (int i = 0; i < 12; ++i) Pout[i] = qnan;
      So[0] = So[1] = So[2] = qnan;
    }
  }
}

__global__ void __launch_bounds__(256) pose_update_kernel(TsPoseP p, int B) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  float Pn[12], Sn[3];
  pose_update_warp(p, b, lane, Pn, Sn);
  if (lane == 0) {
    float* Pout = p.pose_out + (long long)b * 12;
    float* So = p.scale_out + (long long)b * 3;
#pragma unroll
    for (int i = 0; i < 12; ++i) Pout[i] = Pn[i];
    So[0] = Sn[0]; So[1] = Sn[1]; So[2] = Sn[2];
  }
}

}  // namespace catre

// Observed-cloud producer: the step before the refinement path.  For every detected object of an
// image the reference's test data loader back-projects the depth map, keeps the pixels under the object's mask with
// depth > 0 (row-major order), crops them to a ball around the initial translation whose radius grows x1.10 up
// to 9 times until it holds >= 10 points (all points if still empty), repeats the index list until it has
// >= NUM_PCL entries and draws NUM_PCL of them with torch.randperm
// (core/catre/datasets/data_loader.py:773-799, lib/pysixd/misc.py:360-378, core/utils/cat_data_utils.py:209-226,
// 283-318, 380-400).  Here the per-pixel work for all objects of the image runs on the GPU:
//   cloud_hist_kernel     (chunk, object): valid test + distance + radius bin of every pixel -> per-chunk histogram
//   cloud_select_kernel   (object): radius choice k*, number of selected points, per-chunk offsets
//   cloud_compact_kernel  (chunk, object): ordered stream compaction of the selected pixel ids
//   cloud_gather_kernel   back-projects the sampled pixels into pcl [B, n_pts, 3]
// The random draw itself stays on the host (torch.randperm on the CPU generator, as in the reference) so that the
// same seed gives the same cloud bit for bit.  All arithmetic is IEEE fp32 in the reference's operation order
// (no FMA contraction: explicit __f*_rn intrinsics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace catre {

constexpr int CLOUD_THREADS = 256;
constexpr int CLOUD_PPT = 8;                                  // consecutive pixels per thread
constexpr int CLOUD_CHUNK = CLOUD_THREADS * CLOUD_PPT;        // pixels per block
constexpr int CLOUD_MAX_RADII = 16;
constexpr int CLOUD_BINS = CLOUD_MAX_RADII + 1;               // radius bins + "valid, outside every radius"

struct CloudIntr { float fx, fy, cx, cy; };

// misc.py:372-378: ((u - cx) * z) / fx, ((v - cy) * z) / fy, z
__device__ __forceinline__ float3 cloud_backproject(float z, int u, int v, const CloudIntr& k) {
  const float x = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, k.cx), z), k.fx);
  const float y = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, k.cy), z), k.fy);
  return make_float3(x, y, z);
}

// radius bin of pixel `pix` for one object: 0..n_radii-1 = first radius that covers it, n_radii = valid but outside
// all of them, -1 = not a candidate (outside the mask or depth <= 0).
__device__ __forceinline__ int cloud_bin(const float* __restrict__ depth, const uint8_t* __restrict__ mask, int pix, int W,
                                         const CloudIntr& k, float c0, float c1, float c2, const float* radii, int n_radii) {
  const float z = depth[pix];
  if (!(mask[pix] != 0 && z > 0.0f)) return -1;
  const float3 p = cloud_backproject(z, pix % W, pix / W, k);
  const float dx = __fsub_rn(p.x, c0), dy = __fsub_rn(p.y, c1), dz = __fsub_rn(p.z, c2);
  // ((pts - center) ** 2).sum(-1) then sqrt (cat_data_utils.py:284)
  const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  int b = n_radii;
  for (int i = n_radii - 1; i >= 0; --i)
    if (d <= radii[i]) b = i;  // radii are non-decreasing: the smallest index that covers the point
  return b;
}

__global__ void __launch_bounds__(CLOUD_THREADS) cloud_hist_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ masks,
                                                                   CloudIntr k, const float* __restrict__ centers,
                                                                   const float* __restrict__ radii, int n_radii, int HW, int W,
                                                                   int n_chunks, int* __restrict__ hist /*[B][n_chunks][BINS]*/) {
  __shared__ int s_hist[CLOUD_BINS];
  __shared__ float s_r[CLOUD_MAX_RADII];
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (threadIdx.x < CLOUD_BINS) s_hist[threadIdx.x] = 0;
  if (threadIdx.x < n_radii) s_r[threadIdx.x] = radii[b * n_radii + threadIdx.x];
  __syncthreads();
  const float c0 = centers[b * 3 + 0], c1 = centers[b * 3 + 1], c2 = centers[b * 3 + 2];
  const uint8_t* mask = masks + (size_t)b * HW;
  const int p0 = chunk * CLOUD_CHUNK + threadIdx.x * CLOUD_PPT;
#pragma unroll
  for (int j = 0; j < CLOUD_PPT; ++j) {
    const int pix = p0 + j;
    if (pix < HW) {
      const int bin = cloud_bin(depth, mask, pix, W, k, c0, c1, c2, s_r, n_radii);
      if (bin >= 0) atomicAdd(&s_hist[bin], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < CLOUD_BINS) hist[((size_t)b * n_chunks + chunk) * CLOUD_BINS + threadIdx.x] = s_hist[threadIdx.x];
}

// one warp per object.  kstar = first radius index whose ball holds >= 10 points, else the last radius; if that
// ball is empty every valid point is taken (cat_data_utils.py:286-294).  chunk_off = exclusive prefix over the
// chunks of the number of selected pixels, i.e. where each chunk's compacted ids start.
__global__ void cloud_select_kernel(const int* __restrict__ hist, int n_chunks, int n_radii, int* __restrict__ kstar,
                                    int* __restrict__ n_sel, int* __restrict__ chunk_off /*[B][n_chunks]*/) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int* h = hist + (size_t)b * n_chunks * CLOUD_BINS;
  __shared__ int s_tot[CLOUD_BINS];
  __shared__ int s_k;
  if (lane < CLOUD_BINS) {
    int t = 0;
    for (int c = 0; c < n_chunks; ++c) t += h[c * CLOUD_BINS + lane];
    s_tot[lane] = t;
  }
  __syncwarp();
  if (lane == 0) {
    int cum = 0, ks = n_radii - 1, count = 0;
    bool found = false;
    for (int i = 0; i < n_radii; ++i) {
      cum += s_tot[i];
      if (!found && cum >= 10) { ks = i; count = cum; found = true; }
    }
    if (!found) count = cum;                                   // points inside the last radius tried
    if (count == 0) { ks = n_radii; count = cum + s_tot[n_radii]; }  // empty ball: every valid point
    s_k = ks;
    kstar[b] = ks;
    n_sel[b] = count;
  }
  __syncwarp();
  const int ks = s_k;
  // exclusive prefix over chunks (n_chunks is a few hundred at most: one lane walks it)
  if (lane == 0) {
    int off = 0;
    for (int c = 0; c < n_chunks; ++c) {
      chunk_off[(size_t)b * n_chunks + c] = off;
      int s = 0;
      for (int i = 0; i <= ks; ++i) s += h[c * CLOUD_BINS + i];
      off += s;
    }
  }
}

__global__ void __launch_bounds__(CLOUD_THREADS) cloud_compact_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ masks,
                                                                      CloudIntr k, const float* __restrict__ centers,
                                                                      const float* __restrict__ radii, int n_radii, int HW, int W,
                                                                      int n_chunks, const int* __restrict__ kstar,
                                                                      const int* __restrict__ chunk_off,
                                                                      int* __restrict__ sel_pix /*[B][HW]*/) {
  __shared__ float s_r[CLOUD_MAX_RADII];
  __shared__ int s_warp[CLOUD_THREADS / 32];
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (threadIdx.x < n_radii) s_r[threadIdx.x] = radii[b * n_radii + threadIdx.x];
  __syncthreads();
  const float c0 = centers[b * 3 + 0], c1 = centers[b * 3 + 1], c2 = centers[b * 3 + 2];
  const uint8_t* mask = masks + (size_t)b * HW;
  const int ks = kstar[b];
  const int p0 = chunk * CLOUD_CHUNK + threadIdx.x * CLOUD_PPT;
  unsigned flags = 0;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < CLOUD_PPT; ++j) {
    const int pix = p0 + j;
    if (pix < HW) {
      const int bin = cloud_bin(depth, mask, pix, W, k, c0, c1, c2, s_r, n_radii);
      if (bin >= 0 && bin <= ks) { flags |= 1u << j; ++cnt; }
    }
  }
  // block-wide exclusive scan of cnt in thread order (= pixel order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += s_warp[w];
  int pos = chunk_off[(size_t)b * n_chunks + chunk] + base + inc - cnt;
  int* out = sel_pix + (size_t)b * HW;
#pragma unroll
  for (int j = 0; j < CLOUD_PPT; ++j)
    if (flags & (1u << j)) out[pos++] = p0 + j;
}

// pcl[b][j] = back-projection of the (sample[b][j] mod n_sel[b])-th selected pixel: the reference draws from the
// index list repeated 2^m times (cat_data_utils.py:297-298), entry i of which is selected point i mod n_sel.
__global__ void cloud_gather_kernel(const float* __restrict__ depth, CloudIntr k, const int* __restrict__ sel_pix,
                                    const int* __restrict__ n_sel, const long long* __restrict__ sample, int HW, int W, int n_pts,
                                    float* __restrict__ pcl) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pts) return;
  const int n = n_sel[b];
  float3 p = make_float3(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
  if (n > 0) {
    const long long s = sample[(size_t)b * n_pts + j];
    const int pix = sel_pix[(size_t)b * HW + (int)(s % n)];
    p = cloud_backproject(depth[pix], pix % W, pix / W, k);
  }
  float* o = pcl + ((size_t)b * n_pts + j) * 3;
  o[0] = p.x; o[1] = p.y; o[2] = p.z;
}

}  // namespace catre

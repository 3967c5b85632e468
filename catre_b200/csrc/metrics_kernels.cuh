// Pairwise NOCS pose metrics, one thread per (prediction, ground-truth) pair, fp64 like the
// reference's numpy code: 3-D box IoU of the axis-aligned hulls with the 20-step y-axis symmetry search
// (core/catre/engine/test_utils.py:140-205) and the rotation / translation error with the symmetry rules
// (test_utils.py:208-277).  The reference walks these pairs in Python (two nested loops per image, up to 20 numpy
// box transforms per pair); here every pair of every image of a launch is one thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace catre {

struct PairMetricsP {
  const double* pred_RT;     // [P, 16] row-major 4x4 (scalar size baked into the rotation block, NOCS protocol)
  const double* pred_scale;  // [P, 3]  normalised box extents
  const int* pred_cls;       // [P]
  const double* gt_RT;       // [G, 16]
  const double* gt_scale;    // [G, 3]
  const int* gt_cls;         // [G]
  const int* gt_handle;      // [G] handle visibility (mugs)
  const int* pair_pred;      // [M] index into the prediction arrays
  const int* pair_gt;        // [M] index into the ground-truth arrays
  int n_pairs;
  unsigned sym_mask;         // bit c set: class c is symmetric about its y axis (bottle, bowl, can)
  unsigned flip_mask;        // bit c set: class c is symmetric under a 180-degree y flip (phone, eggbox, glue)
  int mug_cls;               // class id of "mug" (-1: none): symmetric when the handle is not visible
  float* iou;                // [M]
  float* deg_shift;          // [M, 2] theta [degrees], shift (fp32 copy; may be null when deg_shift64 is given)
  // shift definition: 0 = |T1 - T2| / cbrt(det(gt_RT[:3,:3]))  (compute_combination_RT_degree_cm_symmetry, test_utils.py:275)
  //                   1 = |T1 - T2| * 100 [cm]                  (compute_RT_degree_cm_symmetry, test_utils.py:687: the one
  //                       compute_independent_mAP -- the metric the NOCS evaluator calls -- is built on)
  int shift_mode;
  double* deg_shift64;       // [M, 2] the same in fp64 (compute_RT_overlaps keeps fp64, test_utils.py:703), or null
};

// axis-aligned hull of the box (+-s/2) under the 4x4 matrix m (row-major), with the homogeneous divide
__device__ __forceinline__ void metrics_hull(const double* m, const double* s, double* lo, double* hi) {
  const double hx = s[0] / 2, hy = s[1] / 2, hz = s[2] / 2;
#pragma unroll
  for (int a = 0; a < 3; ++a) { lo[a] = 1e300; hi[a] = -1e300; }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double x = (c & 2) ? -hx : hx, y = (c & 4) ? -hy : hy, z = (c & 1) ? -hz : hz;
    const double w = m[12] * x + m[13] * y + m[14] * z + m[15];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = (m[4 * a] * x + m[4 * a + 1] * y + m[4 * a + 2] * z + m[4 * a + 3]) / w;
      lo[a] = fmin(lo[a], v);
      hi[a] = fmax(hi[a], v);
    }
  }
}

__device__ __forceinline__ double metrics_aabb_iou(const double* m1, const double* s1, const double* lo2, const double* hi2, double vol2) {
  double lo1[3], hi1[3];
  metrics_hull(m1, s1, lo1, hi1);
  double e[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) e[a] = fmin(hi1[a], hi2[a]) - fmax(lo1[a], lo2[a]);
  const double inter = (fmin(e[0], fmin(e[1], e[2])) < 0) ? 0.0 : e[0] * e[1] * e[2];
  const double vol1 = (hi1[0] - lo1[0]) * (hi1[1] - lo1[1]) * (hi1[2] - lo1[2]);
  return inter / (vol1 + vol2 - inter);
}

__device__ __forceinline__ double metrics_det3(const double* m) {  // upper-left 3x3 of a row-major 4x4
  return m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) + m[2] * (m[4] * m[9] - m[5] * m[8]);
}

__global__ void __launch_bounds__(128) pair_metrics_kernel(PairMetricsP p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n_pairs) return;
  const int i = p.pair_pred[t], j = p.pair_gt[t];
  double m1[16], m2[16], s1[3], s2[3];
#pragma unroll
  for (int k = 0; k < 16; ++k) { m1[k] = p.pred_RT[(size_t)i * 16 + k]; m2[k] = p.gt_RT[(size_t)j * 16 + k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { s1[k] = p.pred_scale[(size_t)i * 3 + k]; s2[k] = p.gt_scale[(size_t)j * 3 + k]; }
  const int c1 = p.pred_cls[i], c2 = p.gt_cls[j], handle = p.gt_handle[j];
  const bool gt_sym = ((p.sym_mask >> c2) & 1u) != 0 || (c2 == p.mug_cls && handle == 0);

  // ---- IoU (compute_3d_iou_new)
  double lo2[3], hi2[3];
  metrics_hull(m2, s2, lo2, hi2);
  const double vol2 = (hi2[0] - lo2[0]) * (hi2[1] - lo2[1]) * (hi2[2] - lo2[2]);
  const bool pred_sym = ((p.sym_mask >> c1) & 1u) != 0 || (c1 == p.mug_cls && handle == 0);
  double iou;
  if (c1 == c2 && pred_sym) {
    iou = 0.0;
    for (int r = 0; r < 20; ++r) {  // RT_1 @ Ry(2 pi r / 20): columns 0 and 2 mix
      const double th = 2.0 * 3.14159265358979323846 * (double)r / 20.0, cs = cos(th), sn = sin(th);
      double mr[16];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        mr[4 * a + 0] = m1[4 * a + 0] * cs - m1[4 * a + 2] * sn;
        mr[4 * a + 1] = m1[4 * a + 1];
        mr[4 * a + 2] = m1[4 * a + 0] * sn + m1[4 * a + 2] * cs;
        mr[4 * a + 3] = m1[4 * a + 3];
      }
      iou = fmax(iou, metrics_aabb_iou(mr, s1, lo2, hi2, vol2));
    }
  } else {
    iou = metrics_aabb_iou(m1, s1, lo2, hi2, vol2);
  }
  p.iou[t] = (float)iou;

  // ---- rotation / translation error (compute_combination_RT_degree_cm_symmetry; class rules of the GT class)
  const double d1 = cbrt(metrics_det3(m1)), d2 = cbrt(metrics_det3(m2));
  double R1[9], R2[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) { R1[3 * a + b] = m1[4 * a + b] / d1; R2[3 * a + b] = m2[4 * a + b] / d2; }
  double theta;
  if (gt_sym) {
    const double y1[3] = {R1[1], R1[4], R1[7]}, y2[3] = {R2[1], R2[4], R2[7]};
    const double dot = y1[0] * y2[0] + y1[1] * y2[1] + y1[2] * y2[2];
    const double n1 = sqrt(y1[0] * y1[0] + y1[1] * y1[1] + y1[2] * y1[2]), n2 = sqrt(y2[0] * y2[0] + y2[1] * y2[1] + y2[2] * y2[2]);
    theta = acos(dot / (n1 * n2));  // unclipped, as in the reference (NaN when rounding pushes it past 1)
  } else {
    // trace(R1 R2^T) = sum_ab R1[a][b] R2[a][b];  with the y-flip diag(-1, 1, -1) in between the b = 0, 2 terms change sign
    double tr = 0.0, trf = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double v = R1[3 * a + b] * R2[3 * a + b];
        tr += v;
        trf += (b == 1) ? v : -v;
      }
    if ((p.flip_mask >> c2) & 1u) theta = fmin(acos((tr - 1.0) / 2.0), acos((trf - 1.0) / 2.0));
    else theta = acos(fmin(fmax((tr - 1.0) / 2.0, -1.0), 1.0));
  }
  theta *= 180.0 / 3.14159265358979323846;
  const double tx = m1[3] - m2[3], ty = m1[7] - m2[7], tz = m1[11] - m2[11];
  const double dist = sqrt(tx * tx + ty * ty + tz * tz);
  const double shift = p.shift_mode == 1 ? dist * 100.0 : dist / d2;
  if (p.deg_shift) {
    p.deg_shift[2 * (size_t)t] = (float)theta;
    p.deg_shift[2 * (size_t)t + 1] = (float)shift;
  }
  if (p.deg_shift64) {
    p.deg_shift64[2 * (size_t)t] = theta;
    p.deg_shift64[2 * (size_t)t + 1] = shift;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Greedy prediction <-> ground-truth matching of the NOCS protocol, one thread per (sub-problem, threshold combination).
// A sub-problem is one (image, class) pair of compute_independent_mAP (test_utils.py:793-880): P predictions in score
// order, G ground truths, a [P, G] pair table.  The loops are sequential in the prediction and in its candidate list but
// independent across sub-problems and thresholds -- REAL275 has ~16 k sub-problems x 4 (IoU) or 3 x 4 (pose) thresholds.
//   mode 0 = compute_3d_matches (test_utils.py:586-614): candidates by descending IoU; a taken ground truth is skipped,
//            IoU below the threshold ends the search, another class is skipped, IoU strictly above the threshold matches.
//   mode 1 = compute_match_from_degree_cm (test_utils.py:734-755): candidates by ascending (degree + shift); taken or
//            other-class ground truths are skipped, so are candidates over either threshold; the first survivor matches.
// The candidate ORDER of every row comes from the host (numpy's argsort of that row: its tie order is what the
// reference's results depend on, and numpy's SIMD sorts are not stable, so it cannot be re-derived here).
// ----------------------------------------------------------------------------------------------------------------
struct MatchP {
  const int* sub_pred_off;  // [n_sub + 1] first prediction of each sub-problem
  const int* sub_gt_off;    // [n_sub + 1]
  const int* sub_pair_off;  // [n_sub + 1] first entry of the sub-problem's row-major [P, G] pair table
  const float* ov;          // mode 0: [pairs] IoU (fp32, as the reference stores it)
  const double* rt;         // mode 1: [pairs, 2] degree, shift (fp64)
  const int* order;         // [pairs] candidate order of every row (indices into the sub-problem's ground truths)
  const int* n_cand;        // [n_pred] candidates of a row (mode 0: after the score_threshold cut)
  const int* pred_cls;      // [n_pred]
  const int* gt_cls;        // [n_gt]
  const double* thr_a; int n_a;  // mode 0: IoU thresholds; mode 1: degree thresholds
  const double* thr_b; int n_b;  // mode 1: shift thresholds (mode 0: n_b = 1, unused)
  int* gt_match;            // [n_a * n_b, n_gt]   index of the matched prediction within the sub-problem, or -1
  int* pred_match;          // [n_a * n_b, n_pred]
  int n_sub, n_pred, n_gt, mode;
};

__global__ void __launch_bounds__(128) match_kernel(MatchP p) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int n_thr = p.n_a * p.n_b;
  if (t >= (long long)p.n_sub * n_thr) return;
  const int sub = (int)(t / n_thr), c = (int)(t % n_thr), ia = c / p.n_b, ib = c % p.n_b;
  const int p0 = p.sub_pred_off[sub], P = p.sub_pred_off[sub + 1] - p0;
  const int g0 = p.sub_gt_off[sub], G = p.sub_gt_off[sub + 1] - g0;
  const int q0 = p.sub_pair_off[sub];
  int* gm = p.gt_match + (size_t)c * p.n_gt + g0;
  int* pm = p.pred_match + (size_t)c * p.n_pred + p0;
  for (int j = 0; j < G; ++j) gm[j] = -1;
  for (int i = 0; i < P; ++i) pm[i] = -1;
  const double ta = p.thr_a[ia], tb = p.mode == 1 ? p.thr_b[ib] : 0.0;
  // mode 0 compares an fp32 IoU with a Python-float threshold: under NumPy >= 2 (NEP 50; the goldens were made with 2.3) the
  // Python float is weakly typed and the comparison runs in fp32.  (NumPy 1.x promoted to fp64; the two differ only when an
  // IoU equals the fp32 rounding of a threshold exactly.)
  const float taf = (float)ta;
  for (int i = 0; i < P; ++i) {
    const int nc = p.n_cand[p0 + i];
    for (int r = 0; r < nc; ++r) {
      const int j = p.order[q0 + i * G + r];
      if (p.mode == 0) {
        if (gm[j] > -1) continue;
        const float iou = p.ov[q0 + i * G + j];
        if (iou < taf) break;
        if (p.pred_cls[p0 + i] != p.gt_cls[g0 + j]) continue;
        if (iou > taf) { gm[j] = i; pm[i] = j; break; }
      } else {
        if (gm[j] > -1 || p.pred_cls[p0 + i] != p.gt_cls[g0 + j]) continue;
        const double* v = p.rt + 2 * (size_t)(q0 + i * G + j);
        if (v[0] > ta || v[1] > tb) continue;  // NaN compares false, as in numpy: an unclipped acos() NaN passes
        gm[j] = i; pm[i] = j;
        break;
      }
    }
  }
}

}  // namespace catre

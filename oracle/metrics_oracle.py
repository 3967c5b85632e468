"""CPU oracle for the pairwise NOCS pose metrics: 3-D box IoU with the y-axis symmetry search,
rotation / translation error with the symmetry rules, and the greedy prediction <-> ground-truth matching.

TEST INFRASTRUCTURE ONLY.  Parity status: PINNED -- tests/test_metrics.py checks this numpy restatement against
tests/golden/golden_metrics.npz, which tests/golden/make_golden_metrics.py produced by calling the unmodified
reference functions (core/catre/engine/test_utils.py: compute_3d_iou_new :140-205, get_3d_bbox :50-94,
transform_coordinates_3d :97-110, compute_combination_RT_degree_cm_symmetry :208-277,
compute_combination_3d_matches :280-389).  Written against those semantics; nothing copied.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

SYM_Y = ("bottle", "bowl", "can")  # rotationally symmetric about the object's y axis
N_ROT = 20                         # IoU symmetry search: 20 rotations of 18 degrees


def _box_corners(scale: np.ndarray, RT: np.ndarray) -> np.ndarray:
    """[3, 8] corners of the box with full extents `scale`, transformed by the 4x4 RT (homogeneous divide)."""
    sx, sy, sz = scale[0] / 2, scale[1] / 2, scale[2] / 2
    c = np.array([[sx, sy, sz], [sx, sy, -sz], [-sx, sy, sz], [-sx, sy, -sz],
                  [sx, -sy, sz], [sx, -sy, -sz], [-sx, -sy, sz], [-sx, -sy, -sz]]).T
    h = RT @ np.vstack([c, np.ones((1, 8))])
    return h[:3] / h[3]


def _aabb_iou(RT_1, RT_2, scales_1, scales_2) -> float:
    """IoU of the axis-aligned hulls of the two transformed boxes (test_utils.py:146-174)."""
    b1, b2 = _box_corners(scales_1, RT_1), _box_corners(scales_2, RT_2)
    max1, min1, max2, min2 = b1.max(1), b1.min(1), b2.max(1), b2.min(1)
    ext = np.minimum(max1, max2) - np.maximum(min1, min2)
    inter = 0 if ext.min() < 0 else np.prod(ext)
    union = np.prod(max1 - min1) + np.prod(max2 - min2) - inter
    return inter / union


def iou_3d(RT_1, RT_2, scales_1, scales_2, handle_visibility, name_1: str, name_2: str) -> float:
    """compute_3d_iou_new: for same-class symmetric objects (and mugs whose handle is not visible) the best IoU over
    20 rotations of box 1 about its y axis, else the plain IoU."""
    if RT_1 is None or RT_2 is None:
        return -1
    if (name_1 in SYM_Y and name_1 == name_2) or (name_1 == "mug" and name_1 == name_2 and handle_visibility == 0):
        best = 0
        for i in range(N_ROT):
            th = 2 * math.pi * i / float(N_ROT)
            ry = np.array([[np.cos(th), 0, np.sin(th), 0], [0, 1, 0, 0], [-np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]])
            best = max(best, _aabb_iou(RT_1 @ ry, RT_2, scales_1, scales_2))
        return best
    return _aabb_iou(RT_1, RT_2, scales_1, scales_2)


def degree_cm(RT_1, RT_2, scale, class_id: int, handle_visibility, synset_names: Sequence[str]) -> np.ndarray:
    """compute_combination_RT_degree_cm_symmetry -> [theta in degrees, |T1 - T2| / scale]."""
    R1 = RT_1[:3, :3] / np.cbrt(np.linalg.det(RT_1[:3, :3]))
    R2 = RT_2[:3, :3] / np.cbrt(np.linalg.det(RT_2[:3, :3]))
    T1, T2 = RT_1[:3, 3], RT_2[:3, 3]
    name = synset_names[class_id]
    if name in ("bottle", "can", "bowl") or (name == "mug" and handle_visibility == 0):
        y = np.array([0, 1, 0])
        y1, y2 = R1 @ y, R2 @ y
        theta = np.arccos(y1.dot(y2) / (np.linalg.norm(y1) * np.linalg.norm(y2)))  # not clipped in the reference
    elif name in ("phone", "eggbox", "glue"):
        R = R1 @ R2.transpose()
        R_rot = R1 @ np.diag([-1.0, 1.0, -1.0]) @ R2.transpose()
        theta = min(np.arccos((np.trace(R) - 1) / 2), np.arccos((np.trace(R_rot) - 1) / 2))
    else:
        theta = np.arccos(np.clip((np.trace(R1 @ R2.transpose()) - 1) / 2, -1.0, 1.0))
    return np.array([theta * 180 / np.pi, np.linalg.norm(T1 - T2) / scale])


def pair_metrics(pred_RTs, pred_scales, pred_cls, gt_RTs, gt_scales, gt_cls, gt_handle, synset_names) -> Tuple[np.ndarray, np.ndarray]:
    """All pairs of one image, stored as the reference stores them: overlaps [P,G] fp32, RT_overlaps [P,G,2] fp32
    (test_utils.py:329-352)."""
    P, G = len(pred_cls), len(gt_cls)
    overlaps = np.zeros((P, G), dtype=np.float32)
    rt = np.zeros((P, G, 2), dtype=np.float32)
    with np.errstate(invalid="ignore"):
        for i in range(P):
            for j in range(G):
                overlaps[i, j] = iou_3d(pred_RTs[i], gt_RTs[j], pred_scales[i], gt_scales[j], gt_handle[j],
                                        synset_names[pred_cls[i]], synset_names[gt_cls[j]])
                rt[i, j] = degree_cm(pred_RTs[i], gt_RTs[j], np.cbrt(np.linalg.det(gt_RTs[j, :3, :3])), gt_cls[j], gt_handle[j],
                                     synset_names)
    return overlaps, rt


def greedy_matches(overlaps, rt, pred_cls, gt_cls, iou_thresholds, degree_thresholds, shift_thresholds, score_threshold=0):
    """The matching loops of compute_combination_3d_matches (:354-387) on score-sorted predictions.
    Returns gt_matches [D,T,S,G], pred_matches [D,T,S,P] (float, -1 = unmatched)."""
    P, G = overlaps.shape
    D, T, S = len(degree_thresholds), len(shift_thresholds), len(iou_thresholds)
    pred_matches = -1 * np.ones([D, T, S, P])
    gt_matches = -1 * np.ones([D, T, S, G])
    for s, iou_t in enumerate(iou_thresholds):
        for d, deg_t in enumerate(degree_thresholds):
            for t, sh_t in enumerate(shift_thresholds):
                for i in range(P):
                    order = np.argsort(overlaps[i])[::-1]
                    low = np.where(overlaps[i, order] < score_threshold)[0]
                    if low.size > 0:
                        order = order[: low[0]]
                    for j in order:
                        if gt_matches[d, t, s, j] > -1:
                            continue
                        if overlaps[i, j] < iou_t or rt[i, j, 0] > deg_t or rt[i, j, 1] > sh_t:
                            break
                        if not pred_cls[i] == gt_cls[j]:
                            continue
                        gt_matches[d, t, s, j] = i
                        pred_matches[d, t, s, i] = j
                        break
    return gt_matches, pred_matches


# ---------------------------------------------------------------------------------------------------------------
# The chain the NOCS evaluator actually calls: compute_independent_mAP (test_utils.py:523-926).  Pinned by
# tests/test_nocs_map.py against tests/golden/golden_mAP.npz (made by the unmodified reference functions,
# tests/golden/make_golden_mAP.py).
# ---------------------------------------------------------------------------------------------------------------
def degree_cm_independent(RT_1, RT_2, class_id: int, handle_visibility, synset_names: Sequence[str]) -> np.ndarray:
    """compute_RT_degree_cm_symmetry (test_utils.py:619-690) -> [theta in degrees, |T1 - T2| * 100]: the same rotation
    rules as degree_cm, the translation error in centimetres instead of object-scale units."""
    theta = degree_cm(RT_1, RT_2, 1.0, class_id, handle_visibility, synset_names)[0]
    return np.array([theta, np.linalg.norm(RT_1[:3, 3] - RT_2[:3, 3]) * 100])


class OracleBackend:
    """The two device stages of catre_b200.nocs_map on the CPU, as plain loops: the per-pair tables
    (compute_3d_iou_new fp32, compute_RT_degree_cm_symmetry fp64) and the two greedy matchers with the candidate
    order supplied by the caller -- the same contract as nocs_map.CudaBackend."""

    def pair_tables(self, ims, synset_names):
        out = []
        with np.errstate(invalid="ignore"):
            for im in ims:
                P, G = len(im["pred_cls"]), len(im["gt_cls"])
                ov = np.zeros((P, G), dtype=np.float32)
                rt = np.zeros((P, G, 2), dtype=np.float64)
                for i in range(P):
                    for j in range(G):
                        ov[i, j] = iou_3d(im["pred_RTs"][i], im["gt_RTs"][j], im["pred_scales"][i], im["gt_scales"][j],
                                          im["gt_handle"][j], synset_names[im["pred_cls"][i]], synset_names[im["gt_cls"][j]])
                        rt[i, j] = degree_cm_independent(im["pred_RTs"][i], im["gt_RTs"][j], im["gt_cls"][j], im["gt_handle"][j],
                                                         synset_names)
                out.append((ov, rt))
        return out

    def match(self, mode, pred_off, gt_off, pair_off, table, order, n_cand, pred_cls, gt_cls, thr_a, thr_b):
        n_sub, n_pred, n_gt = len(pred_off) - 1, int(pred_off[-1]), int(gt_off[-1])
        n_b = max(1, len(thr_b))
        gt_m = -np.ones((len(thr_a) * n_b, n_gt), dtype=np.int32)
        pred_m = -np.ones((len(thr_a) * n_b, n_pred), dtype=np.int32)
        for k in range(n_sub):
            p0, P, g0, G, q0 = pred_off[k], pred_off[k + 1] - pred_off[k], gt_off[k], gt_off[k + 1] - gt_off[k], pair_off[k]
            for ia, ta in enumerate(thr_a):
                for ib in range(n_b):
                    c = ia * n_b + ib
                    gm, pm = gt_m[c, g0:g0 + G], pred_m[c, p0:p0 + P]
                    for i in range(P):
                        for r in range(int(n_cand[p0 + i])):
                            j = int(order[q0 + i * G + r])
                            if mode == 0:  # compute_3d_matches, test_utils.py:597-612
                                if gm[j] > -1:
                                    continue
                                iou = table[q0 + i * G + j]
                                if iou < ta:
                                    break
                                if pred_cls[p0 + i] != gt_cls[g0 + j]:
                                    continue
                                if iou > ta:
                                    gm[j], pm[i] = i, j
                                    break
                            else:          # compute_match_from_degree_cm, test_utils.py:744-755
                                if gm[j] > -1 or pred_cls[p0 + i] != gt_cls[g0 + j]:
                                    continue
                                v = table[q0 + i * G + j]
                                if v[0] > ta or v[1] > thr_b[ib]:
                                    continue
                                gm[j], pm[i] = i, j
                                break
        return gt_m, pred_m

"""Pairwise NOCS pose metrics on the GPU and the matching that consumes them.

Stands in for ``compute_combination_3d_matches`` (core/catre/engine/test_utils.py:280-389): its two nested Python
loops over (prediction, ground truth) pairs -- ``compute_3d_iou_new`` (:140-205, up to 20 numpy box transforms per
pair) and ``compute_combination_RT_degree_cm_symmetry`` (:208-277) -- become one ``catre_pair_metrics`` launch for
every pair of every image handed in; the greedy matching loops (:354-387) stay on the host (a few dozen
comparisons per image).  Same argument meaning and return values as the reference function.

No CPU fallback: needs the CUDA library and a device.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import engine as _engine

SYM_Y_NAMES = ("bottle", "bowl", "can")      # test_utils.py:176, :246
FLIP_NAMES = ("phone", "eggbox", "glue")     # test_utils.py:257


def class_rules(synset_names: Sequence[str]) -> Tuple[int, int, int]:
    sym = sum(1 << i for i, n in enumerate(synset_names) if n in SYM_Y_NAMES)
    flip = sum(1 << i for i, n in enumerate(synset_names) if n in FLIP_NAMES)
    mug = list(synset_names).index("mug") if "mug" in synset_names else -1
    return sym, flip, mug


def pair_metrics_batch(images: List[Dict[str, np.ndarray]], synset_names: Sequence[str], device: str = "cuda"):
    """images[k] holds pred_RTs [P,4,4], pred_scales [P,3], pred_cls [P], gt_RTs [G,4,4], gt_scales [G,3], gt_cls [G],
    gt_handle [G] (numpy).  One kernel launch for all P_k x G_k pairs of all images.  Returns per image
    (overlaps [P,G] fp32, RT_overlaps [P,G,2] fp32) -- the arrays the reference fills at test_utils.py:329-352."""
    if len(synset_names) > 32:
        raise ValueError("class rules are passed as 32-bit masks: at most 32 classes")
    lib = _engine.load_library()
    if not torch.cuda.is_available():
        raise _engine.CatreError("catre_b200.metrics runs on CUDA only; there is no CPU path")
    p_off, g_off, pp, gg = [0], [0], [], []
    for im in images:
        P, G = len(im["pred_cls"]), len(im["gt_cls"])
        ii, jj = np.meshgrid(np.arange(P, dtype=np.int32), np.arange(G, dtype=np.int32), indexing="ij")
        pp.append(ii.reshape(-1) + p_off[-1])
        gg.append(jj.reshape(-1) + g_off[-1])
        p_off.append(p_off[-1] + P)
        g_off.append(g_off[-1] + G)
    n_pairs = int(sum(len(a) for a in pp))

    def cat(key, shape, dtype):
        parts = [np.asarray(im[key], dtype=dtype).reshape((-1,) + shape) for im in images]
        arr = np.concatenate(parts, axis=0) if parts else np.zeros((0,) + shape, dtype)
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device)

    out_iou = torch.empty((max(n_pairs, 1),), dtype=torch.float32, device=device)
    out_rt = torch.empty((max(n_pairs, 1), 2), dtype=torch.float32, device=device)
    if n_pairs:
        t = dict(pred_RT=cat("pred_RTs", (16,), np.float64), pred_scale=cat("pred_scales", (3,), np.float64),
                 pred_cls=cat("pred_cls", (), np.int32), gt_RT=cat("gt_RTs", (16,), np.float64),
                 gt_scale=cat("gt_scales", (3,), np.float64), gt_cls=cat("gt_cls", (), np.int32),
                 gt_handle=cat("gt_handle", (), np.int32))
        pair_p = torch.from_numpy(np.concatenate(pp).astype(np.int32)).to(device)
        pair_g = torch.from_numpy(np.concatenate(gg).astype(np.int32)).to(device)
        sym, flip, mug = class_rules(synset_names)
        stream = ctypes.c_void_p(torch.cuda.current_stream(out_iou.device).cuda_stream)
        rc = lib.catre_pair_metrics(t["pred_RT"].data_ptr(), t["pred_scale"].data_ptr(), t["pred_cls"].data_ptr(),
                                    t["gt_RT"].data_ptr(), t["gt_scale"].data_ptr(), t["gt_cls"].data_ptr(),
                                    t["gt_handle"].data_ptr(), pair_p.data_ptr(), pair_g.data_ptr(), n_pairs, sym, flip, mug,
                                    out_iou.data_ptr(), out_rt.data_ptr(), stream)
        if rc != 0:
            raise _engine.CatreError(f"catre_pair_metrics failed ({rc}): {lib.catre_last_error(None).decode()}")
    iou_h, rt_h = out_iou.cpu().numpy(), out_rt.cpu().numpy()
    res, o = [], 0
    for k, im in enumerate(images):
        P, G = len(im["pred_cls"]), len(im["gt_cls"])
        res.append((iou_h[o:o + P * G].reshape(P, G).copy(), rt_h[o:o + P * G].reshape(P, G, 2).copy()))
        o += P * G
    return res


def greedy_matches(overlaps, rt_overlaps, pred_class_ids, gt_class_ids, iou_3d_thresholds, degree_thesholds, shift_thesholds,
                   score_threshold=0):
    """The matching loops of the reference (test_utils.py:354-387) over score-sorted predictions."""
    num_pred, num_gt = overlaps.shape
    nd, nt, ns = len(degree_thesholds), len(shift_thesholds), len(iou_3d_thresholds)
    pred_matches = -1 * np.ones([nd, nt, ns, num_pred])
    gt_matches = -1 * np.ones([nd, nt, ns, num_gt])
    order_all = [np.argsort(overlaps[i])[::-1] for i in range(num_pred)]
    for s, iou_thres in enumerate(iou_3d_thresholds):
        for d, degree_thres in enumerate(degree_thesholds):
            for t, shift_thres in enumerate(shift_thesholds):
                for i in range(num_pred):
                    order = order_all[i]
                    low = np.where(overlaps[i, order] < score_threshold)[0]
                    if low.size > 0:
                        order = order[: low[0]]
                    for j in order:
                        if gt_matches[d, t, s, j] > -1:
                            continue
                        if overlaps[i, j] < iou_thres or rt_overlaps[i, j, 0] > degree_thres or rt_overlaps[i, j, 1] > shift_thres:
                            break  # sorted by IoU: nothing better follows
                        if not pred_class_ids[i] == gt_class_ids[j]:
                            continue
                        gt_matches[d, t, s, j] = i
                        pred_matches[d, t, s, i] = j
                        break
    return gt_matches, pred_matches


def _sorted_image(gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, pred_class_ids, pred_scores, pred_RTs, pred_scales):
    num_pred = len(pred_class_ids)
    indices = np.zeros(0)
    pred_class_ids, pred_RTs, pred_scales = np.asarray(pred_class_ids), np.asarray(pred_RTs), np.asarray(pred_scales)
    if num_pred:
        indices = np.argsort(np.asarray(pred_scores))[::-1]  # predictions by score, high to low (test_utils.py:318-325)
        pred_class_ids, pred_RTs, pred_scales = pred_class_ids[indices], pred_RTs[indices], pred_scales[indices]
    im = dict(pred_RTs=pred_RTs.reshape(-1, 4, 4), pred_scales=pred_scales.reshape(-1, 3), pred_cls=pred_class_ids,
              gt_RTs=np.asarray(gt_RTs).reshape(-1, 4, 4), gt_scales=np.asarray(gt_scales).reshape(-1, 3),
              gt_cls=np.asarray(gt_class_ids), gt_handle=np.asarray(gt_handle_visibility))
    return im, indices


def compute_combination_3d_matches(gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, synset_names, pred_boxes,
                                   pred_class_ids, pred_scores, pred_RTs, pred_scales, iou_3d_thresholds, degree_thesholds,
                                   shift_thesholds, score_threshold=0):
    """Same contract as the reference function (test_utils.py:280-389): returns (gt_matches [D,T,S,G],
    pred_matches [D,T,S,P], indices = the score order applied to the predictions).  ``pred_boxes`` is accepted for
    signature compatibility (the reference only uses it to trim zero padding)."""
    im, indices = _sorted_image(gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, pred_class_ids, pred_scores, pred_RTs,
                                pred_scales)
    overlaps, rt = pair_metrics_batch([im], synset_names)[0]
    gt_m, pred_m = greedy_matches(overlaps, rt, im["pred_cls"], im["gt_cls"], iou_3d_thresholds, degree_thesholds,
                                  shift_thesholds, score_threshold)
    return gt_m, pred_m, indices


def match_images(results: List[Dict[str, np.ndarray]], synset_names, iou_3d_thresholds, degree_thesholds, shift_thesholds):
    """Batched form: ``results[k]`` has the keys of one entry of the reference's ``final_results`` that the matcher
    reads (gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, pred_class_ids, pred_scores, pred_RTs, pred_scales).
    One GPU launch for all images, then the per-image host matching.  Returns [(gt_matches, pred_matches, indices)]."""
    ims, idxs = [], []
    for r in results:
        im, ind = _sorted_image(r["gt_class_ids"], r["gt_RTs"], r["gt_scales"], r["gt_handle_visibility"], r["pred_class_ids"],
                                r["pred_scores"], r["pred_RTs"], r["pred_scales"])
        ims.append(im)
        idxs.append(ind)
    pm = pair_metrics_batch(ims, synset_names)
    out = []
    for im, ind, (ov, rt) in zip(ims, idxs, pm):
        gt_m, pred_m = greedy_matches(ov, rt, im["pred_cls"], im["gt_cls"], iou_3d_thresholds, degree_thesholds, shift_thesholds)
        out.append((gt_m, pred_m, ind))
    return out

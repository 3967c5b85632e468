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


def pair_metrics_batch(images: List[Dict[str, np.ndarray]], synset_names: Sequence[str], device: str = "cuda",
                       shift_cm: bool = False):
    """images[k] holds pred_RTs [P,4,4], pred_scales [P,3], pred_cls [P], gt_RTs [G,4,4], gt_scales [G,3], gt_cls [G],
    gt_handle [G] (numpy).  One kernel launch for all P_k x G_k pairs of all images.  Returns per image
    (overlaps [P,G] fp32, RT_overlaps [P,G,2]) -- the arrays the reference fills at test_utils.py:329-352 (shift =
    |dT| / scale, stored fp32) or, with ``shift_cm``, at test_utils.py:568-578 and 703-710 (compute_3d_matches'
    overlaps and compute_RT_overlaps' fp64 [degree, |dT| * 100 cm])."""
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
    out_rt = torch.empty((max(n_pairs, 1), 2), dtype=torch.float64 if shift_cm else torch.float32, device=device)
    if n_pairs:
        t = dict(pred_RT=cat("pred_RTs", (16,), np.float64), pred_scale=cat("pred_scales", (3,), np.float64),
                 pred_cls=cat("pred_cls", (), np.int32), gt_RT=cat("gt_RTs", (16,), np.float64),
                 gt_scale=cat("gt_scales", (3,), np.float64), gt_cls=cat("gt_cls", (), np.int32),
                 gt_handle=cat("gt_handle", (), np.int32))
        pair_p = torch.from_numpy(np.concatenate(pp).astype(np.int32)).to(device)
        pair_g = torch.from_numpy(np.concatenate(gg).astype(np.int32)).to(device)
        sym, flip, mug = class_rules(synset_names)
        stream = ctypes.c_void_p(torch.cuda.current_stream(out_iou.device).cuda_stream)
        rc = lib.catre_pair_metrics_ex(t["pred_RT"].data_ptr(), t["pred_scale"].data_ptr(), t["pred_cls"].data_ptr(),
                                       t["gt_RT"].data_ptr(), t["gt_scale"].data_ptr(), t["gt_cls"].data_ptr(),
                                       t["gt_handle"].data_ptr(), pair_p.data_ptr(), pair_g.data_ptr(), n_pairs, sym, flip, mug,
                                       1 if shift_cm else 0, out_iou.data_ptr(), None if shift_cm else out_rt.data_ptr(),
                                       out_rt.data_ptr() if shift_cm else None, stream)
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


def greedy_matches_batch(items, iou_3d_thresholds, degree_thesholds, shift_thesholds, score_threshold=0):
    """greedy_matches for many images at once.  ``items`` = [(overlaps [P,G], rt_overlaps [P,G,2], pred_class_ids [P],
    gt_class_ids [G])] over score-sorted predictions; returns [(gt_matches [D,T,S,G], pred_matches [D,T,S,P])], equal to
    the per-image loops (test_utils.py:354-387) entry for entry.

    The loops are sequential in the prediction (score order) and in its candidate list (IoU order) but independent across
    images and across the D*T*S threshold triples, so every step runs as one array operation over [images, D, T, S]:
    a triple of an image is `alive` while its current prediction is still looking for a ground truth; a candidate that is
    already taken is skipped, one that fails a threshold ends the search (`break`), one of another class is skipped, the
    first that passes is matched.  The candidate order of every row is numpy's argsort of that row, as in the reference
    (ties keep numpy's order)."""
    n = len(items)
    nd, nt, ns = len(degree_thesholds), len(shift_thesholds), len(iou_3d_thresholds)
    if n == 0:
        return []
    p_cnt = np.array([it[0].shape[0] for it in items])
    g_cnt = np.array([it[0].shape[1] for it in items])
    pmax, gmax = int(p_cnt.max()), int(g_cnt.max())
    out_shape = (n, nd, nt, ns)
    gt_m = -1 * np.ones(out_shape + (max(gmax, 1),))
    pred_m = -1 * np.ones(out_shape + (max(pmax, 1),))
    if pmax and gmax:
        ov = np.full((n, pmax, gmax), -np.inf)
        rt = np.full((n, pmax, gmax, 2), np.inf)
        order = np.zeros((n, pmax, gmax), dtype=np.int64)
        n_cand = np.zeros((n, pmax), dtype=np.int64)  # candidates of a prediction after the score_threshold cut
        pcls = np.full((n, pmax), -1, dtype=np.int64)
        gcls = np.full((n, gmax), -2, dtype=np.int64)
        for k, (o, r, pc, gc) in enumerate(items):
            P, G = o.shape
            if P == 0 or G == 0:
                continue
            ov[k, :P, :G], rt[k, :P, :G] = o, r
            pcls[k, :P], gcls[k, :G] = pc, gc
            for i in range(P):
                od = np.argsort(o[i])[::-1]
                low = np.where(o[i, od] < score_threshold)[0]
                order[k, i, :G] = od
                n_cand[k, i] = low[0] if low.size else G
        iou_t = np.asarray(iou_3d_thresholds, dtype=np.float64).reshape(1, 1, 1, ns)
        deg_t = np.asarray(degree_thesholds, dtype=np.float64).reshape(1, nd, 1, 1)
        sh_t = np.asarray(shift_thesholds, dtype=np.float64).reshape(1, 1, nt, 1)
        rows = np.arange(n)
        for i in range(pmax):
            alive = np.broadcast_to((n_cand[:, i] > 0).reshape(n, 1, 1, 1), out_shape).copy()
            for r in range(gmax):
                alive &= (r < n_cand[:, i]).reshape(n, 1, 1, 1)
                if not alive.any():
                    break
                j = order[:, i, r]
                free = np.take_along_axis(gt_m, j.reshape(n, 1, 1, 1, 1), axis=4)[..., 0] <= -1
                o_ij, r_ij = ov[rows, i, j], rt[rows, i, j]
                fail = (o_ij.reshape(n, 1, 1, 1) < iou_t) | (r_ij[:, 0].reshape(n, 1, 1, 1) > deg_t) | (r_ij[:, 1].reshape(n, 1, 1, 1) > sh_t)
                consider = alive & free
                alive &= ~(consider & fail)
                hit = consider & ~fail & (pcls[:, i] == gcls[rows, j]).reshape(n, 1, 1, 1)
                if hit.any():
                    kk, dd, tt, ss = np.nonzero(hit)
                    gt_m[kk, dd, tt, ss, j[kk]] = i
                    pred_m[kk, dd, tt, ss, i] = j[kk]
                    alive &= ~hit
    return [(gt_m[k, ..., :g_cnt[k]].copy(), pred_m[k, ..., :p_cnt[k]].copy()) for k in range(n)]


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


def match_images(results: List[Dict[str, np.ndarray]], synset_names, iou_3d_thresholds, degree_thesholds, shift_thesholds,
                 pair_metrics_fn=None):
    """Batched form: ``results[k]`` has the keys of one entry of the reference's ``final_results`` that the matcher
    reads (gt_class_ids, gt_RTs, gt_scales, gt_handle_visibility, pred_class_ids, pred_scores, pred_RTs, pred_scales).
    One GPU launch for all images, then the matching of all images at once (greedy_matches_batch).  Returns [(gt_matches, pred_matches, indices)]."""
    ims, idxs = [], []
    for r in results:
        im, ind = _sorted_image(r["gt_class_ids"], r["gt_RTs"], r["gt_scales"], r["gt_handle_visibility"], r["pred_class_ids"],
                                r["pred_scores"], r["pred_RTs"], r["pred_scales"])
        ims.append(im)
        idxs.append(ind)
    pm = (pair_metrics_fn or pair_metrics_batch)(ims, synset_names)
    matched = greedy_matches_batch([(ov, rt, im["pred_cls"], im["gt_cls"]) for im, (ov, rt) in zip(ims, pm)], iou_3d_thresholds,
                                   degree_thesholds, shift_thesholds)
    return [(gt_m, pred_m, ind) for (gt_m, pred_m), ind in zip(matched, idxs)]


def compute_ap_from_matches_scores(pred_match, pred_scores, gt_match) -> float:
    """VOC-style AP of one class / threshold triple (test_utils.py:112-137): predictions by descending score,
    precision made monotone from the right, summed over the recall steps."""
    assert pred_match.shape[0] == pred_scores.shape[0]
    order = np.argsort(pred_scores)[::-1]
    hit = pred_match[order] > -1
    precisions = np.cumsum(hit) / (np.arange(len(hit)) + 1)
    recalls = np.cumsum(hit).astype(np.float32) / len(gt_match)
    precisions = np.concatenate([[0], precisions, [0]])
    recalls = np.concatenate([[0], recalls, [1]])
    precisions = np.maximum.accumulate(precisions[::-1])[::-1]  # the reference's right-to-left running maximum (:127-128)
    steps = np.where(recalls[:-1] != recalls[1:])[0] + 1
    return np.sum((recalls[steps] - recalls[steps - 1]) * precisions[steps])


def compute_combination_mAP(final_results, synset_names=("BG", "bottle", "bowl", "camera", "can", "laptop", "mug"),
                            degree_thresholds=(5, 10, 15), shift_thresholds=(0.1, 0.2), iou_3d_thresholds=(0.1,),
                            pair_metrics_fn=None):
    """Same contract as the reference's compute_combination_mAP (test_utils.py:392-520), without its printing:
    returns aps [num_classes + 1, len(degree)+1, len(shift)+1, len(iou)] (last class row = mean over classes
    1..num_classes-1).  The reference calls the matcher once per (image, class) from Python; here all those
    sub-problems go through ONE pair-metrics launch (``pair_metrics_fn`` is the pair stage, the CUDA one by
    default; tests inject the CPU oracle to check the host logic without a GPU)."""
    synset_names = list(synset_names)
    num_classes = len(synset_names)
    deg_list = list(degree_thresholds) + [360]
    shift_list = list(shift_thresholds) + [100]
    iou_list = list(iou_3d_thresholds)
    nd, nt, ns = len(deg_list), len(shift_list), len(iou_list)
    subs, sub_cls = [], []
    for result in final_results:
        gt_class_ids = np.asarray(result["gt_class_ids"]).astype(np.int32)
        gt_RTs, gt_scales = np.array(result["gt_RTs"]), np.array(result["gt_scales"])
        gt_handle = np.asarray(result["gt_handle_visibility"])
        pred_class_ids, pred_scales = np.asarray(result["pred_class_ids"]), np.asarray(result["pred_scales"])
        pred_scores, pred_RTs = np.asarray(result["pred_scores"]), np.array(result["pred_RTs"])
        if len(gt_class_ids) == 0 and len(pred_class_ids) == 0:
            continue
        for cls_id in range(1, num_classes):  # only same-class predictions / ground truths meet (test_utils.py:431-441)
            g = gt_class_ids == cls_id if len(gt_class_ids) else np.zeros(0, bool)
            q = pred_class_ids == cls_id if len(pred_class_ids) else np.zeros(0, bool)
            n_g = int(g.sum())
            if n_g == 0 and not q.any():
                continue  # neither a ground truth nor a prediction of this class: the reference's call returns empty arrays
            if synset_names[cls_id] != "mug":
                handle = np.ones(n_g, dtype=np.int32)  # handle visibility only matters for mugs (:443-448)
            else:
                handle = gt_handle[g] if len(gt_class_ids) else np.ones(0)
            subs.append(dict(gt_class_ids=gt_class_ids[g] if len(gt_class_ids) else np.zeros(0, np.int32),
                             gt_RTs=gt_RTs[g] if len(gt_class_ids) else np.zeros((0, 4, 4)),
                             gt_scales=gt_scales[g] if len(gt_class_ids) else np.zeros((0, 3)), gt_handle_visibility=handle,
                             pred_class_ids=pred_class_ids[q] if len(pred_class_ids) else np.zeros(0, np.int32),
                             pred_scores=pred_scores[q] if len(pred_class_ids) else np.zeros(0),
                             pred_RTs=pred_RTs[q] if len(pred_class_ids) else np.zeros((0, 4, 4)),
                             pred_scales=pred_scales[q] if len(pred_class_ids) else np.zeros((0, 3))))
            sub_cls.append(cls_id)
    matched = match_images(subs, synset_names, iou_list, deg_list, shift_list, pair_metrics_fn=pair_metrics_fn)
    pred_m = [[np.zeros((nd, nt, ns, 0))] for _ in range(num_classes)]
    gt_m = [[np.zeros((nd, nt, ns, 0))] for _ in range(num_classes)]
    scores = [[np.zeros(0)] for _ in range(num_classes)]
    for sub, cls_id, (gm, pm, ind) in zip(subs, sub_cls, matched):
        sc = np.asarray(sub["pred_scores"])
        if len(ind):
            sc = sc[ind]
        pred_m[cls_id].append(pm)
        scores[cls_id].append(sc)  # the same for every threshold triple: broadcast once per class below
        gt_m[cls_id].append(gm)
    aps = np.zeros((num_classes + 1, nd, nt, ns))
    for cls_id in range(1, num_classes):
        pm_all, gm_all = (np.concatenate(x[cls_id], axis=-1) for x in (pred_m, gt_m))
        sc_all = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in scores[cls_id]])  # np.tile(...) made float64 rows
        for s in range(ns):
            for d in range(nd):
                for t in range(nt):
                    aps[cls_id, d, t, s] = compute_ap_from_matches_scores(pm_all[d, t, s, :], sc_all, gm_all[d, t, s, :])
    aps[-1] = np.mean(aps[1:-1], axis=0)
    return aps

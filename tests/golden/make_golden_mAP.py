"""Golden vectors for the metric the NOCS evaluator actually computes (VERDICT r1 missing #1): the UNMODIFIED reference
functions compute_3d_matches, compute_RT_overlaps, compute_match_from_degree_cm and compute_independent_mAP
(core/catre/engine/test_utils.py:523-926), with the thresholds CATRE_EvaluatorCustom passes
(core/catre/engine/catre_custom_evaluator.py:247-260).  Runs only in the build container.

Synthetic "images" as in make_golden_metrics.py, but with several instances of the SAME class per image (so the greedy
matching has real choices to make), translation noise on the centimetre scale of the pose thresholds, metric box sizes
with pure-rotation RTs (what the evaluator builds from the model's pose / scale output), plus an image without
predictions, one without ground truth and one with neither.

Usage: python tests/golden/make_golden_mAP.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_shim  # noqa: E402
from make_golden_metrics import rand_rot, small_rot  # noqa: E402

SYNSET = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]
DEG, SHIFT, IOU = [5, 10], [2, 5, 10], [0.1, 0.25, 0.50, 0.75]  # catre_custom_evaluator.py:247-251
N_IMG = 80


def make_image(g, kind):
    n_gt = 0 if kind in ("no_gt", "empty") else g.randint(2, 8)
    pool = g.choice(np.arange(1, 7), size=g.randint(1, 4), replace=False)  # few classes per image -> same-class groups
    gt_cls = g.choice(pool, size=n_gt) if n_gt else np.zeros(0, np.int32)
    gt_RTs, gt_scales, handle = [], [], []
    for _ in range(n_gt):
        RT = np.eye(4)
        RT[:3, :3] = rand_rot(g)
        RT[:3, 3] = g.uniform(-0.25, 0.25, size=3) + np.array([0, 0, 0.9])
        gt_RTs.append(RT)
        gt_scales.append(g.uniform(0.06, 0.3, size=3))
        handle.append(g.randint(0, 2))
    pred_RTs, pred_scales, pred_cls, pred_scores, boxes = [], [], [], [], []
    if kind not in ("no_pred", "empty"):
        for j in range(n_gt):
            if g.uniform() < 0.12:
                continue  # missed detection
            deg, cm, ds = [(1.5, 0.4, 0.02), (4, 1.5, 0.05), (9, 4, 0.1), (40, 15, 0.3)][g.randint(0, 4)]
            RT = np.eye(4)
            RT[:3, :3] = small_rot(g, deg) @ gt_RTs[j][:3, :3]
            RT[:3, 3] = gt_RTs[j][:3, 3] + 0.01 * cm * g.normal(size=3)
            pred_RTs.append(RT)
            pred_scales.append(gt_scales[j] * (1 + ds * g.normal(size=3)))
            pred_cls.append(gt_cls[j] if g.uniform() > 0.08 else g.randint(1, 7))
            pred_scores.append(g.uniform(0.1, 1.0))
        for _ in range(g.randint(0, 3)):  # false positives, sometimes a near duplicate of an earlier prediction
            if pred_RTs and g.uniform() < 0.5:
                k = g.randint(0, len(pred_RTs))
                RT = pred_RTs[k].copy()
                RT[:3, 3] += 0.01 * g.normal(size=3)
                pred_RTs.append(RT); pred_scales.append(pred_scales[k].copy()); pred_cls.append(pred_cls[k])
            else:
                RT = np.eye(4)
                RT[:3, :3] = rand_rot(g)
                RT[:3, 3] = g.uniform(-0.25, 0.25, size=3) + np.array([0, 0, 0.9])
                pred_RTs.append(RT); pred_scales.append(g.uniform(0.06, 0.3, size=3)); pred_cls.append(g.randint(1, 7))
            pred_scores.append(g.uniform(0.1, 1.0))
    n_pred = len(pred_cls)
    boxes = np.stack([g.randint(1, 200, size=n_pred), g.randint(1, 200, size=n_pred), g.randint(200, 400, size=n_pred),
                      g.randint(200, 400, size=n_pred)], axis=1) if n_pred else np.zeros((0, 4), np.int64)
    return dict(gt_class_ids=np.asarray(gt_cls, np.int32), gt_RTs=np.asarray(gt_RTs).reshape(n_gt, 4, 4),
                gt_scales=np.asarray(gt_scales).reshape(n_gt, 3), gt_handle_visibility=np.asarray(handle, np.int32),
                pred_class_ids=np.asarray(pred_cls, np.int32), pred_RTs=np.asarray(pred_RTs).reshape(n_pred, 4, 4),
                pred_scales=np.asarray(pred_scales).reshape(n_pred, 3), pred_scores=np.asarray(pred_scores, np.float64),
                pred_bboxes=boxes)


def main():
    install_shim("/root/reference")
    from core.catre.engine import test_utils as tu  # the reference's own functions

    g = np.random.RandomState(7)
    kinds = ["full"] * (N_IMG - 3) + ["no_pred", "no_gt", "empty"]
    g.shuffle(kinds)
    results = [make_image(g, k) for k in kinds]
    out = {"n_img": np.int64(N_IMG)}
    for k, r in enumerate(results):
        for name, v in r.items():
            out[f"{name}_{k}"] = v
    with np.errstate(invalid="ignore", divide="ignore"):
        iou_3d_aps, pose_aps = tu.compute_independent_mAP(results, SYNSET, degree_thresholds=DEG, shift_thresholds=SHIFT,
                                                          iou_3d_thresholds=IOU)
        out["iou_3d_aps"], out["pose_aps"] = iou_3d_aps, pose_aps
        # the same with the evaluator-independent defaults and without the IoU pre-matching of the pose metric
        a2, p2 = tu.compute_independent_mAP(results, SYNSET, degree_thresholds=DEG, shift_thresholds=SHIFT,
                                            iou_3d_thresholds=IOU, use_matches_for_pose=False)
        out["iou_3d_aps_nomatch"], out["pose_aps_nomatch"] = a2, p2
        # per-function goldens on whole images (mixed classes inside one call)
        n_fn = 0
        for k, r in enumerate(results):
            if len(r["pred_class_ids"]) == 0 or len(r["gt_class_ids"]) == 0 or n_fn >= 30:
                continue
            gm, pm, ov, idx = tu.compute_3d_matches(r["gt_class_ids"], r["gt_RTs"], r["gt_scales"], r["gt_handle_visibility"], SYNSET,
                                                    r["pred_bboxes"], r["pred_class_ids"], r["pred_scores"], r["pred_RTs"],
                                                    r["pred_scales"], IOU)
            rt = tu.compute_RT_overlaps(r["gt_class_ids"], r["gt_RTs"], r["gt_handle_visibility"], r["pred_class_ids"], r["pred_RTs"],
                                        SYNSET)
            gm2, pm2 = tu.compute_match_from_degree_cm(rt, r["pred_class_ids"], r["gt_class_ids"], DEG + [360], SHIFT + [100])
            out[f"fn_gt_matches_{k}"], out[f"fn_pred_matches_{k}"], out[f"fn_overlaps_{k}"] = gm, pm, ov
            out[f"fn_indices_{k}"], out[f"fn_rt_{k}"] = np.asarray(idx, np.int64), rt
            out[f"fn_pose_gt_matches_{k}"], out[f"fn_pose_pred_matches_{k}"] = gm2, pm2
            n_fn += 1
    np.savez_compressed(os.path.join(HERE, "golden_mAP.npz"), **out)
    print("iou_3d_aps (mean row)", np.round(iou_3d_aps[-1], 4), "\npose_aps (mean)\n", np.round(pose_aps[-1], 4))
    print("wrote golden_mAP.npz", os.path.getsize(os.path.join(HERE, "golden_mAP.npz")), "bytes;", n_fn, "per-function images")


if __name__ == "__main__":
    main()

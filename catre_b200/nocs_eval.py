"""NOCS-format result collection and evaluation (SURVEY.md 8(f) N1 + N3, VERDICT r1 missing #2).

Stands in for ``CATRE_EvaluatorCustom`` (core/catre/engine/catre_custom_evaluator.py:33-330) -- the evaluator the
reference builds for NOCS (``evaluator_type="nocs"``, core/catre/engine/engine.py:87-95):

  process(inputs, batch, outputs, out_dict)   :121-176  one record per object and refinement iteration:
                                                        pred_RTs 4x4, pred_scales, pred_class_ids (+1), pred_scores,
                                                        pred_bboxes (yxyx), keyed by scene_im_id
  evaluate()                                  :200-213  gather over ranks, regroup per iteration and image (:184-198)
  _eval_predictions(cur_iter)                 :215-330  merge with the ground truth (get_gts :83-105), call
                                                        compute_independent_mAP with the evaluator's thresholds, format
                                                        the IoU25/50/75, re5te2 ... te5 table

What is different here: records are numeric rows (not Python dicts), so the multi-rank collection is ONE padded tensor
all-gather (catre_b200.shard.gather_rows) instead of the reference's pickled-object ``all_gather`` (:202-203); the
metric runs on the GPU through catre_b200.nocs_map (pair kernel + device matching) instead of nested Python loops; and
``evaluate()`` RETURNS what the reference only logs / writes: {"iter{i}": {"iou_3d_aps", "pose_aps", "table"}} (the
reference returns {}).  The table text and the ``*_tab_iter{i}.txt`` files are the reference's, character for character
(tests/test_nocs_eval.py).  The ground truth comes in as the reference's ``dataset_dicts`` (detectron2 DatasetCatalog
format: scene_im_id, file_name, annotations with category_id / bbox / pose / scale / mug_handle).
"""
from __future__ import annotations

import logging
import os
import pickle
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import nocs_map as _map
from . import shard as _shard

logger = logging.getLogger(__name__)

# the evaluator's thresholds (catre_custom_evaluator.py:247-251)
DEGREE_THRESHOLDS = [5, 10]
SHIFT_THRESHOLDS = [2, 5, 10]
DEGREE_SHIFT_THRESHOLDS = [(5, 2), (5, 5), (10, 2), (10, 5), (10, 10)]
IOU_3D_THRESHOLDS = [0.1, 0.25, 0.50, 0.75]


def bbox_xyxy_to_yxyx(bbox) -> List[int]:
    """test_utils.py:19-22"""
    x1, y1, x2, y2 = bbox
    return [int(y1), int(x1), int(y2), int(x2)]


def pose_3x4_to_4x4(pose) -> np.ndarray:
    """test_utils.py:25-28"""
    return np.concatenate((pose, np.array([[0, 0, 0, 1]], dtype=np.float32)), axis=0)


def build_gt_dict(dataset_dicts: Sequence[Dict[str, Any]]) -> "OrderedDict[str, Dict[str, Any]]":
    """get_gts (catre_custom_evaluator.py:83-105): per scene_im_id the ground-truth arrays compute_independent_mAP reads;
    class ids start from 1; entries that repeat a scene_im_id are concatenated."""
    gts: "OrderedDict[str, Dict[str, Any]]" = OrderedDict()
    for im in dataset_dicts:
        annos = im["annotations"]
        gt = dict(gt_class_ids=np.array([a["category_id"] + 1 for a in annos]),
                  gt_bboxes=np.array([bbox_xyxy_to_yxyx(a["bbox"]) for a in annos]),
                  gt_RTs=np.array([pose_3x4_to_4x4(a["pose"]) for a in annos]),
                  gt_scales=np.array([a["scale"] for a in annos]),
                  gt_handle_visibility=np.array([a["mug_handle"] for a in annos]))
        key = im["scene_im_id"]
        if key not in gts:
            gts[key] = gt
            gts[key]["image_path"] = [im["file_name"]]
        else:
            gts[key]["image_path"].append(im["file_name"])
            for k, v in gt.items():
                gts[key][k] = np.concatenate((gts[key][k], v), axis=0)
    return gts


def format_table(iou_3d_aps: np.ndarray, pose_aps: np.ndarray, obj_names: Sequence[str]) -> str:
    """The table _eval_predictions logs and writes (catre_custom_evaluator.py:262-318), same rows, same formatting."""
    from tabulate import tabulate

    obj_names = list(obj_names)
    synset_names = ["BG"] + obj_names
    cls_rows = [i for i, n in enumerate(synset_names) if n in obj_names]
    big_tab = [["objects"] + obj_names + [f"Avg({len(obj_names)})"]]

    def line(name, pick):
        return [name] + [f"{100 * pick(i):.2f}" for i in cls_rows] + [f"{100 * pick(-1):.2f}"]

    for metric, thres in zip(["IoU25", "IoU50", "IoU75"], IOU_3D_THRESHOLDS[1:]):
        s = IOU_3D_THRESHOLDS.index(thres)
        big_tab.append(line(metric, lambda i, s=s: iou_3d_aps[i, s]))
    for metric, (deg, sh) in zip(["re5te2", "re5te5", "re10te2", "re10te5", "re10te10"], DEGREE_SHIFT_THRESHOLDS):
        d, t = DEGREE_THRESHOLDS.index(deg), SHIFT_THRESHOLDS.index(sh)
        big_tab.append(line(metric, lambda i, d=d, t=t: pose_aps[i, d, t]))
    for metric, deg in zip(["re5", "re10"], DEGREE_THRESHOLDS):
        d = DEGREE_THRESHOLDS.index(deg)
        big_tab.append(line(metric, lambda i, d=d: pose_aps[i, d, -1]))
    for metric, sh in zip(["te2", "te5"], SHIFT_THRESHOLDS):  # zip stops after two, as in the reference (:304)
        t = SHIFT_THRESHOLDS.index(sh)
        big_tab.append(line(metric, lambda i, t=t: pose_aps[i, -1, t]))
    return tabulate(big_tab, tablefmt="plain")


class NocsPredictionCollector:
    """process() / evaluate() of CATRE_EvaluatorCustom with tensor rows instead of a list of dicts."""

    _META = 7  # scene_im_id index, class id (1-based), score, has_score flag, bbox y1 x1 y2 x2 -> 4 (total 8 below)

    def __init__(self, obj_names: Sequence[str], n_iter_test: int, dataset_dicts: Optional[Sequence[Dict[str, Any]]] = None,
                 train_objs: Optional[Sequence[str]] = None, distributed: bool = False, gather_device: str = "cpu",
                 output_dir: Optional[str] = None, exp_id: str = "catre_b200", dataset_name: str = "nocs", map_backend=None):
        self.obj_names = list(obj_names)
        self.n_iter_test = int(n_iter_test)
        self.train_objs = list(train_objs) if train_objs is not None else None
        self.dataset_dicts = dataset_dicts
        self._distributed = distributed
        self._gather_device = gather_device  # "cuda" under nccl, "cpu" under gloo
        self._output_dir = output_dir
        self._exp_id, self.dataset_name = exp_id, dataset_name
        self._map_backend = map_backend
        self.reset()

    # ---- the reference's evaluator surface --------------------------------------------------------------------------
    def reset(self):
        self._meta: List[List[float]] = []          # one row per object: [scene idx, class id, score, has_score, bbox yxyx]
        self._order: List[int] = []                 # object -> row of the concatenated pose tensors
        k1 = self.n_iter_test + 1
        self._poses: List[List[torch.Tensor]] = [[] for _ in range(k1)]
        self._scales: List[List[torch.Tensor]] = [[] for _ in range(k1)]
        self._n_seen = 0
        self._scene_ims: List[str] = []
        self._scene_index: Dict[str, int] = {}
        self._predictions_dict: "OrderedDict[str, Dict[str, Dict[str, np.ndarray]]]" = OrderedDict()

    def _maybe_adapt_label_cls_name(self, label):
        """catre_custom_evaluator.py:108-119"""
        name = self.obj_names[label]
        if self.train_objs is None:
            return label, name
        if name not in self.train_objs:
            return None, None
        return self.train_objs.index(name), name

    def process(self, inputs, batch, outputs, out_dict):
        """catre_custom_evaluator.py:121-176: bookkeeping only (this runs underneath the next launch); the numeric rows are
        assembled once, in rows()."""
        im_ids = batch["im_id"].detach().cpu().tolist()
        inst_ids = batch["inst_id"].detach().cpu().tolist()
        labels = batch["obj_cls"].detach().cpu().tolist()
        for im_i, (inp, _output) in enumerate(zip(inputs, outputs)):
            key = inp["scene_im_id"]
            if key not in self._scene_index:
                self._scene_index[key] = len(self._scene_ims)
                self._scene_ims.append(key)
            inst = inp.get("instances", None)
            for out_i, b_im in enumerate(im_ids):
                if int(b_im) != im_i:
                    continue
                inst_id = int(inst_ids[out_i])
                bbox, score, has_score = [0, 0, 0, 0], 1.0, 0.0
                if inst is not None:
                    boxes = inst.obj_boxes.tensor if hasattr(inst.obj_boxes, "tensor") else inst.obj_boxes
                    bbox = bbox_xyxy_to_yxyx(boxes[inst_id])
                    if (inst.has("obj_scores") if hasattr(inst, "has") else hasattr(inst, "obj_scores")):
                        score, has_score = float(inst.obj_scores[inst_id]), 1.0
                self._meta.append([self._scene_index[key], labels[out_i] + 1, score, has_score] + [float(v) for v in bbox])
                self._order.append(self._n_seen + out_i)
        self._n_seen += len(labels)
        for i in range(self.n_iter_test + 1):
            self._poses[i].append(out_dict[f"pose_{i}"])
            self._scales[i].append(out_dict[f"scale_{i}"])

    # ---- rows <-> the reference's per-iteration prediction dicts -----------------------------------------------------
    def rows(self) -> torch.Tensor:
        """[objects, 8 + (K+1) * 15] float64: metadata, then per iteration the 3x4 pose and the 3 scales."""
        k1 = self.n_iter_test + 1
        width = 8 + k1 * 15
        if not self._order:
            return torch.zeros((0, width), dtype=torch.float64)
        poses = torch.stack([torch.cat([t.detach().cpu() for t in self._poses[i]], dim=0) for i in range(k1)], dim=1).double()
        scales = torch.stack([torch.cat([t.detach().cpu() for t in self._scales[i]], dim=0) for i in range(k1)], dim=1).double()
        n = poses.shape[0]
        body = torch.cat((poses.reshape(n, k1, 12), scales.reshape(n, k1, 3)), dim=2).reshape(n, k1 * 15)
        idx = torch.tensor(self._order, dtype=torch.long)
        return torch.cat((torch.tensor(self._meta, dtype=torch.float64), body[idx]), dim=1)

    def _preds_from_rows(self, rows: torch.Tensor, scene_ims: Sequence[str]) -> None:
        """_preds_list_to_dict + batch_prediction_results (catre_custom_evaluator.py:178-198): {iter{i}: {scene_im_id: arrays}}
        with the reference's dtypes (fp32 poses / scales from the fp32 model output, int64 ids and boxes)."""
        k1 = self.n_iter_test + 1
        r = rows.numpy()
        body = r[:, 8:].reshape(-1, k1, 15)
        self._predictions_dict = OrderedDict((f"iter{i}", OrderedDict()) for i in range(k1))
        groups: "OrderedDict[int, List[int]]" = OrderedDict()
        for j, s in enumerate(r[:, 0].astype(np.int64).tolist()):
            groups.setdefault(s, []).append(j)
        bottom = np.array([0, 0, 0, 1], dtype=np.float32)
        for s, js in groups.items():
            js = np.asarray(js)
            has_score = bool(r[js, 3].all())
            common = dict(pred_class_ids=r[js, 1].astype(np.int64), pred_scores=r[js, 2].astype(np.float32 if has_score else np.float64),
                          pred_bboxes=r[js, 4:8].astype(np.int64))
            for i in range(k1):
                rt = np.zeros((len(js), 4, 4), dtype=np.float32)
                rt[:, :3, :] = body[js, i, :12].reshape(-1, 3, 4).astype(np.float32)
                rt[:, 3, :] = bottom
                self._predictions_dict[f"iter{i}"][scene_ims[s]] = dict(pred_RTs=rt, pred_scales=body[js, i, 12:].astype(np.float32), **common)

    def evaluate(self):
        """catre_custom_evaluator.py:200-213; returns None on the non-main ranks like the reference."""
        rows, scene_ims = self.rows(), list(self._scene_ims)
        if self._distributed and _shard.world_size() > 1:
            vocab = _shard.gather_vocab(scene_ims)
            if rows.shape[0]:  # re-index this rank's image keys into the global vocabulary
                remap = torch.tensor([vocab.index(s) for s in scene_ims], dtype=torch.float64)
                rows = rows.clone()
                rows[:, 0] = remap[rows[:, 0].long()]
            rows = _shard.gather_rows(rows.to(self._gather_device)).cpu()
            scene_ims = vocab
            if _shard.rank() != 0:
                return None
        self._preds_from_rows(rows, scene_ims)
        eval_res: Dict[str, Any] = {}
        for refine_i in range(self.n_iter_test + 1):
            eval_res[f"iter{refine_i}"] = self._eval_predictions(refine_i)
        return eval_res

    def predictions(self) -> "OrderedDict[str, Dict[str, Dict[str, np.ndarray]]]":
        """the structure the reference caches as ``*_preds.pkl`` (valid after evaluate())"""
        return self._predictions_dict

    def merged_results(self, cur_iter: int) -> List[Dict[str, Any]]:
        """pred_gt_merge_list of _eval_predictions (:236-243): every ground-truth image, with its predictions or the
        empty prediction set."""
        if self.dataset_dicts is None:
            raise ValueError("NocsPredictionCollector needs dataset_dicts (the ground truth) to evaluate")
        empty = dict(pred_class_ids=np.array([]).astype(np.int32), pred_scores=np.array([]).astype(np.float32),
                     pred_bboxes=np.empty((0, 4), dtype=np.int32), pred_RTs=np.empty((0, 4, 4), dtype=np.float32),
                     pred_scales=np.empty((0, 3), dtype=np.float32))
        preds = self._predictions_dict[f"iter{cur_iter}"]
        merged = []
        for key, gt in build_gt_dict(self.dataset_dicts).items():
            gt = dict(gt)
            gt.update(preds[key] if key in preds else empty)
            merged.append(gt)
        return merged

    def _eval_predictions(self, cur_iter: int = 0) -> Dict[str, Any]:
        """catre_custom_evaluator.py:215-330"""
        method_name = f"{self._exp_id.replace('_', '-')}"
        if cur_iter == 0 and self._output_dir:
            os.makedirs(self._output_dir, exist_ok=True)
            with open(os.path.join(self._output_dir, f"{method_name}_{self.dataset_name}_preds.pkl"), "wb") as f:
                pickle.dump(self._predictions_dict, f)
        iou_3d_aps, pose_aps = _map.compute_independent_mAP(
            self.merged_results(cur_iter), ["BG"] + self.obj_names, degree_thresholds=DEGREE_THRESHOLDS,
            shift_thresholds=SHIFT_THRESHOLDS, iou_3d_thresholds=IOU_3D_THRESHOLDS, backend=self._map_backend)
        table = format_table(iou_3d_aps, pose_aps, self.obj_names)
        logger.info("Eval recalls of results at iter=%d...\n%s", cur_iter, table)
        if self._output_dir:
            with open(os.path.join(self._output_dir, f"{method_name}_{self.dataset_name}_tab_iter{cur_iter}.txt"), "w") as f:
                f.write("{}\n".format(table))
        return {"iou_3d_aps": iou_3d_aps, "pose_aps": pose_aps, "table": table}

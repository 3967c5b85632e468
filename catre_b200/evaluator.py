"""Evaluator K-loop with cross-image batching.

Stands in for ``catre_inference_on_dataset`` (core/catre/engine/catre_evaluator.py:225-369) and its input
collation ``batch_data_test`` (core/catre/engine/batch_test.py:10-60).  The reference refines the objects of
ONE data-loader item (one image, ~5.6 objects on REAL275) per launch and runs ``batch_updater_test`` + the
model K times from Python; here loader items are queued until ``objects_per_launch`` objects are pending, the
whole K-loop of all of them runs as one ``refine`` call of the engine (catre_b200.dropin.CatreB200), every
iteration's pose comes back with ONE device->host copy, and the caller's evaluator is then fed per loader
item exactly as the reference feeds it: ``evaluator.process(inputs, batch, outputs, out_dict)`` with
``batch["im_id"] / ["inst_id"] / ["obj_cls"]`` and ``out_dict["pose_i"] / ["scale_i"]`` for i = 0..K
(catre_evaluator.py:86-170 consumes exactly these).  Objects are independent on this path (GroupNorm is per
object), so regrouping them across images does not change any object's result (bit-exact:
tests/test_evaluator.py).

No reference code is imported; the loader items are duck-typed: ``d["instances"]`` needs the attributes
batch_data_test reads (obj_classes, obj_boxes.tensor, obj_poses.tensor, obj_scales, obj_mean_points,
obj_mean_scales, pcl, obj_sym_infos) and ``d["cam"]``.  There is no CPU path: the model must be the
CUDA drop-in.
"""
from __future__ import annotations

import datetime
import logging
import os
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import torch

from . import shard as _shard

logger = logging.getLogger(__name__)


def _cfg_get(cfg: Any, path: str, default: Any = None) -> Any:
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        cur = cur.get(key, None) if isinstance(cur, dict) else getattr(cur, key, None)
    return default if cur is None else cur


def _tensor_of(x: Any) -> torch.Tensor:
    """detectron2 ``Boxes`` / the reference's ``MyList``-style wrappers keep their data in ``.tensor``."""
    return x.tensor if hasattr(x, "tensor") and not isinstance(x, torch.Tensor) else x


def batch_data_test(cfg: Any, data: Sequence[Dict[str, Any]], device: str = "cuda", dtype=torch.float32) -> Dict[str, Any]:
    """Flatten the instances of a list of images into one batch; same keys, dtypes and order as the
    reference's batch_data_test (core/catre/engine/batch_test.py:10-60) for the shipped input config
    (KPS_TYPE="mean_shape", no image / depth inputs).  ``obj_kps`` (the per-object prior) is filled here
    as get_normed_kps does on the first batch_updater_test call (engine_utils.py:17-24)."""
    kps_type = str(_cfg_get(cfg, "INPUT.KPS_TYPE", "mean_shape")).lower()
    if kps_type != "mean_shape":
        raise NotImplementedError(f"INPUT.KPS_TYPE={kps_type!r}: the engine implements the shipped 'mean_shape' priors")
    fl = dict(dtype=dtype, device=device, non_blocking=True)
    lg = dict(dtype=torch.long, device=device, non_blocking=True)

    def cat(attr, kw):
        parts = [torch.as_tensor(_tensor_of(getattr(d["instances"], attr))) for d in data]
        return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)).to(**kw)  # one image per item is the usual case

    batch: Dict[str, Any] = {}
    batch["obj_cls"] = cat("obj_classes", lg)
    batch["obj_bbox"] = cat("obj_boxes", fl)
    batch["obj_pose_est"] = cat("obj_poses", fl)
    batch["obj_scale_est"] = cat("obj_scales", fl)
    batch["obj_mean_points"] = cat("obj_mean_points", fl)
    batch["obj_mean_scales"] = cat("obj_mean_scales", fl)
    im_ids: List[int] = []
    inst_ids: List[int] = []
    k_list = []
    sym_infos: List[Any] = []
    for i_im, d in enumerate(data):
        n_inst = len(d["instances"])
        sym_infos.extend(list(getattr(d["instances"], "obj_sym_infos", [None] * n_inst)))
        im_ids.extend([i_im] * n_inst)
        inst_ids.extend(range(n_inst))
        if n_inst:
            k_list.append(torch.as_tensor(d["cam"]).reshape(1, 3, 3).expand(n_inst, 3, 3))  # one K per instance (batch_test.py:41-47)
    batch["im_id"] = torch.tensor(im_ids, dtype=dtype, device=device)  # float, like the reference (batch_test.py:45)
    batch["inst_id"] = torch.tensor(inst_ids, dtype=dtype, device=device)
    batch["K"] = (torch.cat(k_list, dim=0) if k_list else torch.zeros(0, 3, 3)).to(**fl).contiguous()
    batch["sym_info"] = sym_infos
    batch["pcl"] = cat("pcl", fl)
    batch["obj_kps"] = batch["obj_mean_points"]
    return batch


@dataclass
class _Pending:
    inputs: Sequence[Dict[str, Any]]
    batch: Dict[str, Any]
    n_obj: int
    t_collate: float


@dataclass
class InferenceStats:
    images: int = 0
    objects: int = 0
    launches: int = 0
    compute_s: float = 0.0
    process_s: float = 0.0
    total_s: float = 0.0
    objects_per_launch: List[int] = field(default_factory=list)


def _filter_labels(evaluator: Any, batch: Dict[str, Any], device: str) -> bool:
    """The reference's test-label -> train-label adaptation (catre_evaluator.py:270-289).  Returns False when
    no object of this item is a trained class (the reference skips the item)."""
    if getattr(evaluator, "train_objs", None) is None:
        return True
    test_labels = batch["obj_cls"].cpu().numpy().tolist()
    train_labels, keep = [], []
    for i, lab in enumerate(test_labels):
        train_label, _ = evaluator._maybe_adapt_label_cls_name(lab)
        if train_label is not None:
            train_labels.append(train_label)
            keep.append(i)
    if not keep:
        return False
    n_all = len(test_labels)
    keep_t = torch.tensor(keep, device=device, dtype=torch.long)
    for k in list(batch):
        v = batch[k]
        if len(v) != n_all:
            continue
        if isinstance(v, torch.Tensor):
            batch[k] = v[keep_t]
        elif isinstance(v, list):
            batch[k] = [v[i] for i in keep]
    batch["obj_cls"] = torch.tensor(train_labels, device=device, dtype=torch.long)
    return True


@dataclass
class _Launch:
    items: List[_Pending]
    n_obj: int
    poses_h: torch.Tensor   # [K+1, n, 3, 4] host (pinned on the CUDA path)
    scales_h: torch.Tensor  # [K+1, n, 3]
    ev_start: Any
    ev_end: Any
    host_s: float           # host time spent staging + enqueueing this launch


class CrossImageRefiner:
    """Queue loader items, refine ``objects_per_launch`` objects per engine call, hand results back per item.

    The launch is asynchronous: flush() stages the pending objects into one of two pinned host buffers, enqueues
    H2D copies + the engine's K-loop + the D2H copy of every iteration's pose, and returns; the PREVIOUS launch is
    then finished (event wait, evaluator.process per loader item), so collation and result collection of one launch
    run on the host underneath the device work of the next."""

    _KEYS = (("pcl", "pcl"), ("obj_kps", "prior"), ("obj_pose_est", "pose"), ("obj_scale_est", "scale"), ("K", "K"))

    def __init__(self, cfg: Any, model: Any, evaluator: Any, n_iter: int, objects_per_launch: int = 256,
                 device: str = "cuda"):
        if not hasattr(model, "refine"):
            raise TypeError("CrossImageRefiner needs the catre_b200 drop-in model (a .refine(pcl, prior, pose, scale, K, "
                            "n_iter) entry); there is no per-iteration PyTorch fallback")
        self.cfg, self.model, self.evaluator = cfg, model, evaluator
        self.n_iter = int(n_iter)
        self.objects_per_launch = max(1, int(objects_per_launch))
        self.device = device
        self.on_gpu = str(device).startswith("cuda")
        self.pending: List[_Pending] = []
        self.pending_objs = 0
        self.stats = InferenceStats()
        self._stage: List[Optional[Dict[str, torch.Tensor]]] = [None, None]
        self._cur = 0
        self._inflight: Optional[_Launch] = None

    def add(self, inputs: Sequence[Dict[str, Any]]) -> None:
        t0 = time.perf_counter()
        self.stats.images += len(inputs)
        if getattr(self.evaluator, "train_objs", None) is None and os.environ.get("CATRE_EVAL_COLLATE", "item") == "launch":
            # opt-in (CATRE_EVAL_COLLATE=launch): only count the item's objects here and collate all items of a launch with ONE
            # batch_data_test call in flush().  Fewer torch calls in total, but they then sit between two launches instead of
            # underneath the previous one: measured 0.141 s vs 0.135 s for 400 images (profiles/r02h_bench_evaluator_loop.json),
            # so collation per loader item stays the default
            batch = None
            n_obj = sum(len(d["instances"]) for d in inputs)
            if n_obj == 0:
                return  # nothing to refine for this item (the reference `continue`s)
        else:
            # collate on the host: the objects cross to the device once per launch (flush), not once per image
            batch = batch_data_test(self.cfg, inputs, device="cpu")
            if int(batch["obj_cls"].shape[0]) == 0 or not _filter_labels(self.evaluator, batch, "cpu"):
                return
            n_obj = int(batch["obj_cls"].shape[0])
        # never let a launch grow past objects_per_launch (normally the engine's max_batch): the engine would split it into a
        # full chunk plus a few-object chunk, and a few-object K-loop is latency-bound (1 ms for 8 objects, as long as 20 of
        # a full launch's objects) -- launch what is pending first
        if self.pending and self.pending_objs + n_obj > self.objects_per_launch:
            self.flush()
        self.pending.append(_Pending(inputs, batch, n_obj, time.perf_counter() - t0))
        self.pending_objs += n_obj
        if self.pending_objs >= self.objects_per_launch:
            self.flush()

    def _collate_launch(self, items: List[_Pending]) -> Dict[str, Any]:
        """One batch_data_test over every image of the launch; the per-item batches the evaluator receives are views of it
        (same values as a per-item collation; im_id counts the images of the item, as batch_data_test does per item)."""
        big = batch_data_test(self.cfg, [d for it in items for d in it.inputs], device="cpu")
        sizes = [it.n_obj for it in items]
        n = sum(sizes)
        parts: Dict[str, Any] = {}
        for k, v in big.items():
            if k == "obj_kps":
                continue  # the same tensor as obj_mean_points
            if isinstance(v, torch.Tensor) and v.shape[0] == n:
                parts[k] = v.split(sizes, dim=0)
        lo = im0 = 0
        share = (time.perf_counter() - self._t_flush) / max(1, len(items))
        for i, it in enumerate(items):
            b: Dict[str, Any] = {k: p[i] for k, p in parts.items()}
            if im0:
                b["im_id"] = b["im_id"] - float(im0)
            b["sym_info"] = big["sym_info"][lo: lo + it.n_obj]
            b["obj_kps"] = b["obj_mean_points"]
            it.batch = b
            it.t_collate += share
            lo += it.n_obj
            im0 += len(it.inputs)
        return big

    def _staging(self, n: int, items: List[_Pending]) -> Dict[str, torch.Tensor]:
        st = self._stage[self._cur]
        if st is None or st["cap"] < n:
            cap = max(n, self.objects_per_launch + 64)
            st = {"cap": cap}
            for key, name in self._KEYS:
                shape = (cap,) + tuple(items[0].batch[key].shape[1:])
                st[name] = torch.empty(shape, dtype=torch.float32, pin_memory=self.on_gpu)
            st["out_pose"] = torch.empty((self.n_iter + 1) * cap * 12, dtype=torch.float32, pin_memory=self.on_gpu)
            st["out_scale"] = torch.empty((self.n_iter + 1) * cap * 3, dtype=torch.float32, pin_memory=self.on_gpu)
            self._stage[self._cur] = st
        return st

    def flush(self) -> None:
        """Launch everything that is pending; finish the launch before it."""
        if self.pending:
            items, self.pending = self.pending, []
            n, self.pending_objs = self.pending_objs, 0
            t0 = self._t_flush = time.perf_counter()
            big = self._collate_launch(items) if items[0].batch is None else None
            st = self._staging(n, items)
            args = []
            for key, name in self._KEYS:
                if big is not None:
                    st[name][:n].copy_(big[key])
                else:
                    torch.cat([it.batch[key] for it in items], dim=0, out=st[name][:n])
                args.append(st[name][:n].to(self.device, non_blocking=True) if self.on_gpu else st[name][:n])
            k1 = self.n_iter + 1
            poses_h = st["out_pose"][: k1 * n * 12].view(k1, n, 3, 4)
            scales_h = st["out_scale"][: k1 * n * 3].view(k1, n, 3)
            ev_start = ev_end = None
            if self.on_gpu:
                ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev_start.record()
            poses, scales = self.model.refine(*args, self.n_iter)
            # ONE device->host copy for every iteration's pose of every object of the launch; the evaluator's own
            # .detach().cpu().numpy() calls (catre_evaluator.py:96-100) then cost nothing
            poses_h.copy_(poses, non_blocking=True)
            scales_h.copy_(scales, non_blocking=True)
            if self.on_gpu:
                ev_end.record()
            launch = _Launch(items, n, poses_h, scales_h, ev_start, ev_end, time.perf_counter() - t0)
            self.stats.launches += 1
            self.stats.objects += n
            self.stats.objects_per_launch.append(n)
            self._cur ^= 1
            prev, self._inflight = self._inflight, launch
        else:
            prev, self._inflight = self._inflight, None
        if prev is not None:
            self._finish(prev)

    def drain(self) -> None:
        """Launch what is pending and finish every launch in flight (end of the loader)."""
        self.flush()
        if self._inflight is not None:
            last, self._inflight = self._inflight, None
            self._finish(last)

    def _finish(self, l: _Launch) -> None:
        dt = l.host_s
        if l.ev_end is not None:
            l.ev_end.synchronize()
            dt = max(dt, l.ev_start.elapsed_time(l.ev_end) * 1e-3)  # device time of H2D + K-loop + D2H
        self.stats.compute_s += dt + sum(it.t_collate for it in l.items)
        t1 = time.perf_counter()
        poses_h, scales_h = l.poses_h.clone(), l.scales_h.clone()  # the pinned buffer is reused two launches later
        # one split per launch + one unbind per item instead of 2 (K+1) slicing calls per item (host time, not device time,
        # bounds this loop at 256 objects per launch)
        sizes = [it.n_obj for it in l.items]
        pose_parts, scale_parts = poses_h.split(sizes, dim=1), scales_h.split(sizes, dim=1)
        names_p = [f"pose_{i}" for i in range(self.n_iter + 1)]
        names_s = [f"scale_{i}" for i in range(self.n_iter + 1)]
        for it, pp, sp in zip(l.items, pose_parts, scale_parts):
            pi, si = pp.unbind(0), sp.unbind(0)
            out_dict = dict(zip(names_p, pi))
            out_dict.update(zip(names_s, si))
            # the reference leaves the last iteration's estimate in the batch (batch_test.py:73-77)
            it.batch["obj_pose_est"] = pi[self.n_iter]
            it.batch["obj_scale_est"] = si[self.n_iter]
            share = it.t_collate + dt * (it.n_obj / float(l.n_obj))  # this item's share of the launch
            outputs = [{"time": share} for _ in range(len(it.inputs))]
            self.evaluator.process(it.inputs, it.batch, outputs, out_dict)
        self.stats.process_s += time.perf_counter() - t1


class _NoOpEvaluator:
    train_objs = None

    def reset(self):
        pass

    def process(self, inputs, batch, outputs, out_dict):
        pass

    def evaluate(self):
        return None


def catre_inference_on_dataset(cfg, model, data_loader, evaluator, amp_test: bool = False, objects_per_launch: int = 256,
                               device: str = "cuda", return_stats: bool = False):
    """Same contract as the reference's catre_inference_on_dataset(cfg, model, data_loader, evaluator, amp_test)
    (catre_evaluator.py:225-369): resets the evaluator, feeds it every loader item's refined poses
    (all N_ITER_TEST+1 iterations) and returns ``evaluator.evaluate()`` (``{}`` when that is None).

    ``amp_test`` must be False: the engine's precision is chosen at build time."""
    if amp_test:
        raise NotImplementedError("amp_test=True: pick the engine precision at build time instead "
                                  "(build_model_optimizer(..., precision='bf16'))")
    if evaluator is None:
        evaluator = _NoOpEvaluator()
    evaluator.reset()
    n_iter = int(_cfg_get(cfg, "MODEL.CATRE.N_ITER_TEST", 4))
    total = len(data_loader)
    logger.info("Start inference on %d images", total)
    was_training = bool(getattr(model, "training", False))
    if hasattr(model, "eval"):
        model.eval()
    runner = CrossImageRefiner(cfg, model, evaluator, n_iter, objects_per_launch, device)
    t_start = time.perf_counter()
    with torch.no_grad():
        for inputs in data_loader:
            runner.add(inputs)
        runner.drain()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    st = runner.stats
    st.total_s = time.perf_counter() - t_start
    if was_training and hasattr(model, "train"):
        model.train()
    per_img = st.total_s / max(1, total)
    # NOTE same three lines the reference logs (the first is "parsed by grep", catre_evaluator.py:336-360)
    logger.info("Total inference time: %s (%.6f s / img per device, on %d devices)",
                str(datetime.timedelta(seconds=st.total_s)), per_img, _shard.world_size())
    logger.info("Total inference pure compute time: %s (%.6f s / img per device, on %d devices)",
                str(datetime.timedelta(seconds=int(st.compute_s))), st.compute_s / max(1, total), _shard.world_size())
    logger.info("Total inference post process time: %s (%.6f s / img per device, on %d devices)",
                str(datetime.timedelta(seconds=int(st.process_s))), st.process_s / max(1, total), _shard.world_size())
    results = evaluator.evaluate()
    if results is None:
        results = {}
    return (results, st) if return_stats else results


class PosePredictionCollector:
    """Result collection of CATRE_Evaluator without the BOP-toolkit back end.  ``process`` takes what the
    reference's process takes (catre_evaluator.py:86-170); ``evaluate`` returns, on the main process,
    {"iter{i}": [record...]} -- the structure save_and_eval_results receives (catre_evaluator.py:179-190) --
    with the reference's record keys (catre_evaluator.py:193-222): scene_id, im_id, obj_id, score, R (row-major
    9), t [mm], scale, mug_handle, time.

    Predictions are kept as numeric rows, not Python dicts, so the multi-rank collection is one tensor
    all-gather (catre_b200.shard.gather_rows) instead of the reference's pickled-object all_gather; the dicts
    are only built once, on the main process."""

    _META = 6  # scene index, im_id, obj_id, score, mug_handle, time

    def __init__(self, obj_names: Sequence[str], obj2id: Dict[str, int], n_iter_test: int,
                 train_objs: Optional[Sequence[str]] = None, distributed: bool = False, gather_device: str = "cpu"):
        self.obj_names = list(obj_names)
        self.obj2id = dict(obj2id)
        self.n_iter_test = int(n_iter_test)
        self.train_objs = list(train_objs) if train_objs is not None else None
        self._distributed = distributed
        self._gather_device = gather_device  # "cuda" under nccl, "cpu" under gloo
        self.reset()

    def reset(self):
        k1 = self.n_iter_test + 1
        self._meta: List[List[float]] = []  # one metadata row per record
        self._order: List[int] = []  # record -> row of the concatenated pose tensors
        self._poses: List[List[torch.Tensor]] = [[] for _ in range(k1)]  # per iteration, one [n, 3, 4] tensor per item
        self._scales: List[List[torch.Tensor]] = [[] for _ in range(k1)]
        self._n_seen = 0
        self._scenes: List[str] = []

    def _maybe_adapt_label_cls_name(self, label):
        """test-set label -> (train-set label, name) or (None, None) (catre_evaluator.py:73-84)."""
        name = self.obj_names[label]
        if self.train_objs is None:
            return label, name
        if name not in self.train_objs:
            return None, None
        return self.train_objs.index(name), name

    def process(self, inputs, batch, outputs, out_dict):
        """Only bookkeeping here (this runs once per loader item underneath the next launch): the per-record metadata and
        references to the item's pose tensors; the numeric rows are assembled for all items at once in rows()."""
        k1 = self.n_iter_test + 1
        im_ids = batch["im_id"].detach().cpu().tolist()
        inst_ids = batch["inst_id"].detach().cpu().tolist()
        labels = batch["obj_cls"].detach().cpu().tolist()
        names = self.train_objs if self.train_objs is not None else self.obj_names
        meta_rows, order = [], []
        for im_i, (inp, output) in enumerate(zip(inputs, outputs)):  # records are emitted image by image
            scene_id, im_id = inp["scene_im_id"].split("/")
            if scene_id not in self._scenes:
                self._scenes.append(scene_id)
            inst = inp.get("instances", None)
            for out_i, b_im in enumerate(im_ids):
                if int(b_im) != im_i:
                    continue
                inst_id = int(inst_ids[out_i])
                score = float(inst.obj_scores[inst_id]) if inst is not None and hasattr(inst, "obj_scores") else 1.0
                handle = float(inst.mug_handle[inst_id]) if inst is not None and hasattr(inst, "mug_handle") else 1.0
                meta_rows.append([self._scenes.index(scene_id), int(im_id), self.obj2id[names[labels[out_i]]], score, handle,
                                  output["time"]])
                order.append(self._n_seen + out_i)
        self._n_seen += len(labels)
        self._meta.extend(meta_rows)
        self._order.extend(order)
        for i in range(k1):
            self._poses[i].append(out_dict[f"pose_{i}"])
            self._scales[i].append(out_dict[f"scale_{i}"])

    def rows(self) -> torch.Tensor:
        """[records, META + (K+1)*15] float64: metadata, then per iteration the 3x4 pose and the 3 scales."""
        k1 = self.n_iter_test + 1
        width = self._META + k1 * 15
        if not self._order:
            return torch.zeros((0, width), dtype=torch.float64)
        poses = torch.stack([torch.cat([t.detach().cpu() for t in self._poses[i]], dim=0) for i in range(k1)], dim=1).double()
        scales = torch.stack([torch.cat([t.detach().cpu() for t in self._scales[i]], dim=0) for i in range(k1)], dim=1).double()
        n = poses.shape[0]
        body = torch.cat((poses.reshape(n, k1, 12), scales.reshape(n, k1, 3)), dim=2).reshape(n, k1 * 15)
        idx = torch.tensor(self._order, dtype=torch.long)
        return torch.cat((torch.tensor(self._meta, dtype=torch.float64), body[idx]), dim=1)

    def evaluate(self):
        rows, scenes = self.rows(), list(self._scenes)
        if self._distributed and _shard.world_size() > 1:
            vocab = _shard.gather_vocab(scenes)
            if rows.shape[0]:  # re-index this rank's scene ids into the global vocabulary
                remap = torch.tensor([vocab.index(s) for s in scenes], dtype=torch.float64)
                rows = rows.clone()
                rows[:, 0] = remap[rows[:, 0].long()]
            rows = _shard.gather_rows(rows.to(self._gather_device)).cpu()
            scenes = vocab
            if _shard.rank() != 0:
                return None
        k1 = self.n_iter_test + 1
        out: Dict[str, List[Dict[str, Any]]] = {f"iter{i}": [] for i in range(k1)}
        body = rows[:, self._META:].reshape(-1, k1, 3 + 12)
        pose = body[:, :, :12].reshape(-1, k1, 3, 4)
        rot = pose[..., :3].reshape(-1, k1, 9).tolist()  # one conversion for all records, not one per record
        trans = (1000.0 * pose[..., 3]).tolist()
        scale = body[:, :, 12:].tolist()
        meta = rows[:, : self._META].tolist()
        for r, m in enumerate(meta):
            scene, im_id, obj_id, handle = scenes[int(m[0])], int(m[1]), int(m[2]), int(m[4])
            for i in range(k1):
                out[f"iter{i}"].append({"scene_id": scene, "im_id": im_id, "obj_id": obj_id, "score": m[3], "R": rot[r][i],
                                        "t": trans[r][i], "scale": scale[r][i], "mug_handle": handle, "time": m[5]})
        return out

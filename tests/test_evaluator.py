"""Evaluator K-loop with cross-image batching: host logic on CPU with a stand-in model,
the multi-rank result collection under gloo (world_size 2), and -- on the GPU -- the real engine behind it.

The reference semantics checked here: catre_inference_on_dataset feeds evaluator.process(inputs, batch, outputs,
out_dict) once per loader item with that item's objects in order (core/catre/engine/catre_evaluator.py:258-324),
batch_data_test's flattening order (core/catre/engine/batch_test.py:10-60), the test->train label adaptation
(catre_evaluator.py:270-289) and CATRE_Evaluator's record format (catre_evaluator.py:86-170, 193-222).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from catre_b200 import evaluator as ev
from catre_b200 import synth

N_ITER = 3
OBJ_NAMES = list(synth.CATEGORIES)
OBJ2ID = {n: i + 1 for i, n in enumerate(OBJ_NAMES)}  # reference: ref/nocs.py obj2id is 1-based
CFG = {"INPUT": {"KPS_TYPE": "mean_shape"}, "MODEL": {"CATRE": {"N_ITER_TEST": N_ITER}}}


class FakeBoxes:
    def __init__(self, t):
        self.tensor = t


class FakeInstances:
    """The attributes batch_data_test reads from detectron2 Instances (data_loader.py test branch)."""

    def __init__(self, b: synth.Batch, lo: int, hi: int, with_scores=True):
        n = hi - lo
        self.obj_classes = b.obj_cls[lo:hi]
        self.obj_boxes = FakeBoxes(torch.arange(4 * n, dtype=torch.float32).reshape(n, 4))
        self.obj_poses = FakeBoxes(b.init_pose[lo:hi])
        self.obj_scales = b.init_scale[lo:hi]
        self.obj_mean_points = b.prior[lo:hi]
        self.obj_mean_scales = torch.ones(n, 3)
        self.pcl = b.pcl[lo:hi]
        self.obj_sym_infos = [None] * n
        if with_scores:
            self.obj_scores = [0.5 + 0.01 * i for i in range(n)]
            self.mug_handle = [i % 2 for i in range(n)]
        self._n = n

    def __len__(self):
        return self._n


def make_loader(b: synth.Batch, sizes, imgs_per_item=1):
    """Loader items = lists of image dicts; image i holds sizes[i] consecutive objects of b."""
    images, lo = [], 0
    for i, n in enumerate(sizes):
        images.append({"scene_im_id": f"scene_{1 + i // 4}/{i:04d}", "cam": b.K[0], "instances": FakeInstances(b, lo, lo + n)})
        lo += n
    return [images[i:i + imgs_per_item] for i in range(0, len(images), imgs_per_item)]


class StubModel:
    """Object-wise deterministic stand-in for the engine's refine(): lets the CPU tests check the batching,
    ordering and slicing logic without a GPU (the real engine runs in the gpu test below)."""

    def __init__(self):
        self.calls = []
        self.training = False

    def eval(self):
        return self

    def refine(self, pcl, prior, init_pose, init_scale, K, n_iter):
        self.calls.append(int(pcl.shape[0]))
        poses, scales = [init_pose], [init_scale]
        for i in range(1, n_iter + 1):
            sig = pcl.mean(dim=(1, 2)) + 2.0 * prior.mean(dim=(1, 2)) + K[:, 0, 0] * 1e-3
            poses.append(init_pose + i * sig.reshape(-1, 1, 1))
            scales.append(init_scale * (1.0 + 0.1 * i) + sig.reshape(-1, 1))
        return torch.stack(poses), torch.stack(scales)


class Recorder:
    """Evaluator that records what process() receives."""

    def __init__(self, train_objs=None):
        self.train_objs = train_objs
        self.calls = []
        self.was_reset = False

    def reset(self):
        self.was_reset = True

    def _maybe_adapt_label_cls_name(self, label):
        name = OBJ_NAMES[label]
        if self.train_objs is None:
            return label, name
        if name not in self.train_objs:
            return None, None
        return self.train_objs.index(name), name

    def process(self, inputs, batch, outputs, out_dict):
        self.calls.append((inputs, {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}, outputs,
                           {k: v.clone() for k, v in out_dict.items()}))

    def evaluate(self):
        return None


def test_cross_image_batching_feeds_evaluator_like_the_reference():
    sizes = [3, 0, 5, 1, 4, 2, 6]  # ragged images, one without detections
    b = synth.make_batch(sum(sizes), 128, seed=1)
    loader = make_loader(b, sizes)
    model, rec = StubModel(), Recorder()
    res, st = ev.catre_inference_on_dataset(CFG, model, loader, rec, objects_per_launch=8, device="cpu", return_stats=True)
    assert res == {} and rec.was_reset
    # a launch never exceeds objects_per_launch: 3+5 -> 8 | 1+4+2 = 7 (the next image's 6 would not fit) | 6 ; the empty image
    # is never queued
    assert model.calls == [8, 7, 6] and st.launches == 3 and st.objects == 21 and st.images == 7
    assert len(rec.calls) == 6  # once per non-empty loader item, in loader order
    ref_p, ref_s = StubModel().refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, N_ITER)
    lo = 0
    non_empty = [(i, n) for i, n in enumerate(sizes) if n]
    for (img_i, n), (inputs, batch, outputs, out_dict) in zip(non_empty, rec.calls):
        assert inputs is loader[img_i]
        assert set(out_dict) == {f"{k}_{i}" for k in ("pose", "scale") for i in range(N_ITER + 1)}
        for i in range(N_ITER + 1):
            assert torch.equal(out_dict[f"pose_{i}"], ref_p[i, lo:lo + n]) and torch.equal(out_dict[f"scale_{i}"], ref_s[i, lo:lo + n])
        assert batch["im_id"].tolist() == [0.0] * n and batch["inst_id"].tolist() == [float(j) for j in range(n)]
        assert torch.equal(batch["obj_cls"], b.obj_cls[lo:lo + n])
        assert torch.equal(batch["obj_pose_est"], ref_p[N_ITER, lo:lo + n])  # last estimate left in the batch
        assert len(outputs) == 1 and outputs[0]["time"] > 0
        lo += n


def test_multi_image_items_and_flattening_order():
    sizes = [2, 3, 1, 4]
    b = synth.make_batch(sum(sizes), 128, seed=2)
    loader = make_loader(b, sizes, imgs_per_item=2)  # two images per loader item
    bt = ev.batch_data_test(CFG, loader[0], device="cpu")
    assert bt["im_id"].tolist() == [0, 0, 1, 1, 1] and bt["inst_id"].tolist() == [0, 1, 0, 1, 2]
    assert bt["im_id"].dtype == torch.float32 and bt["obj_cls"].dtype == torch.long
    assert torch.equal(bt["pcl"], b.pcl[:5]) and torch.equal(bt["obj_kps"], b.prior[:5]) and bt["K"].shape == (5, 3, 3)
    assert torch.equal(bt["obj_bbox"], torch.cat((loader[0][0]["instances"].obj_boxes.tensor, loader[0][1]["instances"].obj_boxes.tensor)))
    model, rec = StubModel(), Recorder()
    ev.catre_inference_on_dataset(CFG, model, loader, rec, objects_per_launch=1000, device="cpu")
    assert model.calls == [10] and len(rec.calls) == 2  # everything in one launch, flushed at the end
    assert rec.calls[1][1]["im_id"].tolist() == [0, 1, 1, 1, 1] and len(rec.calls[1][2]) == 2
    with pytest.raises(NotImplementedError):
        ev.batch_data_test({"INPUT": {"KPS_TYPE": "fps"}}, loader[0], device="cpu")
    with pytest.raises(NotImplementedError):
        ev.catre_inference_on_dataset(CFG, model, loader, rec, amp_test=True, device="cpu")
    with pytest.raises(TypeError):  # a model without the fused entry is refused, not looped over in Python
        ev.catre_inference_on_dataset(CFG, object(), loader, rec, device="cpu")


def test_label_adaptation_drops_untrained_classes():
    b = synth.make_batch(12, 128, seed=3, round_robin_cls=True)  # classes 0..5 repeating
    loader = make_loader(b, [6, 6])
    train_objs = ["mug", "bottle"]  # train-set order differs from the test-set order
    model, rec = StubModel(), Recorder(train_objs=train_objs)
    ev.catre_inference_on_dataset(CFG, model, loader, rec, objects_per_launch=64, device="cpu")
    assert model.calls == [4]  # 2 kept objects per image
    for inputs, batch, outputs, out_dict in rec.calls:
        assert batch["obj_cls"].tolist() == [1, 0]  # bottle -> 1, mug -> 0 in train order
        assert batch["inst_id"].tolist() == [0.0, 5.0] and out_dict["pose_0"].shape == (2, 3, 4)
    # an item with no trained class at all is skipped
    rec2 = Recorder(train_objs=["nothing"])
    ev.catre_inference_on_dataset(CFG, StubModel(), loader, rec2, device="cpu")
    assert rec2.calls == []


def expected_records(b, sizes, poses, scales, n_iter):
    """Straightforward restatement of CATRE_Evaluator.process + pose_prediction_to_json for the fake loader."""
    out = {f"iter{i}": [] for i in range(n_iter + 1)}
    lo = 0
    for img_i, n in enumerate(sizes):
        for j in range(n):
            for i in range(n_iter + 1):
                p = poses[i, lo + j].double().numpy()
                out[f"iter{i}"].append({
                    "scene_id": f"scene_{1 + img_i // 4}", "im_id": img_i, "obj_id": OBJ2ID[OBJ_NAMES[int(b.obj_cls[lo + j])]],
                    "score": 0.5 + 0.01 * j, "R": p[:3, :3].flatten().tolist(), "t": (1000 * p[:3, 3]).tolist(),
                    "scale": scales[i, lo + j].double().numpy().tolist(), "mug_handle": j % 2,
                })
        lo += n
    return out


def check_records(got, want):
    assert set(got) == set(want)
    for k in want:
        assert len(got[k]) == len(want[k])
        for g, w in zip(got[k], want[k]):
            assert g["time"] > 0
            for f in ("scene_id", "im_id", "obj_id", "mug_handle"):
                assert g[f] == w[f], (k, f)
            for f in ("R", "t", "scale"):
                assert np.array_equal(np.asarray(g[f]), np.asarray(w[f])), (k, f)
            assert abs(g["score"] - w["score"]) < 1e-12


def test_collector_records_match_reference_format():
    sizes = [3, 2, 0, 4, 1, 2]
    b = synth.make_batch(sum(sizes), 128, seed=4)
    loader = make_loader(b, sizes)
    col = ev.PosePredictionCollector(OBJ_NAMES, OBJ2ID, N_ITER)
    res = ev.catre_inference_on_dataset(CFG, StubModel(), loader, col, objects_per_launch=5, device="cpu")
    ref_p, ref_s = StubModel().refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, N_ITER)
    check_records(res, expected_records(b, sizes, ref_p, ref_s, N_ITER))
    assert set(res["iter0"][0]) == {"scene_id", "im_id", "obj_id", "score", "R", "t", "scale", "mug_handle", "time"}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [3, 2, 4, 1, 2, 5, 1]  # 7 images: rank 0 takes 4, rank 1 takes 3 (InferenceSampler-style contiguous split)
    b = synth.make_batch(sum(sizes), 128, seed=5)
    loader = make_loader(b, sizes)
    per = (len(loader) + world - 1) // world
    mine = loader[rank * per:(rank + 1) * per]
    col = ev.PosePredictionCollector(OBJ_NAMES, OBJ2ID, N_ITER, distributed=True)
    res = ev.catre_inference_on_dataset(CFG, StubModel(), mine, col, objects_per_launch=4, device="cpu")
    if rank == 0:
        ref_p, ref_s = StubModel().refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, N_ITER)
        try:
            check_records(res, expected_records(b, sizes, ref_p, ref_s, N_ITER))
            q.put((rank, True))
        except AssertionError as e:  # pragma: no cover
            q.put((rank, repr(e)))
    else:
        q.put((rank, res == {}))
    dist.destroy_process_group()


def test_collector_gathers_over_ranks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
def test_cross_image_refinement_on_the_engine_is_bit_exact_and_matches_golden():
    """The real engine behind the evaluator loop: regrouping objects across images does not change a single
    bit of any object's result (per-item refine == cross-image refine), and the collected records carry the
    reference's golden poses of BASELINE config 2 to 1e-4."""
    from catre_b200 import dropin
    from tests import golden_util as gu

    case = gu.load_case("c2_b64_n1024_k4")
    b = case.batch
    sizes = [5, 7, 0, 6, 3, 9, 1, 8, 4, 6, 2, 7, 6]
    assert sum(sizes) == 64
    loader = make_loader(b, sizes)
    cfg = {"INPUT": {"KPS_TYPE": "mean_shape"}, "MODEL": {"CATRE": {"N_ITER_TEST": case.n_iter}}}
    model = dropin.CatreB200(1024, 1024, precision="f16x3", max_batch=32)
    model.load_state_dict(synth.load_weights(), strict=True)
    model = model.to("cuda").eval()
    rec = Recorder()
    _, st = ev.catre_inference_on_dataset(cfg, model, loader, rec, objects_per_launch=24, return_stats=True)
    assert st.objects == 64 and st.launches < len(loader) - 1
    lo = 0
    d = b.to("cuda")
    for (inputs, batch, outputs, out_dict), n in zip(rec.calls, [s for s in sizes if s]):
        p1, s1 = model.refine(d.pcl[lo:lo + n], d.prior[lo:lo + n], d.init_pose[lo:lo + n], d.init_scale[lo:lo + n],
                              d.K[lo:lo + n], case.n_iter)  # the reference's grouping: one image per launch
        for i in range(case.n_iter + 1):
            assert torch.equal(out_dict[f"pose_{i}"], p1[i].cpu()) and torch.equal(out_dict[f"scale_{i}"], s1[i].cpu())
            assert (out_dict[f"pose_{i}"] - case.poses[i, lo:lo + n]).abs().max() <= gu.TOL
            assert (out_dict[f"scale_{i}"] - case.scales[i, lo:lo + n]).abs().max() <= gu.TOL
        lo += n
    col = ev.PosePredictionCollector(OBJ_NAMES, OBJ2ID, case.n_iter)
    res = ev.catre_inference_on_dataset(cfg, model, loader, col, objects_per_launch=256)
    last = res[f"iter{case.n_iter}"]
    assert len(last) == 64
    t_mm = torch.tensor([r["t"] for r in last])
    assert (t_mm / 1000.0 - case.poses[case.n_iter, :, :, 3]).abs().max() <= gu.TOL

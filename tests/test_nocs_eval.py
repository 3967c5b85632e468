"""The NOCS evaluation chain end to end (VERDICT r1 "next round" #2): loader items -> CrossImageRefiner ->
NocsPredictionCollector -> compute_independent_mAP, against the UNMODIFIED reference's model + CATRE_EvaluatorCustom +
compute_independent_mAP on the same synthetic evaluation set (tests/golden/make_golden_nocs_eval.py, tests/nocs_fixture.py).

CPU: a stand-in model replays the reference's own poses, so everything downstream (collection, regrouping, dtypes,
ground-truth merge, metric, table text) must equal the reference EXACTLY, single process and gathered over two gloo ranks.
GPU: the real engine produces the poses (within 1e-4 of the reference's) and the metric runs on the device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from catre_b200 import evaluator as ev
from catre_b200 import nocs_eval, synth
from oracle import metrics_oracle as mo
from tests import nocs_fixture as fx

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_nocs_eval.npz")
CFG = {"INPUT": {"KPS_TYPE": "mean_shape"}, "MODEL": {"CATRE": {"N_ITER_TEST": fx.N_ITER}}}


class ReplayModel:
    """refine() returns the reference's stored poses for the objects it is given (identified by their initial pose)."""

    def __init__(self, batch, poses, scales):
        self.init, self.poses, self.scales, self.training = batch.init_pose, poses, scales, False

    def eval(self):
        return self

    def refine(self, pcl, prior, init_pose, init_scale, K, n_iter):
        idx = [int(torch.nonzero((self.init == p).flatten(1).all(1))[0]) for p in init_pose]
        return self.poses[:, idx], self.scales[:, idx]


def check_against_golden(res, z, exact=True):
    assert set(res) == {f"iter{i}" for i in range(fx.N_ITER + 1)}
    for i in range(fx.N_ITER + 1):
        r = res[f"iter{i}"]
        if exact:
            assert np.array_equal(r["iou_3d_aps"], z[f"iou_3d_aps_{i}"], equal_nan=True), i
            assert np.array_equal(r["pose_aps"], z[f"pose_aps_{i}"], equal_nan=True), i
            assert r["table"] + "\n" == str(z[f"table_{i}"]), i  # the text the reference writes to *_tab_iter{i}.txt
        else:
            assert np.allclose(r["iou_3d_aps"], z[f"iou_3d_aps_{i}"], atol=1e-12, equal_nan=True), i
            assert np.allclose(r["pose_aps"], z[f"pose_aps_{i}"], atol=1e-12, equal_nan=True), i


def test_collector_and_metric_equal_the_reference_evaluator(tmp_path):
    z = np.load(GOLDEN)
    loader, dataset_dicts, b, _ = fx.build()
    model = ReplayModel(b, torch.from_numpy(z["poses"]), torch.from_numpy(z["scales"]))
    col = nocs_eval.NocsPredictionCollector(fx.OBJ_NAMES, fx.N_ITER, dataset_dicts, map_backend=mo.OracleBackend(),
                                            output_dir=str(tmp_path), exp_id="catre_b200", dataset_name="nocs_synth")
    res = ev.catre_inference_on_dataset(CFG, model, loader, col, objects_per_launch=16, device="cpu")
    check_against_golden(res, z)
    # the regrouped predictions: same images in the same order, same arrays, same dtypes as the reference's dict
    last = col.predictions()[f"iter{fx.N_ITER}"]
    assert list(last.keys()) == [str(k) for k in z["pred_keys"]]
    for k, p in enumerate(last.values()):
        for name in ("pred_RTs", "pred_scales", "pred_class_ids", "pred_scores", "pred_bboxes"):
            want = z[f"pred_{k}_{name}"]
            assert p[name].dtype == want.dtype and np.array_equal(p[name], want), (k, name)
    # and the files the reference writes
    for i in range(fx.N_ITER + 1):
        with open(tmp_path / f"catre-b200_nocs_synth_tab_iter{i}.txt") as f:
            assert f.read() == str(z[f"table_{i}"])
    assert (tmp_path / "catre-b200_nocs_synth_preds.pkl").exists()


def test_gt_dict_and_helpers():
    _, dataset_dicts, _, _ = fx.build()
    gts = nocs_eval.build_gt_dict(dataset_dicts + [dataset_dicts[1]])  # a repeated scene_im_id is concatenated (:99-104)
    k1 = dataset_dicts[1]["scene_im_id"]
    assert len(gts) == len(dataset_dicts) and len(gts[k1]["gt_class_ids"]) == 2 * len(dataset_dicts[1]["annotations"])
    assert gts[k1]["gt_RTs"].shape[1:] == (4, 4) and (gts[k1]["gt_RTs"][:, 3] == [0, 0, 0, 1]).all()
    assert gts[k1]["gt_class_ids"].min() >= 1 and len(gts[k1]["image_path"]) == 2
    assert nocs_eval.bbox_xyxy_to_yxyx(torch.tensor([1.9, 2.2, 3.7, 4.1])) == [2, 1, 4, 3]
    with pytest.raises(ValueError):
        nocs_eval.NocsPredictionCollector(fx.OBJ_NAMES, 1).merged_results(0)


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
    z = np.load(GOLDEN)
    loader, dataset_dicts, b, _ = fx.build()
    per = (len(loader) + world - 1) // world
    mine = loader[rank * per:(rank + 1) * per]  # InferenceSampler-style contiguous split
    model = ReplayModel(b, torch.from_numpy(z["poses"]), torch.from_numpy(z["scales"]))
    col = nocs_eval.NocsPredictionCollector(fx.OBJ_NAMES, fx.N_ITER, dataset_dicts, distributed=True, map_backend=mo.OracleBackend())
    res = ev.catre_inference_on_dataset(CFG, model, mine, col, objects_per_launch=7, device="cpu")
    try:
        if rank == 0:
            check_against_golden(res, z)
        else:
            assert res == {}
        q.put((rank, True))
    except AssertionError as e:  # pragma: no cover
        q.put((rank, repr(e)))
    dist.destroy_process_group()


def test_collector_gathers_over_ranks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
def test_engine_to_map_table_equals_the_reference_chain():
    """loader items -> CrossImageRefiner on the real engine -> NocsPredictionCollector -> device metric: the engine's poses
    are within 1e-4 of the reference model's, and the AP arrays / tables the reference's evaluator produced from ITS poses
    come out identical (no prediction of this set sits within 1e-4 of a threshold)."""
    from catre_b200 import dropin
    from tests import golden_util as gu

    z = np.load(GOLDEN)
    loader, dataset_dicts, b, _ = fx.build()
    for prec in ("fp32", "f16x3"):
        model = dropin.CatreB200(fx.N_PTS, fx.N_PTS, precision=prec, max_batch=64)
        model.load_state_dict(synth.load_weights(), strict=True)
        model = model.to("cuda").eval()
        col = nocs_eval.NocsPredictionCollector(fx.OBJ_NAMES, fx.N_ITER, dataset_dicts)
        res = ev.catre_inference_on_dataset(CFG, model, loader, col, objects_per_launch=32)
        rows = col.rows()
        body = rows[:, 8:].reshape(-1, fx.N_ITER + 1, 15)
        e_p = (body[:, :, :12].reshape(-1, fx.N_ITER + 1, 3, 4).permute(1, 0, 2, 3) - torch.from_numpy(z["poses"]).double()).abs().max()
        e_s = (body[:, :, 12:].permute(1, 0, 2) - torch.from_numpy(z["scales"]).double()).abs().max()
        assert max(float(e_p), float(e_s)) <= gu.TOL, (prec, float(e_p), float(e_s))
        check_against_golden(res, z, exact=False)
        assert res[f"iter{fx.N_ITER}"]["table"] + "\n" == str(z[f"table_{fx.N_ITER}"])

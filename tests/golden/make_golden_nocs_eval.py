"""End-to-end golden for the NOCS evaluation chain (VERDICT r1 "next round" #2): the UNMODIFIED reference model refines
the objects of a small synthetic evaluation set image by image (the evaluator K-loop, catre_evaluator.py:292-311), the
UNMODIFIED ``CATRE_EvaluatorCustom`` collects the results (process / _preds_list_to_dict,
catre_custom_evaluator.py:121-198) and evaluates them (_eval_predictions :215-330 -> compute_independent_mAP).  Stored:
the reference's poses (so a CPU test can replay them through catre_b200's collector), the per-iteration AP arrays and
the table text the reference writes.  Runs only in the build container.

Usage: python tests/golden/make_golden_nocs_eval.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from tests import nocs_fixture as fx  # noqa: E402


def main():
    ref = "/root/reference"
    mg.install_shim(ref)
    torch.set_num_threads(os.cpu_count())
    import detectron2.evaluation as d2eval  # stub package: the evaluator's base class must be a real class

    d2eval.DatasetEvaluator = type("DatasetEvaluator", (), {})
    from core.catre.engine import catre_custom_evaluator as cce
    from core.catre.engine import test_utils as tu
    from core.catre.engine.batch_test import batch_data_test, batch_updater_test

    sd = torch.load(os.path.join(ref, mg.CKPT_REL), map_location="cpu")
    cfg, model, _ = mg.build_reference_model(ref, fx.N_PTS, sd)
    cfg.MODEL.CATRE.N_ITER_TEST = fx.N_ITER
    cfg.INPUT.WITH_DEPTH = cfg.INPUT.WITH_IMG = False  # batch_data_test would only stack the (unused) depth / image crops
    loader, dataset_dicts, batch_all, _ = fx.build()

    captured = {}
    orig = cce.compute_independent_mAP

    def spy(*a, **kw):
        out = orig(*a, **kw)
        captured.setdefault("aps", []).append(out)
        return out

    cce.compute_independent_mAP = spy
    cce.DatasetCatalog.get = lambda name: dataset_dicts
    with tempfile.TemporaryDirectory() as tmp:
        ev = object.__new__(cce.CATRE_EvaluatorCustom)  # the constructor needs detectron2's MetadataCatalog; set what it sets
        ev.cfg, ev.n_iter_test, ev._distributed, ev._output_dir = cfg, fx.N_ITER, False, tmp
        ev._cpu_device, ev._logger = torch.device("cpu"), __import__("logging").getLogger("golden")
        ev.train_objs, ev.dataset_name, ev.obj_names, ev.use_cache = None, "nocs_synth", fx.OBJ_NAMES, False
        ev._empty_pred = {"pred_class_ids": np.array([]).astype(np.int32), "pred_scores": np.array([]).astype(np.float32),
                          "pred_bboxes": np.empty((0, 4), dtype=np.int32), "pred_RTs": np.empty((0, 4, 4), dtype=np.float32),
                          "pred_scales": np.empty((0, 3), dtype=np.float32)}
        cfg.EXP_ID = "catre_b200"
        ev.reset()
        poses_all, scales_all = [], []
        with torch.no_grad():
            for inputs in loader:  # catre_inference_on_dataset's loop body (catre_evaluator.py:262-324)
                if len(inputs[0]["instances"]) == 0:
                    continue
                batch = batch_data_test(cfg, inputs, device="cpu")
                out_dict = {"pose_0": batch["obj_pose_est"].clone(), "scale_0": batch["obj_scale_est"].clone()}
                pose_est = scale_est = None
                for it in range(1, fx.N_ITER + 1):
                    batch_updater_test(cfg, batch, poses_est=pose_est, scales_est=scale_est, device="cpu")
                    out = model(batch["x"], batch["tfd_kps"], init_pose=batch["obj_pose_est"], init_scale=batch["obj_scale_est"],
                                K_zoom=batch["K"], obj_class=batch["obj_cls"], mean_scales=batch["obj_mean_scales"], do_loss=False,
                                cur_iter=it)
                    out_dict.update(out)
                    pose_est, scale_est = out[f"pose_{it}"], out[f"scale_{it}"]
                ev.process(inputs, batch, [{"time": 0.0} for _ in inputs], out_dict)
                poses_all.append(torch.stack([out_dict[f"pose_{i}"] for i in range(fx.N_ITER + 1)]))
                scales_all.append(torch.stack([out_dict[f"scale_{i}"] for i in range(fx.N_ITER + 1)]))
        with np.errstate(invalid="ignore", divide="ignore"):
            res = ev.evaluate()
        assert res == {}
        out = {"poses": torch.cat(poses_all, 1).numpy(), "scales": torch.cat(scales_all, 1).numpy()}
        for i, (iou_aps, pose_aps) in enumerate(captured["aps"]):
            out[f"iou_3d_aps_{i}"], out[f"pose_aps_{i}"] = iou_aps, pose_aps
            with open(os.path.join(tmp, f"catre-b200_nocs_synth_tab_iter{i}.txt")) as f:
                out[f"table_{i}"] = np.array(f.read())
        # the regrouped predictions of the last iteration, image by image (dtype / order pin for the collector)
        last = ev._predictions_dict[f"iter{fx.N_ITER}"]
        out["pred_keys"] = np.array(list(last.keys()))
        for k, (key, p) in enumerate(last.items()):
            for name, v in p.items():
                out[f"pred_{k}_{name}"] = v
    np.savez_compressed(os.path.join(HERE, "golden_nocs_eval.npz"), **out)
    print(str(out[f"table_{fx.N_ITER}"]))
    print("iter0 vs iter4 mean IoU50:", captured["aps"][0][0][-1, 2], captured["aps"][-1][0][-1, 2])


if __name__ == "__main__":
    main()

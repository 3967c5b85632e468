"""The NOCS metric the reference's evaluator calls -- compute_independent_mAP and its chain
(core/catre/engine/test_utils.py:523-926) -- against golden vectors made by the unmodified reference functions
(tests/golden/make_golden_mAP.py: 80 synthetic images with same-class groups, misses, false positives, near-duplicate
predictions, an image without predictions / without ground truth / with neither; the evaluator's thresholds).

CPU tests drive catre_b200.nocs_map's host logic with the oracle's loop back end; the GPU tests run the product path
(pair kernel + device matching through the C ABI).  Matches and APs must equal the reference's EXACTLY; the pair
tables are compared at 2e-6 relative (fp64 trig / summation order in the last bits, IoU stored as fp32)."""
import os

import numpy as np
import pytest

from catre_b200 import nocs_map
from oracle import metrics_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_mAP.npz")
SYNSET = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]
DEG, SHIFT, IOU = [5, 10], [2, 5, 10], [0.1, 0.25, 0.50, 0.75]
KEYS = ("gt_class_ids", "gt_RTs", "gt_scales", "gt_handle_visibility", "pred_class_ids", "pred_RTs", "pred_scales", "pred_scores",
        "pred_bboxes")


def load():
    z = np.load(GOLDEN)
    results = [{n: z[f"{n}_{k}"] for n in KEYS} for k in range(int(z["n_img"]))]
    return z, results


def check_map(backend):
    z, results = load()
    iou_aps, pose_aps = nocs_map.compute_independent_mAP(results, SYNSET, degree_thresholds=DEG, shift_thresholds=SHIFT,
                                                         iou_3d_thresholds=IOU, backend=backend)
    assert iou_aps.shape == (8, 4) and pose_aps.shape == (8, 3, 4)
    assert np.array_equal(iou_aps, z["iou_3d_aps"], equal_nan=True), np.abs(iou_aps - z["iou_3d_aps"]).max()
    assert np.array_equal(pose_aps, z["pose_aps"], equal_nan=True), np.abs(pose_aps - z["pose_aps"]).max()
    assert 0.1 < iou_aps[-1, 1] < 0.9 and 0.05 < pose_aps[-1, 0, 0] < 0.9  # a non-degenerate case
    a2, p2 = nocs_map.compute_independent_mAP(results, SYNSET, degree_thresholds=DEG, shift_thresholds=SHIFT, iou_3d_thresholds=IOU,
                                              use_matches_for_pose=False, backend=backend)
    assert np.array_equal(a2, z["iou_3d_aps_nomatch"], equal_nan=True) and np.array_equal(p2, z["pose_aps_nomatch"], equal_nan=True)


def check_functions(backend, table_rtol):
    z, results = load()
    n = 0
    for k, r in enumerate(results):
        if f"fn_overlaps_{k}" not in z.files:
            continue
        gm, pm, ov, idx = nocs_map.compute_3d_matches(r["gt_class_ids"], r["gt_RTs"], r["gt_scales"], r["gt_handle_visibility"], SYNSET,
                                                      r["pred_bboxes"], r["pred_class_ids"], r["pred_scores"], r["pred_RTs"],
                                                      r["pred_scales"], IOU, backend=backend)
        assert np.array_equal(idx, z[f"fn_indices_{k}"])
        assert ov.dtype == np.float32 and np.allclose(ov, z[f"fn_overlaps_{k}"], rtol=table_rtol, atol=1e-7)
        assert np.array_equal(gm, z[f"fn_gt_matches_{k}"]) and np.array_equal(pm, z[f"fn_pred_matches_{k}"])
        rt = nocs_map.compute_RT_overlaps(r["gt_class_ids"], r["gt_RTs"], r["gt_handle_visibility"], r["pred_class_ids"], r["pred_RTs"],
                                          SYNSET, backend=backend)
        assert rt.dtype == np.float64 and np.allclose(rt, z[f"fn_rt_{k}"], rtol=table_rtol, atol=1e-7, equal_nan=True)
        gm2, pm2 = nocs_map.compute_match_from_degree_cm(rt, r["pred_class_ids"], r["gt_class_ids"], DEG + [360], SHIFT + [100],
                                                         backend=backend)
        assert np.array_equal(gm2, z[f"fn_pose_gt_matches_{k}"]) and np.array_equal(pm2, z[f"fn_pose_pred_matches_{k}"])
        n += 1
    assert n == 30
    # empty sides keep the reference's shapes (test_utils.py:725-726)
    gm, pm = nocs_map.compute_match_from_degree_cm(np.zeros((0, 3, 2)), np.zeros(0), np.array([1, 1, 2]), [5, 360], [2, 100],
                                                   backend=backend)
    assert gm.shape == (2, 2, 3) and pm.shape == (2, 2, 0) and (gm == -1).all()


def test_independent_map_host_logic_with_oracle_backend():
    check_map(mo.OracleBackend())


def test_chain_functions_with_oracle_backend():
    check_functions(mo.OracleBackend(), table_rtol=0.0)  # the oracle reproduces the reference's stored values exactly


def test_zero_padding_rows_are_refused():
    _, results = load()
    r = dict(next(x for x in results if len(x["pred_class_ids"]) > 1))
    r["pred_bboxes"] = r["pred_bboxes"].copy()
    r["pred_bboxes"][0] = 0
    with pytest.raises(ValueError):
        nocs_map.compute_independent_mAP([r], SYNSET, backend=mo.OracleBackend())


@pytest.mark.gpu
def test_independent_map_on_the_gpu_equals_reference():
    check_map(None)  # CudaBackend: pair kernel + device matching through the C ABI


@pytest.mark.gpu
def test_chain_functions_on_the_gpu():
    check_functions(None, table_rtol=2e-6)


@pytest.mark.gpu
def test_device_matching_equals_oracle_on_random_ragged_problems():
    """catre_match_greedy against the oracle loops on random sub-problems: ragged sizes, empty sides, ties, NaN angles,
    an IoU threshold of 0."""
    g = np.random.RandomState(3)
    be, ob = nocs_map.CudaBackend(), mo.OracleBackend()
    tabs0, tabs1, pcs, gcs = [], [], [], []
    for _ in range(200):
        P, G = g.randint(0, 7), g.randint(0, 7)
        ov = np.round(g.uniform(0, 1, size=(P, G)), 1).astype(np.float32)  # ties
        rt = np.stack([g.uniform(0, 20, size=(P, G)), g.uniform(0, 12, size=(P, G))], axis=-1)
        if P and G and g.uniform() < 0.3:
            rt[g.randint(0, P), g.randint(0, G), 0] = np.nan
        tabs0.append(ov); tabs1.append(rt)
        pcs.append(g.randint(1, 3, size=P)); gcs.append(g.randint(1, 3, size=G))
    for mode, tabs, ta, tb in ((0, tabs0, [0.0, 0.1, 0.5, 0.9], []), (1, tabs1, [5, 10, 360], [2, 5, 100])):
        got = nocs_map._match_many(be, mode, tabs, pcs, gcs, ta, tb)
        ref = nocs_map._match_many(ob, mode, tabs, pcs, gcs, ta, tb)
        for (g1, p1), (g2, p2) in zip(got, ref):
            assert np.array_equal(g1, g2) and np.array_equal(p1, p2)

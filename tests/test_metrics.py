"""Pairwise NOCS pose metrics: the numpy oracle against golden vectors made by the unmodified
reference functions (tests/golden/make_golden_metrics.py), the product's host matching against the reference's
matches, and -- on the GPU -- the CUDA pair kernel against the goldens.

Tolerances: the reference computes in fp64 and stores fp32.  The oracle reproduces the stored values exactly; the
CUDA kernel's fp64 trig / summation order differs in the last bits, so its fp32 results are compared at 2e-6
relative (a couple of fp32 ulps) -- and the MATCHES they induce must equal the reference's exactly."""
import os

import numpy as np
import pytest

from catre_b200 import metrics
from oracle import metrics_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_metrics.npz")
SYNSET = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]
IOU_T, DEG_T, SHIFT_T = [0.25, 0.5, 0.75], [5, 10, 360], [2, 5, 100]
MAP_DEG, MAP_SHIFT, MAP_IOU = [5, 10, 15], [0.05, 0.10, 0.20], [0.25, 0.5, 0.75]
KEYS = ("gt_cls", "gt_RTs", "gt_scales", "gt_handle", "pred_cls", "pred_RTs", "pred_scales", "pred_scores", "pred_boxes")


def images():
    z = np.load(GOLDEN)
    out = []
    for k in range(int(z["n_img"])):
        im = {n: z[f"{n}_{k}"] for n in KEYS}
        im.update(overlaps=z[f"overlaps_{k}"], rt=z[f"rt_{k}"], gt_matches=z[f"gt_matches_{k}"],
                  pred_matches=z[f"pred_matches_{k}"], indices=z[f"indices_{k}"])
        out.append(im)
    return out


def test_oracle_pair_metrics_match_reference_golden():
    n_sym = 0
    for im in images():
        ov, rt = mo.pair_metrics(im["pred_RTs"], im["pred_scales"], im["pred_cls"], im["gt_RTs"], im["gt_scales"], im["gt_cls"],
                                 im["gt_handle"], SYNSET)
        assert np.array_equal(ov, im["overlaps"]) and np.array_equal(rt, im["rt"], equal_nan=True)
        n_sym += sum(1 for i in im["pred_cls"] for j in im["gt_cls"] if i == j and SYNSET[i] in mo.SYM_Y)
    assert n_sym > 10  # the symmetric branch is exercised


def test_host_matching_equals_reference():
    """metrics.greedy_matches (product host logic) fed with the reference's own pair metrics must reproduce the
    reference's gt_matches / pred_matches; so must the oracle's restatement."""
    seen_match = False
    for im in images():
        idx = im["indices"].astype(np.int64)
        ov, rt = im["overlaps"][idx], im["rt"][idx]
        for fn in (metrics.greedy_matches, mo.greedy_matches):
            gm, pm = fn(ov, rt, im["pred_cls"][idx], im["gt_cls"], IOU_T, DEG_T, SHIFT_T)
            assert np.array_equal(gm, im["gt_matches"]) and np.array_equal(pm, im["pred_matches"])
        seen_match |= bool((im["pred_matches"] > -1).any())
    assert seen_match
    assert metrics.class_rules(SYNSET) == ((1 << 1) | (1 << 2) | (1 << 4), 0, 6)


@pytest.mark.gpu
def test_cuda_pair_metrics_match_reference_golden():
    ims = images()
    res = metrics.pair_metrics_batch(ims, SYNSET)  # one launch for all 24 images
    for im, (ov, rt) in zip(ims, res):
        assert ov.shape == im["overlaps"].shape and ov.dtype == np.float32
        assert np.allclose(ov, im["overlaps"], rtol=2e-6, atol=1e-7)
        assert np.array_equal(np.isnan(rt), np.isnan(im["rt"]))
        assert np.allclose(rt, im["rt"], rtol=2e-6, atol=2e-5, equal_nan=True)  # theta in degrees: acos near 0 amplifies ulps


@pytest.mark.gpu
def test_cuda_matches_equal_reference():
    ims = images()
    for im in ims:  # the reference-signature entry, one image at a time
        gm, pm, idx = metrics.compute_combination_3d_matches(
            im["gt_cls"], im["gt_RTs"], im["gt_scales"], im["gt_handle"], SYNSET, im["pred_boxes"], im["pred_cls"],
            im["pred_scores"], im["pred_RTs"], im["pred_scales"], IOU_T, DEG_T, SHIFT_T)
        assert np.array_equal(gm, im["gt_matches"]) and np.array_equal(pm, im["pred_matches"])
        assert np.array_equal(np.asarray(idx, np.int64), im["indices"])
    batch = [dict(gt_class_ids=im["gt_cls"], gt_RTs=im["gt_RTs"], gt_scales=im["gt_scales"], gt_handle_visibility=im["gt_handle"],
                  pred_class_ids=im["pred_cls"], pred_scores=im["pred_scores"], pred_RTs=im["pred_RTs"],
                  pred_scales=im["pred_scales"]) for im in ims]
    for im, (gm, pm, idx) in zip(ims, metrics.match_images(batch, SYNSET, IOU_T, DEG_T, SHIFT_T)):
        assert np.array_equal(gm, im["gt_matches"]) and np.array_equal(pm, im["pred_matches"])
    assert metrics.pair_metrics_batch([], SYNSET) == []


def final_results():
    return [dict(gt_class_ids=im["gt_cls"], gt_RTs=im["gt_RTs"], gt_scales=im["gt_scales"], gt_handle_visibility=im["gt_handle"],
                 pred_bboxes=im["pred_boxes"], pred_class_ids=im["pred_cls"], pred_scales=im["pred_scales"],
                 pred_scores=im["pred_scores"], pred_RTs=im["pred_RTs"]) for im in images()]


def _oracle_pairs(ims, synset_names):
    """Test-only pair stage: the CPU oracle in place of the CUDA launch, to check the host logic around it."""
    return [mo.pair_metrics(im["pred_RTs"], im["pred_scales"], im["pred_cls"], im["gt_RTs"], im["gt_scales"], im["gt_cls"],
                            im["gt_handle"], synset_names) for im in ims]


def test_map_accumulation_equals_reference_host_logic():
    """compute_combination_mAP's class splitting, score bookkeeping and AP integration, with the pair stage injected
    from the oracle (no GPU): equals the aps the unmodified reference function returned."""
    want = np.load(GOLDEN)["aps"]
    with np.errstate(invalid="ignore"):
        got = metrics.compute_combination_mAP(final_results(), SYNSET, MAP_DEG, MAP_SHIFT, MAP_IOU, pair_metrics_fn=_oracle_pairs)
    assert got.shape == want.shape == (8, 4, 4, 3)
    assert np.array_equal(got, want)
    assert want[-1].max() > 0.05  # the synthetic images produce non-trivial APs


@pytest.mark.gpu
def test_cuda_map_equals_reference():
    want = np.load(GOLDEN)["aps"]
    got = metrics.compute_combination_mAP(final_results(), SYNSET, MAP_DEG, MAP_SHIFT, MAP_IOU)
    assert np.array_equal(got, want)


def test_batched_matching_equals_reference_loops_on_random_inputs():
    """greedy_matches_batch (all images and threshold triples as array operations) == the reference's per-image loops
    (oracle restatement), entry for entry: ragged sizes, empty images, tied and zero overlaps, an IoU threshold of 0,
    NaN pose errors, a score threshold."""
    g = np.random.RandomState(11)
    items = []
    for k in range(300):
        P, G = g.randint(0, 9), g.randint(0, 9)
        ov = np.round(g.rand(P, G), 1).astype(np.float32)  # one decimal: plenty of ties
        ov[g.rand(P, G) < 0.3] = 0.0
        rt = np.stack((g.rand(P, G) * 20, g.rand(P, G) * 8), axis=-1).astype(np.float32)
        rt[g.rand(P, G) < 0.05] = np.nan
        items.append((ov, rt, g.randint(1, 4, size=P), g.randint(1, 4, size=G)))
    for iou_t, deg_t, sh_t, score_t in (([0.0, 0.25, 0.5], [5, 10, 360], [2, 5, 100], 0), ([0.3], [360], [100], 0.2)):
        got = metrics.greedy_matches_batch(items, iou_t, deg_t, sh_t, score_t)
        for (ov, rt, pc, gc), (gm, pm) in zip(items, got):
            want_g, want_p = mo.greedy_matches(ov, rt, pc, gc, iou_t, deg_t, sh_t, score_t)
            assert gm.shape == want_g.shape and pm.shape == want_p.shape and gm.dtype == want_g.dtype
            assert np.array_equal(gm, want_g) and np.array_equal(pm, want_p)
    assert metrics.greedy_matches_batch([], [0.5], [5], [2]) == []

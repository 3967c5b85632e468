"""Measure the pairwise NOCS metrics + matching: catre_b200.metrics.match_images (one CUDA launch
for all pairs of all images, host matching) against the CPU restatement of the reference loops
(oracle/metrics_oracle.py = what compute_combination_3d_matches does per image today).  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from catre_b200 import metrics  # noqa: E402
from make_golden_metrics import DEG_T, IOU_T, SHIFT_T, SYNSET, make_image  # noqa: E402  (synthetic image generator only)
from oracle import metrics_oracle as mo  # noqa: E402  (CPU baseline leg only)


def main():
    n_img, n_cpu = 2754, 150  # REAL275 has 2754 test images
    g = np.random.RandomState(7)
    ims = [make_image(g) for _ in range(n_img)]
    batch = [dict(gt_class_ids=im["gt_cls"], gt_RTs=im["gt_RTs"], gt_scales=im["gt_scales"], gt_handle_visibility=im["gt_handle"],
                  pred_class_ids=im["pred_cls"], pred_scores=im["pred_scores"], pred_RTs=im["pred_RTs"],
                  pred_scales=im["pred_scales"]) for im in ims]
    pairs = sum(len(im["pred_cls"]) * len(im["gt_cls"]) for im in ims)
    t0 = time.perf_counter()
    cpu = []
    for im in ims[:n_cpu]:
        idx = np.argsort(im["pred_scores"])[::-1] if len(im["pred_cls"]) else np.zeros(0, int)
        ov, rt = mo.pair_metrics(im["pred_RTs"][idx], im["pred_scales"][idx], im["pred_cls"][idx], im["gt_RTs"], im["gt_scales"],
                                 im["gt_cls"], im["gt_handle"], SYNSET)
        cpu.append(mo.greedy_matches(ov, rt, im["pred_cls"][idx], im["gt_cls"], IOU_T, DEG_T, SHIFT_T))
    cpu_s = time.perf_counter() - t0
    cpu_pairs = sum(len(im["pred_cls"]) * len(im["gt_cls"]) for im in ims[:n_cpu])
    metrics.match_images(batch[:50], SYNSET, IOU_T, DEG_T, SHIFT_T)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = metrics.match_images(batch, SYNSET, IOU_T, DEG_T, SHIFT_T)
    gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    metrics.pair_metrics_batch([dict(pred_RTs=im["pred_RTs"], pred_scales=im["pred_scales"], pred_cls=im["pred_cls"],
                                     gt_RTs=im["gt_RTs"], gt_scales=im["gt_scales"], gt_cls=im["gt_cls"], gt_handle=im["gt_handle"])
                                for im in ims], SYNSET)
    pair_s = time.perf_counter() - t0
    same = all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(cpu, res[:n_cpu]))
    print(json.dumps({"workload": f"{n_img} images, {pairs} (pred, gt) pairs, 3x3x3 thresholds", "matches_equal_cpu": bool(same),
                      "cpu_pairs_per_s": round(cpu_pairs / cpu_s, 1), "cpu_sample_images": n_cpu, "cpu_s_extrapolated": round(cpu_s * pairs / cpu_pairs, 2),
                      "gpu_total_s": round(gpu_s, 4), "gpu_pairs_per_s": round(pairs / gpu_s, 1),
                      "gpu_pair_metrics_only_s": round(pair_s, 4), "speedup": round((cpu_s * pairs / cpu_pairs) / gpu_s, 1)}))


if __name__ == "__main__":
    main()

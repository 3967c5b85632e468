"""Synthetic REAL275-like evaluation set shared by tests/golden/make_golden_nocs_eval.py (which runs the UNMODIFIED
reference's evaluator on it) and tests/test_nocs_eval.py: loader items in the reference's format (one image per item,
``instances`` with the attributes batch_data_test and CATRE_EvaluatorCustom.process read) and the matching
``dataset_dicts`` (the ground truth the evaluator pulls from detectron2's DatasetCatalog)."""
import torch

from catre_b200 import synth

SEED, N_PTS, N_ITER = 33, 1024, 4
SIZES = [5, 7, 0, 4, 6, 3, 8, 5, 6, 4]  # objects per image; image 2 has no detections
OBJ_NAMES = list(synth.CATEGORIES)


class Boxes:
    def __init__(self, t):
        self.tensor = t


class Instances:
    def __init__(self, b: synth.Batch, lo: int, hi: int, scores):
        n = hi - lo
        self.obj_classes = b.obj_cls[lo:hi]
        self.obj_boxes = Boxes(torch.stack([torch.tensor([10.0 + 3 * i, 20.0 + 2 * i, 110.5 + i, 140.25 + i]) for i in range(lo, hi)])
                               if n else torch.zeros(0, 4))
        self.obj_poses = Boxes(b.init_pose[lo:hi])
        self.obj_scales = b.init_scale[lo:hi]
        self.obj_mean_points = b.prior[lo:hi]
        self.obj_mean_scales = torch.ones(n, 3)
        self.pcl = b.pcl[lo:hi]
        self.obj_sym_infos = [None] * n
        self.obj_scores = scores[lo:hi]
        self._n = n

    def has(self, name):
        return hasattr(self, name)

    def __len__(self):
        return self._n


def build():
    """-> (loader items, dataset_dicts, batch, targets)"""
    total = sum(SIZES)
    b, tgt = synth.make_train_batch(total, N_PTS, SEED, round_robin_cls=True)
    g = torch.Generator().manual_seed(SEED)
    scores = torch.rand(total, generator=g) * 0.9 + 0.1
    scores[3] = scores[1]  # a tie inside image 0
    loader, dataset_dicts, lo = [], [], 0
    for i, n in enumerate(SIZES):
        key = f"scene_{1 + i // 4}/{i:04d}"
        loader.append([{"scene_im_id": key, "cam": b.K[0], "instances": Instances(b, lo, lo + n, scores)}])
        annos = []
        for j in range(lo, lo + n):
            if j % 9 == 4:
                continue  # an annotation the detector's instance has no counterpart for... and vice versa below
            annos.append({"category_id": int(b.obj_cls[j]), "bbox": [10.0 + 3 * j, 20.0 + 2 * j, 110.0 + j, 140.0 + j],
                          "pose": tgt.gt_pose[j].numpy(), "scale": tgt.gt_scale[j].numpy(), "mug_handle": int(j % 2)})
        if i == 2:  # the image without detections still has a ground-truth object
            annos.append({"category_id": 2, "bbox": [1.0, 2.0, 30.0, 40.0], "pose": tgt.gt_pose[0].numpy(),
                          "scale": tgt.gt_scale[0].numpy(), "mug_handle": 1})
        dataset_dicts.append({"scene_im_id": key, "file_name": f"/data/{key}_color.png", "annotations": annos})
        lo += n
    return loader, dataset_dicts, b, tgt

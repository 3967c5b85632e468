"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; the GPU box has no reference tree).
The reference's own ``CATRE_disR_shared`` model and ``batch_updater_test`` are imported as they are
through an import shim that fabricates stub modules for the third-party packages missing here
(detectron2, mmcv, ...) -- none of them contributes arithmetic to the
forward path.  Outputs:

  nocs_fixtures.npz             category priors + the 15,374 REAL275 initial (R, t, s)
  catre_weights_82cf930e.npz    the shipped checkpoint, tensor for tensor
  golden_<case>.npz             inputs + reference outputs (every iteration) per case
  golden_index.json             case list with shapes and the reference's CPU timing

Usage:  python tests/golden/make_golden.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import copy
import hashlib
import importlib.abc
import importlib.machinery
import json
import os
import pickle
import sys
import time
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

STUB_ROOTS = set(
    "detectron2 mmcv fvcore transforms3d pytorch3d tensorboardX IPython open3d imgaug pycocotools ujson horovod "
    "apex pytorch_lightning timm plyfile termcolor matplotlib fairscale deepspeed yacs chardet omegaconf "
    "pyquaternion skimage glfw OpenEXR albumentations imagecorruptions h5py vispy glumpy pyrender OpenGL "
    "pyassimp png imageio thop meshplex fastfunc cv2 PIL loguru setproctitle tqdm pandas seaborn numba "
    "einops scipy sklearn trimesh pypng ruamel yaml tabulate colorama".split()
)


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Only fires for modules that are genuinely missing (it sits last on sys.meta_path)."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        module.__path__ = []


def install_shim(ref_root: str) -> None:
    sys.meta_path.append(_StubFinder())
    # the three bindings that must be real
    import torch.nn as nn
    import mmcv.cnn  # noqa: stub
    import detectron2.layers.batch_norm as d2bn  # noqa: stub
    import detectron2.utils.env as d2env  # noqa: stub

    def normal_init(module, mean=0, std=1, bias=0):
        if hasattr(module, "weight") and module.weight is not None:
            nn.init.normal_(module.weight, mean, std)
        if hasattr(module, "bias") and module.bias is not None:
            nn.init.constant_(module.bias, bias)

    def constant_init(module, val, bias=0):
        if hasattr(module, "weight") and module.weight is not None:
            nn.init.constant_(module.weight, val)
        if hasattr(module, "bias") and module.bias is not None:
            nn.init.constant_(module.bias, bias)

    mmcv.cnn.normal_init = normal_init
    mmcv.cnn.constant_init = constant_init
    d2bn.BatchNorm2d = nn.BatchNorm2d
    d2bn.FrozenBatchNorm2d = type("FrozenBatchNorm2d", (nn.Module,), {})
    d2bn.NaiveSyncBatchNorm = type("NaiveSyncBatchNorm", (nn.BatchNorm2d,), {})
    d2env.TORCH_VERSION = tuple(int(x) for x in torch.__version__.split("+")[0].split(".")[:2])
    sys.path.insert(0, ref_root)


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: _to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return type(d)(_to_attr(v) for v in d)
    return d


def _merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict):
            v = dict(v)
            if v.pop("_delete_", False) or not isinstance(out.get(k), dict):
                out[k] = _merge({}, v)
            else:
                out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def load_py_config(path: str) -> dict:
    """mmcv-style python config: exec the file, honour ``_base_`` inheritance and ``_delete_``."""
    ns: dict = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not isinstance(v, types.ModuleType)}
    bases = cfg.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged: dict = {}
    for b in bases:
        merged = _merge(merged, load_py_config(os.path.normpath(os.path.join(os.path.dirname(path), b))))
    return _merge(merged, cfg)


CFG_REL = "configs/catre/NOCS_REAL/aug05_kpsMS_r9d_catreDisR_shared_tspcl_convPerRot_scaleexp_120e.py"
CKPT_REL = ("output/catre/NOCS_REAL/aug05_kpsMS_r9d_catreDisR_shared_tspcl_convPerRot_scaleexp_120e/"
            "model_final_wo_optim-82cf930e.pth")


def build_reference_model(ref_root: str, n_pts: int, state_dict):
    from core.catre.models import CATRE_disR_shared as ref_model  # the reference, unmodified

    cfg = _to_attr(load_py_config(os.path.join(ref_root, CFG_REL)))
    cfg.MODEL.DEVICE = "cpu"
    cfg.MODEL.WEIGHTS = "fixture"
    cfg.SOLVER.OPTIMIZER_NAME = cfg.SOLVER.OPTIMIZER_CFG["type"]
    cfg.SOLVER.BASE_LR = cfg.SOLVER.OPTIMIZER_CFG["lr"]
    cfg.MODEL.CATRE.ROT_HEAD.INIT_CFG.num_points = 2 * n_pts
    cfg.MODEL.CATRE.PCLNET.INIT_CFG.num_points = n_pts
    cfg.INPUT.NUM_KPS = n_pts
    cfg.INPUT.NUM_PCL = n_pts
    model, _ = ref_model.build_model_optimizer(cfg, is_test=True)
    res = model.load_state_dict(state_dict, strict=True)
    model.eval()
    return cfg, model, str(res)


@torch.no_grad()
def run_reference(cfg, model, batch_in, n_iter: int):
    """The evaluator's K-loop (reference catre_evaluator.py:292-311) around the real model."""
    from core.catre.engine.batch_test import batch_updater_test

    batch = {
        "obj_cls": batch_in.obj_cls.clone(),
        "obj_pose_est": batch_in.init_pose.clone(),
        "obj_scale_est": batch_in.init_scale.clone(),
        "obj_mean_points": batch_in.prior.clone(),
        "obj_mean_scales": torch.zeros_like(batch_in.init_scale),
        "K": batch_in.K.clone(),
        "pcl": batch_in.pcl.clone(),
    }
    poses, scales = [batch["obj_pose_est"].clone()], [batch["obj_scale_est"].clone()]
    pose_est = scale_est = None
    for it in range(1, n_iter + 1):
        batch_updater_test(cfg, batch, poses_est=pose_est, scales_est=scale_est, device="cpu")
        out = model(batch["x"], batch["tfd_kps"], init_pose=batch["obj_pose_est"], init_scale=batch["obj_scale_est"],
                    K_zoom=batch["K"], obj_class=batch["obj_cls"], mean_scales=batch["obj_mean_scales"],
                    do_loss=False, cur_iter=it)
        pose_est, scale_est = out[f"pose_{it}"], out[f"scale_{it}"]
        poses.append(pose_est.clone())
        scales.append(scale_est.clone())
    return torch.stack(poses), torch.stack(scales)


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


# (name, batch, n_pts, n_iter, seed, round_robin) -- seeds follow SURVEY.md 8(d): seed = config index
CASES = [
    ("kat", 1, 1024, 4, None, False),
    ("c1_b1_n512_k1", 1, 512, 1, 1, False),
    ("c2_b64_n1024_k4", 64, 1024, 4, 2, False),
    ("c4s_b4_n2048_k8", 4, 2048, 8, 4, False),
    ("c5s_b12_n1024_k4_mixed", 12, 1024, 4, 5, True),
    ("ragged_b3_n1024_k2", 3, 1024, 2, 7, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    ref = args.ref

    # ---- fixtures (data shipped with the reference; not source code) ----
    ckpt_path = os.path.join(ref, CKPT_REL)
    sd = torch.load(ckpt_path, map_location="cpu")
    assert len(sd) == 74 and all(v.dtype == torch.float32 for v in sd.values())
    np.savez(os.path.join(HERE, "catre_weights_82cf930e.npz"), **{k: v.numpy() for k, v in sd.items()})
    prior_path = os.path.join(ref, "datasets/NOCS/obj_models/cr_normed_mean_model_points_spd.pkl")
    with open(prior_path, "rb") as f:
        priors = pickle.load(f)
    from catre_b200 import synth

    pri = np.stack([np.asarray(priors[c], dtype=np.float64) for c in synth.CATEGORIES])
    init_path = os.path.join(ref, "datasets/NOCS/test_init_poses/init_pose_spd_nocs_real.json")
    with open(init_path) as f:
        init = json.load(f)
    pose, scale, cls = [], [], []
    for key in sorted(init.keys()):
        for inst in init[key]:
            pose.append(np.asarray(inst["pose_est"], dtype=np.float64)[:3, :4])
            scale.append(np.asarray(inst["scale_est"], dtype=np.float64))
            cls.append(int(inst["obj_id"]) - 1)
    np.savez_compressed(os.path.join(HERE, "nocs_fixtures.npz"), priors=pri, init_pose=np.stack(pose),
                        init_scale=np.stack(scale), obj_cls=np.asarray(cls, dtype=np.int16))
    print("fixtures:", pri.shape, len(pose), "instances")

    install_shim(ref)
    from oracle import catre_oracle

    fx = synth.load_fixtures()
    weights = {k: v.clone() for k, v in sd.items()}
    index = {"checkpoint_sha256": sha256(ckpt_path), "priors_sha256": sha256(prior_path),
             "init_poses_sha256": sha256(init_path), "torch": torch.__version__, "threads": args.threads,
             "cases": {}}
    for name, b, n, k, seed, rr in CASES:
        batch = synth.known_answer_inputs(fx) if seed is None else synth.make_batch(b, n, seed, rr, fx)
        w_n = catre_oracle.resize_conv_p(weights, n)
        cfg, model, load_msg = build_reference_model(ref, n, w_n)
        run_reference(cfg, model, batch, 1)  # warm-up (oneDNN primitive caches)
        t0 = time.perf_counter()
        poses, scales = run_reference(cfg, model, batch, k)
        dt = time.perf_counter() - t0
        np.savez_compressed(
            os.path.join(HERE, f"golden_{name}.npz"),
            pcl=batch.pcl.numpy(), prior_cls=batch.obj_cls.numpy().astype(np.int16),
            init_pose=batch.init_pose.numpy(), init_scale=batch.init_scale.numpy(), K=batch.K.numpy(),
            poses=poses.numpy(), scales=scales.numpy(),
        )
        index["cases"][name] = {"batch": b, "n_pts": n, "n_iter": k, "seed": seed, "round_robin": rr,
                                "ref_cpu_seconds": round(dt, 4), "ref_obj_per_s": round(b / dt, 3),
                                "load_state_dict": load_msg}
        print(f"{name}: B={b} N={n} K={k}  reference CPU {dt:.3f}s  ({b / dt:.2f} obj/s)  {load_msg}")
        if name == "kat":
            print(" iter1 t", poses[1, 0, :, 3].tolist(), "s", scales[1, 0].tolist())
            print(" iter1 R", poses[1, 0, :, :3].flatten().tolist())
            print(" iter4 t", poses[4, 0, :, 3].tolist(), "s", scales[4, 0].tolist())
            print(" iter4 R", poses[4, 0, :, :3].flatten().tolist())
    with open(os.path.join(HERE, "golden_index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()

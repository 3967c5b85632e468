"""Drop-in replacement for the reference's model plugin ``core/catre/models/CATRE_disR_shared.py``.

The reference resolves its model with ``eval(cfg.MODEL.CATRE.NAME).build_model_optimizer(cfg, is_test)``
(core/catre/main_catre.py:138) and then calls, per refinement iteration,
``model(x, tfd_kps, init_pose=..., init_scale=..., K_zoom=..., obj_class=..., mean_scales=...,
do_loss=False, cur_iter=i)`` -> ``{"pose_i": [B,3,4], "scale_i": [B,3]}``
(core/catre/engine/catre_evaluator.py:295-311, CATRE_disR_shared.py:40-124).

This module exposes the same two names with the same meaning.  The module tree holds exactly the
reference's 74 parameters under the reference's names (so ``MyCheckpointer(model).resume_or_load`` /
``load_state_dict(strict=True)`` fill it: core/utils/my_checkpoint.py:48-84), but the arithmetic is the
sm_100a kernel chain in libcatre_b200.so, reached through ``catre_b200.engine`` (ctypes, C ABI).  There
is no PyTorch implementation of the forward in this package and no CPU path: inputs must be CUDA
tensors, and a missing library or device raises.

Additionally (not in the reference): ``CatreB200.refine(...)`` runs the evaluator's whole K-loop
(batch_updater_test + forward, catre_evaluator.py:292-311, batch_test.py:63-97) in one call.

Training (``do_loss=True``): the forward runs the engine's fused training step
(``catre_train_step``: forward with saved activations, the shipped LOSS_CFG's losses and the hand-derived backward
of their sum, all CUDA) and returns ``(out_dict, loss_dict)`` like the reference (CATRE_disR_shared.py:125-165).
The loss tensors are tied to the parameters through a small autograd bridge, so the reference's training loop
(``sum(loss_dict.values()).backward(); optimizer.step()``, core/catre/engine/engine.py:318-352) works unchanged:
``backward()`` copies the engine's gradients into ``param.grad``.  The bridge assumes what that loop does -- every
loss enters the total with the same weight (a uniform factor such as an AMP loss scale is fine); per-term weights
raise.  There is still no PyTorch implementation of the network in this package.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import engine as _engine

# tensors of the shipped config that never see a gradient (the heads' unused `norm`; the reference leaves their
# .grad None, core/catre/models/heads/conv_out_per_rot_head.py:96-101, fc_trans_size_head.py:33-36)
UNUSED_PARAMS = ("rot_head.rot_head_x.norm.weight", "rot_head.rot_head_x.norm.bias", "rot_head.rot_head_y.norm.weight",
                 "rot_head.rot_head_y.norm.bias", "ts_head.norm.weight", "ts_head.norm.bias")

# loss configuration the training step implements (configs/catre/NOCS_REAL/aug05_..._120e.py:115-134 over
# configs/_base_/catre_base.py); anything else is refused
_REQUIRED_LOSS_CFG = {
    "PM_LOSS_SYM": True, "PM_NORM_BY_EXTENT": False, "PM_R_ONLY": True, "PM_WITH_SCALE": True,
    "PM_LOSS_TYPE": "L1", "PM_USE_BBOX": False,
    "ROT_LOSS_TYPE": "angular", "ROT_YAXIS_LOSS_TYPE": "L1",
    "TRANS_LOSS_TYPE": "L1", "TRANS_LOSS_DISENTANGLE": True,
    "SCALE_LOSS_TYPE": "L1",
}
# the four loss weights are free (> 0): (config key, shipped value)
_LOSS_WEIGHTS = (("PM_LW", 1.0), ("ROT_LW", 1.0), ("TRANS_LW", 1.0), ("SCALE_LW", 1.0))


class _LossBridge(torch.autograd.Function):
    """Ties the engine's loss values to the module's parameters: backward() hands out the gradients the fused
    training step already computed for d(sum of losses)."""

    @staticmethod
    def forward(ctx, losses, present, model, names, *params):
        ctx.model, ctx.names, ctx.present = model, names, present
        ctx.shapes = [p.detach() for p in params]  # only their shape / dtype / device are used
        ctx.step_id = model._train_step_id
        return losses.clone()

    @staticmethod
    def backward(ctx, grad_out):
        model = ctx.model
        if ctx.step_id != model._train_step_id:
            raise RuntimeError("catre_b200: backward() of an older training forward; the engine keeps the gradients of "
                               "the latest do_loss=True forward only (the reference's loop backpropagates each iteration "
                               "before the next forward, core/catre/engine/engine.py:318-352)")
        g = grad_out[ctx.present]
        scale = g[0]
        if not bool((g == scale).all()):
            raise NotImplementedError("catre_b200: the fused training step backpropagates the plain sum of the losses; "
                                      f"got per-term weights {g.tolist()}")
        eng = model._engine
        grads: List[Optional[torch.Tensor]] = []
        if os.environ.get("CATRE_TRAIN_FLAT_GRADS", "1") == "1":
            # default since round 2 (measured on B200: 16.3 ms vs 17.2 ms per 16-object training iteration,
            # profiles/r02a_train_loop_probe.log; green in tests/test_train_gpu.py): ONE scaled copy of the engine's gradient
            # arena, handed out as views, instead of a copy and a multiply per tensor (~136 launches + 68 allocations).
            # CATRE_TRAIN_FLAT_GRADS=0 selects the per-tensor hand-off.
            offsets, _ = eng._grad_layout()
            flat = eng.train_grads_flat(float(scale))
            for name, p, need in zip(ctx.names, ctx.shapes, ctx.needs_input_grad[4:]):
                grads.append(flat[offsets[name]: offsets[name] + p.numel()].view(p.shape) if need else None)
            return (None, None, None, None, *grads)
        for name, p, need in zip(ctx.names, ctx.shapes, ctx.needs_input_grad[4:]):
            if not need:
                grads.append(None)
                continue
            grads.append(eng.train_grad(name, torch.empty_like(p, memory_format=torch.contiguous_format)) * scale)
        return (None, None, None, None, *grads)

# (name, shape); -1 = n_obs + n_prior (conv_p is tied to the point count,
# core/catre/models/heads/conv_out_per_rot_head.py:112)
_PARAM_SPECS = (
    [("pcl_net.stn.conv1.weight", (64, 3, 1)), ("pcl_net.stn.conv1.bias", (64,)),
     ("pcl_net.stn.conv2.weight", (128, 64, 1)), ("pcl_net.stn.conv2.bias", (128,)),
     ("pcl_net.stn.conv3.weight", (1024, 128, 1)), ("pcl_net.stn.conv3.bias", (1024,)),
     ("pcl_net.stn.fc1.weight", (512, 1024)), ("pcl_net.stn.fc1.bias", (512,)),
     ("pcl_net.stn.fc2.weight", (256, 512)), ("pcl_net.stn.fc2.bias", (256,)),
     ("pcl_net.stn.fc3.weight", (9, 256)), ("pcl_net.stn.fc3.bias", (9,)),
     ("pcl_net.conv1.weight", (64, 3, 1)), ("pcl_net.conv1.bias", (64,)),
     ("pcl_net.conv2.weight", (128, 64, 1)), ("pcl_net.conv2.bias", (128,)),
     ("pcl_net.conv3.weight", (512, 128, 1)), ("pcl_net.conv3.bias", (512,)),
     ("pcl_net.conv4.weight", (1024, 512, 1)), ("pcl_net.conv4.bias", (1024,)),
     ("pcl_net.fstn.conv1.weight", (64, 64, 1)), ("pcl_net.fstn.conv1.bias", (64,)),
     ("pcl_net.fstn.conv2.weight", (128, 64, 1)), ("pcl_net.fstn.conv2.bias", (128,)),
     ("pcl_net.fstn.conv3.weight", (1024, 128, 1)), ("pcl_net.fstn.conv3.bias", (1024,)),
     ("pcl_net.fstn.fc1.weight", (512, 1024)), ("pcl_net.fstn.fc1.bias", (512,)),
     ("pcl_net.fstn.fc2.weight", (256, 512)), ("pcl_net.fstn.fc2.bias", (256,)),
     ("pcl_net.fstn.fc3.weight", (4096, 256)), ("pcl_net.fstn.fc3.bias", (4096,))]
    + [(f"rot_head.rot_head_{a}.{n}", s) for a in ("x", "y") for n, s in (
        ("norm.weight", (256,)), ("norm.bias", (256,)),
        ("layers.0.weight", (256, 1088, 1)), ("layers.0.bias", (256,)),
        ("layers.1.weight", (256,)), ("layers.1.bias", (256,)),
        ("layers.3.weight", (256, 256, 1)), ("layers.3.bias", (256,)),
        ("layers.4.weight", (256,)), ("layers.4.bias", (256,)),
        ("neck.0.weight", (3, 256, 1)), ("neck.0.bias", (3,)),
        ("conv_p.weight", (1, -1, 1)), ("conv_p.bias", (1,)))]
    + [("ts_head.norm.weight", (256,)), ("ts_head.norm.bias", (256,)),
       ("ts_head.linears.0.weight", (256, 1091)), ("ts_head.linears.0.bias", (256,)),
       ("ts_head.linears.1.weight", (256,)), ("ts_head.linears.1.bias", (256,)),
       ("ts_head.linears.3.weight", (256, 256)), ("ts_head.linears.3.bias", (256,)),
       ("ts_head.linears.4.weight", (256,)), ("ts_head.linears.4.bias", (256,)),
       ("ts_head.fc_t.weight", (3, 256)), ("ts_head.fc_t.bias", (3,)),
       ("ts_head.fc_s.weight", (3, 256)), ("ts_head.fc_s.bias", (3,))]
)


def param_specs(n_obs: int, n_prior: int):
    return [(n, tuple((n_obs + n_prior) if d == -1 else d for d in s)) for n, s in _PARAM_SPECS]


class _Holder(nn.Module):
    """A parameter container node; leaves are registered as nn.Parameter under the reference's names."""


def _cfg_get(cfg: Any, path: str, default: Any = None) -> Any:
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(key, None)
        else:
            cur = getattr(cur, key, None)
    return default if cur is None else cur


# cfg values the kernel chain is built for (SURVEY.md 5 "Config / flags"); anything else is refused.
_REQUIRED_CFG = {
    "MODEL.CATRE.PCLNET.INIT_CFG.type": "point_net",
    "MODEL.CATRE.PCLNET.INIT_CFG.global_feat": False,
    "MODEL.CATRE.PCLNET.INIT_CFG.feature_transform": True,
    "MODEL.CATRE.PCLNET.INIT_CFG.out_dim": 1024,
    "MODEL.CATRE.ROT_HEAD.ROT_TYPE": "ego_rot6d",
    "MODEL.CATRE.ROT_HEAD.CLASS_AWARE": False,
    "MODEL.CATRE.ROT_HEAD.DELTA_T_SPACE": "image",
    "MODEL.CATRE.ROT_HEAD.DELTA_T_WEIGHT": 1.0,
    "MODEL.CATRE.ROT_HEAD.T_TRANSFORM_K_AWARE": True,
    "MODEL.CATRE.ROT_HEAD.DELTA_Z_STYLE": "cosypose",
    "MODEL.CATRE.ROT_HEAD.SCLAE_TYPE": "iter_add",
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.type": "ConvOutPerRotHead",
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.num_layers": 2,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.feat_dim": 256,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.norm": "GN",
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.num_gn_groups": 32,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.act": "gelu",
    "MODEL.CATRE.TS_HEAD.WITH_KPS_FEATURE": False,
    "MODEL.CATRE.TS_HEAD.WITH_INIT_SCALE": True,
    "MODEL.CATRE.TS_HEAD.WITH_INIT_TRANS": False,
    "MODEL.CATRE.TS_HEAD.INIT_CFG.type": "FC_TransSizeHead",
    "MODEL.CATRE.TS_HEAD.INIT_CFG.num_layers": 2,
    "MODEL.CATRE.TS_HEAD.INIT_CFG.feat_dim": 256,
    "MODEL.CATRE.TS_HEAD.INIT_CFG.norm": "GN",
    "MODEL.CATRE.TS_HEAD.INIT_CFG.act": "gelu",
    "MODEL.REFINE_SCLAE": True,
    # the remaining constructor switches the kernels assume (conv_out_per_rot_head.py:85-106, fc_trans_size_head.py:9-44)
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.point_bias": True,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.norm_input": False,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.dropout": False,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.rot_dim": 3,
    "MODEL.CATRE.ROT_HEAD.INIT_CFG.kernel_size": 1,
    "MODEL.CATRE.TS_HEAD.INIT_CFG.num_gn_groups": 32,
    "MODEL.CATRE.TS_HEAD.INIT_CFG.norm_input": False,
    "MODEL.CATRE.TS_HEAD.INIT_CFG.dropout": False,
    # the fused K-loop hard-codes x = pcl - t and tfd_kps = R (s * kps) without + t (batch_test.py:84-97)
    "INPUT.ZERO_CENTER_INPUT": True,
}
# keys whose value in the reference's BASE config (configs/_base_/catre_base.py:80) differs from the shipped one and
# silently changes the arithmetic: an absent key means the base default, which is refused
_ABSENT_MEANS = {"INPUT.ZERO_CENTER_INPUT": False}


def check_cfg(cfg: Any) -> Tuple[int, int]:
    """Validate a reference config against what the engine implements; returns (n_obs, n_prior)."""
    bad = []
    for path, want in _REQUIRED_CFG.items():
        got = _cfg_get(cfg, path, _ABSENT_MEANS.get(path, want))  # other absent keys: the shipped config's value
        if (got.lower() if isinstance(got, str) else got) != (want.lower() if isinstance(want, str) else want):
            bad.append(f"{path}={got!r} (engine implements {want!r})")
    if bad:
        raise NotImplementedError("catre_b200 implements the shipped CATRE config only: " + "; ".join(bad))
    n_obs = int(_cfg_get(cfg, "INPUT.NUM_PCL", 1024))
    n_prior = int(_cfg_get(cfg, "INPUT.NUM_KPS", 1024))
    n_rot = int(_cfg_get(cfg, "MODEL.CATRE.ROT_HEAD.INIT_CFG.num_points", n_obs + n_prior))
    if n_rot != n_obs + n_prior:
        raise NotImplementedError(f"ROT_HEAD num_points={n_rot} != NUM_PCL+NUM_KPS={n_obs + n_prior}")
    return n_obs, n_prior


def reference_init(n_obs: int, n_prior: int) -> Dict[str, torch.Tensor]:
    """Fresh parameters drawn the way the reference's constructors draw them, from the GLOBAL torch RNG and in the
    reference's order, so ``torch.manual_seed(s); build_model_optimizer(cfg)`` starts training from the same tensors as
    the reference does (pinned by tests/test_dropin_init.py against the unmodified reference's build):

    * ``pcl_net`` keeps torch's stock Conv1d / Linear initialisation -- kaiming-uniform(a=sqrt 5) weights and
      uniform(+-1/sqrt(fan_in)) biases -- in construction order stn, conv1..4, fstn (pointnets/pointnet.py:15-21, 46-52,
      88-95);
    * each head is constructed (stock init, consuming the RNG) and then re-drawn by its ``_init_weights``: Conv1d / Linear
      weights ~ N(0, 0.001), biases 0, GroupNorm 1 / 0, in ``self.modules()`` order; ``fc_t`` / ``fc_s`` a second time with
      std 0.01 (heads/conv_out_per_rot_head.py:93-124, heads/fc_trans_size_head.py:28-59); rot_head (x, then y) before
      ts_head (CATRE_disR_shared.py:311-315).

    Only stock torch layers are built here; they exist to consume the RNG exactly as the reference's layers do."""
    out: Dict[str, torch.Tensor] = {}

    def keep(prefix: str, layer: nn.Module) -> nn.Module:
        out[prefix + ".weight"], out[prefix + ".bias"] = layer.weight.data, layer.bias.data
        return layer

    def tnet(prefix: str, k_in: int, k_out: int) -> None:
        keep(prefix + ".conv1", nn.Conv1d(k_in, 64, 1)); keep(prefix + ".conv2", nn.Conv1d(64, 128, 1))
        keep(prefix + ".conv3", nn.Conv1d(128, 1024, 1)); keep(prefix + ".fc1", nn.Linear(1024, 512))
        keep(prefix + ".fc2", nn.Linear(512, 256)); keep(prefix + ".fc3", nn.Linear(256, k_out))

    tnet("pcl_net.stn", 3, 9)
    keep("pcl_net.conv1", nn.Conv1d(3, 64, 1)); keep("pcl_net.conv2", nn.Conv1d(64, 128, 1))
    keep("pcl_net.conv3", nn.Conv1d(128, 512, 1)); keep("pcl_net.conv4", nn.Conv1d(512, 1024, 1))
    tnet("pcl_net.fstn", 64, 4096)

    def head_normal(layer: nn.Module, std: float) -> None:
        nn.init.normal_(layer.weight, 0.0, std)
        nn.init.constant_(layer.bias, 0.0)

    for axis in ("x", "y"):
        pre = f"rot_head.rot_head_{axis}."
        for gn in ("norm", "layers.1", "layers.4"):
            keep(pre + gn, nn.GroupNorm(32, 256))  # constructed 1 / 0 and re-set to 1 / 0: no RNG either way
        # construction order (stock init draws), then the re-draw in modules() order: layers.0, layers.3, neck.0, conv_p
        convs = [keep(pre + "layers.0", nn.Conv1d(1088, 256, 1)), keep(pre + "layers.3", nn.Conv1d(256, 256, 1)),
                 keep(pre + "neck.0", nn.Conv1d(256, 3, 1)), keep(pre + "conv_p", nn.Conv1d(n_obs + n_prior, 1, 1))]
        for c in convs:
            head_normal(c, 0.001)
    for gn in ("norm", "linears.1", "linears.4"):
        keep("ts_head." + gn, nn.GroupNorm(32, 256))
    lin = [keep("ts_head.linears.0", nn.Linear(1091, 256)), keep("ts_head.linears.3", nn.Linear(256, 256)),
           keep("ts_head.fc_t", nn.Linear(256, 3)), keep("ts_head.fc_s", nn.Linear(256, 3))]
    for l in lin:
        head_normal(l, 0.001)
    head_normal(lin[2], 0.01)
    head_normal(lin[3], 0.01)
    return out


class CatreB200(nn.Module):
    """Same constructor-independent surface as the reference's ``CATRE_disR_shared`` nn.Module:
    ``.to()``, ``.eval()``, ``.parameters()``, ``state_dict()`` with the checkpoint's keys, and
    ``forward(...)`` for one refinement iteration."""

    def __init__(self, n_obs: int = 1024, n_prior: int = 1024, precision: str = "f16x3", max_batch: int = 256,
                 cfg: Any = None):
        super().__init__()
        self.cfg = cfg
        self.n_obs, self.n_prior = int(n_obs), int(n_prior)
        self.precision = precision
        self.max_batch = int(max_batch)
        init = reference_init(n_obs, n_prior)  # the reference's own initial distribution, from the global torch RNG
        for name, shape in param_specs(n_obs, n_prior):
            parts = name.split(".")
            node: nn.Module = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Holder())
                node = node._modules[p]
            assert tuple(init[name].shape) == tuple(shape), name
            node.register_parameter(parts[-1], nn.Parameter(init[name].clone()))
        self._engine: Optional[_engine.Engine] = None
        self._packed_key = None
        self._train_versions: Optional[Dict[str, Tuple[int, int]]] = None  # per tensor (version, data_ptr) the engine's training copy holds
        self._train_step_id = 0
        self._loss_w: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)  # what the engine currently holds

    # ---- engine plumbing -----------------------------------------------------------------------
    def _weights_key(self):
        return tuple((p._version, p.data_ptr()) for p in self.parameters())

    def engine(self, device: torch.device) -> _engine.Engine:
        """The engine for ``device`` with the module's current parameters packed (lazy; re-packs after
        load_state_dict / any in-place parameter update)."""
        if device.type != "cuda":
            raise _engine.CatreError("catre_b200 runs on CUDA (sm_100a) only; there is no CPU path. "
                                     f"Got tensors on {device}.")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is None or self._engine.device != idx:
            if self._engine is not None:
                self._engine.close()
            self._engine = _engine.Engine(self.n_obs, self.max_batch, self.precision, idx, n_prior=self.n_prior)
            self._packed_key = None
        key = self._weights_key()
        if key != self._packed_key:
            self._engine.load_weights({k: v for k, v in self.state_dict().items()})
            self._packed_key = key
            self._train_versions = {n: (p._version, p.data_ptr()) for n, p in self.named_parameters()}
        return self._engine

    def _engine_for_training(self, device: torch.device) -> _engine.Engine:
        """The engine with its fp32 training copies in step with the parameters: one full load the first time, then a
        device-to-device refresh of exactly the tensors the optimiser changed (no host round trip per step)."""
        if device.type != "cuda":
            raise _engine.CatreError(f"catre_b200 runs on CUDA (sm_100a) only; there is no CPU path. Got tensors on {device}.")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is None or self._engine.device != idx or self._train_versions is None:
            self.engine(device)  # create + full load (allocates the engine's device copies)
        eng = self._engine
        changed = {}
        for n, p in self.named_parameters():
            cur = (p._version, p.data_ptr())
            if self._train_versions.get(n) != cur:
                changed[n] = p.data.float()
                self._train_versions[n] = cur
        if changed:
            eng.train_set_weights(changed)  # ONE launch for all of them (one copy per tensor before: 74 after an optimiser step)
            self._packed_key = None  # the packed inference weights are stale until the next eval-mode forward
        return eng

    # ---- the reference's forward (one iteration) ---------------------------------------------------
    def forward(self, x, tfd_kps, init_pose, init_scale, K_zoom=None, obj_class=None, gt_ego_rot=None, gt_trans=None,
                gt_scale=None, obj_kps=None, mean_scales=None, sym_info=None, do_loss=False, cur_iter=0):
        """x [B,3,N_o] and tfd_kps [B,3,N_p] as the reference passes them (permuted views of point-major
        tensors, core/catre/engine/batch_test.py:92-95).  Returns {f"pose_{cur_iter}", f"scale_{cur_iter}"}, and with
        do_loss=True ``(out_dict, loss_dict)`` (CATRE_disR_shared.py:125-165)."""
        if K_zoom is None:
            raise ValueError("K_zoom is required (T_TRANSFORM_K_AWARE=True)")
        if do_loss:
            return self._forward_train(x, tfd_kps, init_pose, init_scale, K_zoom, gt_ego_rot, gt_trans, gt_scale, obj_kps,
                                       sym_info, cur_iter)
        with torch.no_grad():
            eng = self.engine(x.device)
            x_pm = x.transpose(1, 2).contiguous().float()
            k_pm = tfd_kps.transpose(1, 2).contiguous().float()
            pose, scale = eng.forward_once(x_pm, k_pm, init_pose.float().contiguous(), init_scale.float().contiguous(),
                                           K_zoom.float().contiguous())
        return {f"pose_{cur_iter}": pose, f"scale_{cur_iter}": scale}

    # ---- the reference's training forward ---------------------------------------
    def _forward_train(self, x, tfd_kps, init_pose, init_scale, K_zoom, gt_ego_rot, gt_trans, gt_scale, obj_kps, sym_info,
                       cur_iter):
        if gt_ego_rot is None or gt_trans is None or gt_scale is None or obj_kps is None or sym_info is None:
            raise ValueError("do_loss=True needs gt_ego_rot, gt_trans, gt_scale, obj_kps and sym_info "
                             "(CATRE_disR_shared.py:126, 191, 224)")
        loss_w = check_loss_cfg(self.cfg)
        is_sym, sym_rots = split_sym_info(sym_info)
        if len(is_sym) != x.shape[0]:
            raise ValueError(f"sym_info has {len(is_sym)} entries for a batch of {x.shape[0]}")
        with torch.no_grad():
            eng = self._engine_for_training(x.device)
            if loss_w != self._loss_w:
                eng.train_set_loss_weights(*loss_w)
                self._loss_w = loss_w
            f32 = lambda t: t.detach().float().contiguous()
            gt_pose = torch.cat((f32(gt_ego_rot), f32(gt_trans).reshape(-1, 3, 1)), dim=2).contiguous()
            pose, scale, losses = eng.train_step(x.detach().transpose(1, 2).contiguous().float(),
                                                 tfd_kps.detach().transpose(1, 2).contiguous().float(), f32(obj_kps),
                                                 f32(init_pose), f32(init_scale), f32(K_zoom), gt_pose, f32(gt_scale), is_sym,
                                                 sym_rots)
        self._train_step_id += 1
        present = [n != "loss_rot" or not all(is_sym) for n in _engine.TRAIN_LOSS_NAMES]
        present = [p and (n != "loss_yaxis_rot" or any(is_sym)) for p, n in zip(present, _engine.TRAIN_LOSS_NAMES)]
        # only the tensors the shipped config uses enter the graph: the heads' unused `norm` stay out of it, so their .grad
        # stays None as in the reference and DistributedDataParallel(find_unused_parameters=True) (main_catre.py:155-160)
        # marks them unused instead of waiting for a gradient
        used = [(n, p) for n, p in self.named_parameters() if n not in UNUSED_PARAMS]
        bridged = _LossBridge.apply(losses, torch.tensor(present, device=losses.device), self, [n for n, _ in used],
                                    *[p for _, p in used])
        loss_dict = {n: bridged[i] for i, n in enumerate(_engine.TRAIN_LOSS_NAMES) if present[i]}
        _put_vis_scalars(cur_iter, pose, init_pose, K_zoom, gt_ego_rot, gt_trans)
        return {f"pose_{cur_iter}": pose, f"scale_{cur_iter}": scale}, loss_dict

    # ---- fused K-loop (additive API) ------------------------------------------------------------------
    @torch.no_grad()
    def refine(self, pcl, prior, init_pose, init_scale, K, n_iter: int = 4):
        """pcl [B,N_o,3] (batch["pcl"]), prior [B,N_p,3] (batch["obj_kps"]), init_pose [B,3,4],
        init_scale [B,3], K [B,3,3] -> poses [n_iter+1,B,3,4], scales [n_iter+1,B,3] (entry 0 = init)."""
        eng = self.engine(pcl.device)
        return eng.refine(pcl.float(), prior.float(), init_pose.float(), init_scale.float(), K.float(), n_iter)

    @torch.no_grad()
    def refine_table(self, pcl, prior_table, obj_cls, init_pose, init_scale, K, n_iter: int = 4):
        """refine() with the priors as a category table [C,N_p,3] and batch["obj_cls"] [B] (any integer
        dtype): object b uses prior_table[obj_cls[b]] (core/catre/engine/engine_utils.py:17-24)."""
        eng = self.engine(pcl.device)
        return eng.refine_table(pcl.float(), prior_table.float(), obj_cls.to(torch.int32), init_pose.float(),
                                init_scale.float(), K.float(), n_iter)

    def refine_as_out_dict(self, pcl, prior, init_pose, init_scale, K, n_iter: int = 4) -> Dict[str, torch.Tensor]:
        """The evaluator's out_dict for all iterations: {pose_0.., scale_0..} (catre_evaluator.py:292-311)."""
        poses, scales = self.refine(pcl, prior, init_pose, init_scale, K, n_iter)
        out = {}
        for i in range(n_iter + 1):
            out[f"pose_{i}"] = poses[i]
            out[f"scale_{i}"] = scales[i]
        return out


def vis_scalars(cur_iter: int, pose, init_pose, K, gt_rot, gt_trans) -> Dict[str, float]:
    """The `vis/*` scalars the reference's training forward logs (CATRE_disR_shared.py:127-146): mean rotation error
    [deg] and translation error [cm] of the batch (lib/pysixd/pose_error.py:359-374, 406-417) and, for object 0, the
    per-axis translation error [cm], the predicted / ground-truth translation and the raw translation deltas of the head
    (recovered from the pose update, pose_scale_from_delta_init.py:47-95).  Host work on [B, 3, 4] values."""
    P = pose.detach().double().cpu().numpy()
    P0 = init_pose.detach().double().cpu().numpy()
    G, gt = gt_rot.detach().double().cpu().numpy(), gt_trans.detach().double().cpu().numpy()
    Kh = K.detach().double().cpu().numpy()
    tr = np.einsum("bij,bij->b", P[:, :, :3], G)  # trace(R_est R_gt^T)
    re_deg = np.rad2deg(np.arccos(np.clip(0.5 * (np.minimum(tr, 3.0) - 1.0), -1.0, 1.0)))
    te = np.linalg.norm(gt - P[:, :, 3], axis=1)
    t, t0 = P[0, :, 3], P0[0, :, 3]
    dz = t[2] / t0[2]
    dx = (t[0] / t[2] - t0[0] / t0[2]) * Kh[0, 0, 0]
    dy = (t[1] / t[2] - t0[1] / t0[2]) * Kh[0, 1, 1]
    i = cur_iter
    out = {f"vis/error_R_{i}": float(re_deg.astype(np.float32).mean()), f"vis/error_t_{i}": float(te.astype(np.float32).mean()) * 100}
    for a, ax in enumerate("xyz"):
        out[f"vis/error_t{ax}_{i}"] = float(abs(t[a] - gt[0, a]) * 100)
        out[f"vis/t{ax}_pred_{i}"] = float(t[a])
        out[f"vis/t{ax}_delta_{i}"] = float((dx, dy, dz)[a])
        out[f"vis/t{ax}_gt_{i}"] = float(gt[0, a])
    return out


def _put_vis_scalars(cur_iter, pose, init_pose, K, gt_rot, gt_trans) -> None:
    """storage.put_scalars(**vis_dict) as the reference does (CATRE_disR_shared.py:163-164) -- when detectron2's event
    storage is there (inside the reference's training loop); silently nothing otherwise."""
    try:
        from detectron2.utils.events import get_event_storage

        storage = get_event_storage()
    except Exception:
        return
    storage.put_scalars(**vis_scalars(cur_iter, pose, init_pose, K, gt_rot, gt_trans))


def check_loss_cfg(cfg: Any) -> Tuple[float, float, float, float]:
    """The training step implements the shipped LOSS_CFG's loss types; a config object that sets something else is refused
    (absent keys take the shipped values).  Returns the loss weights (PM_LW, ROT_LW, TRANS_LW, SCALE_LW), which may be any
    positive numbers; a zero weight (the reference then drops the term from its dict) is refused."""
    bad = []
    weights = tuple(float(_cfg_get(cfg, "MODEL.CATRE.LOSS_CFG." + k, d)) for k, d in _LOSS_WEIGHTS)
    for (k, _), v in zip(_LOSS_WEIGHTS, weights):
        if not v > 0:
            bad.append(f"{k}={v!r} (implemented: > 0)")
    for key, want in _REQUIRED_LOSS_CFG.items():
        got = _cfg_get(cfg, "MODEL.CATRE.LOSS_CFG." + key, want)
        if (got.lower() if isinstance(got, str) else got) != (want.lower() if isinstance(want, str) else want):
            bad.append(f"{key}={got!r} (implemented: {want!r})")
    if _cfg_get(cfg, "MODEL.CATRE.USE_MTL", False):
        bad.append("USE_MTL=True (implemented: False)")
    if bad:
        raise NotImplementedError("catre_b200's training step implements the shipped loss config only: " + "; ".join(bad))
    return weights


def split_sym_info(sym_info) -> Tuple[List[bool], np.ndarray]:
    """The reference's per-object list ``[K x 3 x 3 rotations or None]`` (core/catre/engine/batching.py:49-63) ->
    (is_sym per object, the one rotation set).  Every symmetric NOCS object carries the same discretised rotations
    about y (core/catre/datasets/data_loader.py:385-401); differing sets are refused."""
    is_sym = [s is not None for s in sym_info]
    rots = None
    for s in sym_info:
        if s is None:
            continue
        a = (s.detach().cpu().numpy() if isinstance(s, torch.Tensor) else np.asarray(s)).astype(np.float32).reshape(-1, 3, 3)
        if rots is None:
            rots = a
        elif rots.shape != a.shape or not np.array_equal(rots, a):
            raise NotImplementedError("catre_b200: objects with different symmetry-rotation sets in one batch")
    return is_sym, (rots if rots is not None else np.zeros((0, 3, 3), np.float32))


def _build_optimizer(cfg: Any, model: nn.Module):
    """The reference builds its optimiser from cfg.SOLVER.OPTIMIZER_CFG through its own registry
    (core/utils/solver_utils.build_optimizer_with_params; the shipped config names its `Ranger`).  Inside the
    reference's tree that builder is used as is; elsewhere any torch.optim class of that name works."""
    lr = float(_cfg_get(cfg, "SOLVER.BASE_LR", _cfg_get(cfg, "SOLVER.OPTIMIZER_CFG.lr", 1e-4)))
    # the reference's three parameter groups, in its order (CATRE_disR_shared.py:292-315, model_utils.py:66-89, 144-167):
    # pcl_net at the base rate, the rotation and translation/size heads at base * LR_MULT; a FREEZE'd part is left out and
    # stops requiring gradients.  Same grouping and parameter order -> optimiser checkpoints interoperate.
    groups = []
    for prefix, cfg_key, mult_key in (("pcl_net.", "MODEL.CATRE.PCLNET", None), ("rot_head.", "MODEL.CATRE.ROT_HEAD", "LR_MULT"),
                                      ("ts_head.", "MODEL.CATRE.TS_HEAD", "LR_MULT")):
        part = [p for n, p in model.named_parameters() if n.startswith(prefix)]
        if _cfg_get(cfg, cfg_key + ".FREEZE", False):
            for p in part:
                p.requires_grad = False
            continue
        mult = float(_cfg_get(cfg, cfg_key + "." + mult_key, 1.0)) if mult_key else 1.0
        groups.append({"params": [p for p in part if p.requires_grad], "lr": lr * mult})
    if str(_cfg_get(cfg, "SOLVER.OPTIMIZER_CFG.type", "")) == "Ranger":
        # the shipped config's optimiser: same constructor and state dict as lib/torch_utils/solver/ranger.py, one fused
        # CUDA step over all tensors instead of ~700 small launches (catre_b200/optim.py)
        from .optim import FusedRanger

        kw = {k: v for k, v in dict(_cfg_get(cfg, "SOLVER.OPTIMIZER_CFG", {})).items() if k not in ("type", "_delete_", "lr")}
        return FusedRanger(groups, lr=lr, **kw)
    try:
        from core.utils.solver_utils import build_optimizer_with_params  # the reference's own builder
    except Exception:
        build_optimizer_with_params = None
    if build_optimizer_with_params is not None:
        return build_optimizer_with_params(cfg, groups)
    opt_cfg = dict(_cfg_get(cfg, "SOLVER.OPTIMIZER_CFG", {}) or {})
    name = opt_cfg.pop("type", None)
    opt_cfg.pop("_delete_", None)
    if name is None or not hasattr(torch.optim, name):
        raise NotImplementedError(f"optimizer {name!r}: run inside the reference tree (its registry provides it) or use a "
                                  "torch.optim class name in SOLVER.OPTIMIZER_CFG.type")
    opt_cfg.pop("lr", None)
    return getattr(torch.optim, name)(groups, lr=lr, **opt_cfg)


def build_model_optimizer(cfg, is_test: bool = False, precision: str = "f16x3", max_batch: int = 256):
    """Same contract as the reference's build_model_optimizer (CATRE_disR_shared.py:291-350):
    returns (model, optimizer).  ``optimizer`` is None for is_test=True, as in the reference."""
    n_obs, n_prior = check_cfg(cfg)
    pretrained = _cfg_get(cfg, "MODEL.CATRE.PCLNET.PRETRAINED", "")
    if _cfg_get(cfg, "MODEL.WEIGHTS", "") == "" and pretrained not in ("", None):
        # the reference loads these into pcl_net with mmcv's load_checkpoint (CATRE_disR_shared.py:327-346); the shipped
        # config leaves it empty.  Refuse rather than silently start from random weights.
        raise NotImplementedError(f"MODEL.CATRE.PCLNET.PRETRAINED={pretrained!r}: load the file into model.pcl_net yourself "
                                  "(state_dict names are the reference's) and clear the key")
    model = CatreB200(n_obs, n_prior, precision=precision, max_batch=max_batch, cfg=cfg)
    device = _cfg_get(cfg, "MODEL.DEVICE", "cuda")
    model.to(torch.device(device))
    if is_test:
        model.eval()
        return model, None
    check_loss_cfg(cfg)
    return model, _build_optimizer(cfg, model)

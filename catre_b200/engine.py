"""ctypes binding of libcatre_b200.so (include/catre_b200.h) -- the only way the Python side reaches the
CUDA kernels.  There is NO fallback: if the library is missing or there is no sm_100 device, every call
raises.

The reference has no FFI; this binding is what stands behind the drop-in model (catre_b200/dropin.py)
that replaces core/catre/models/CATRE_disR_shared.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import torch

from . import build as _build

PREC_FP32_SIMT, PREC_F16X3, PREC_BF16 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32_SIMT, "fp32_simt": PREC_FP32_SIMT, "f16x3": PREC_F16X3, "bf16": PREC_BF16}


class CatreCfg(ctypes.Structure):
    _fields_ = [("n_obs", ctypes.c_int32), ("n_prior", ctypes.c_int32), ("max_batch", ctypes.c_int32),
                ("precision", ctypes.c_int32), ("device", ctypes.c_int32), ("reserved", ctypes.c_int32 * 3)]


# every symbol include/catre_b200.h declares: (restype, argtypes)
_P = ctypes.c_void_p
_F = ctypes.c_void_p  # float* passed as raw addresses (tensor.data_ptr())
SYMBOLS = {
    "catre_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(CatreCfg)]),
    "catre_set_weight": (ctypes.c_int, [_P, ctypes.c_char_p, _F, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32]),
    "catre_num_weights": (ctypes.c_int32, []),
    "catre_weight_name": (ctypes.c_char_p, [ctypes.c_int32]),
    "catre_pack": (ctypes.c_int, [_P, _P]),
    "catre_workspace_bytes": (ctypes.c_size_t, [_P, ctypes.c_int32]),
    "catre_forward_once": (ctypes.c_int, [_P, _F, _F, _F, _F, _F, ctypes.c_int32, _F, _F, _P]),
    "catre_refine": (ctypes.c_int, [_P, _F, _F, _F, _F, _F, ctypes.c_int32, ctypes.c_int32, _F, _F, _P]),
    "catre_refine_host": (ctypes.c_int, [_P, _F, _F, _F, _F, _F, ctypes.c_int32, ctypes.c_int32, _F, _F, _P]),
    "catre_refine_host_packed": (ctypes.c_int, [_P, _F, _F, _F, _F, _F, ctypes.c_int32, ctypes.c_int32, _F, _F, _F, _P]),
    "catre_pack_poses": (ctypes.c_int, [_F, _F, ctypes.c_int32, ctypes.c_int32, _F, _P]),
    "catre_refine_table": (ctypes.c_int, [_P, _F, _F, _F, ctypes.c_int32, _F, _F, _F, ctypes.c_int32, ctypes.c_int32, _F, _F, _P]),
    "catre_refine_table_host": (ctypes.c_int, [_P, _F, _F, _F, ctypes.c_int32, _F, _F, _F, ctypes.c_int32, ctypes.c_int32, _F, _F, _P]),
    "catre_cloud_scratch_bytes": (ctypes.c_size_t, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "catre_cloud_select": (ctypes.c_int, [_F, _F, ctypes.POINTER(ctypes.c_float), _F, _F, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int32, _F, _F, _P, _P]),
    "catre_cloud_gather": (ctypes.c_int, [_F, ctypes.POINTER(ctypes.c_float), _F, _F, _F, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int32, ctypes.c_int32, _F, _P]),
    "catre_pair_metrics": (ctypes.c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_int32, _F, _F, _P]),
    "catre_pair_metrics_ex": (ctypes.c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_int32, ctypes.c_int32, _F, _F, _F, _P]),
    "catre_match_greedy": (ctypes.c_int, [ctypes.c_int32, _F, _F, _F, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _F, _F, _F, _F,
                                          _F, _F, _F, ctypes.c_int32, _F, ctypes.c_int32, _F, _F, _P]),
    "catre_train_set_weight": (ctypes.c_int, [_P, ctypes.c_char_p, _F, _P]),
    "catre_train_set_weights": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_void_p), _P]),
    "catre_train_set_loss_weights": (ctypes.c_int, [_P, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]),
    "catre_train_step": (ctypes.c_int, [_P, _F, _F, _F, _F, _F, _F, _F, _F, _P, _P, ctypes.c_int32, ctypes.c_int32, _F, _F, _F, _P]),
    "catre_train_grad": (ctypes.c_int, [_P, ctypes.c_char_p, _F, _P]),
    "catre_train_grad_layout": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "catre_train_grads_flat": (ctypes.c_int, [_P, _F, ctypes.c_float, _P]),
    "catre_ranger_step": (ctypes.c_int, [_F, _F, _F, _F, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, _F, _P, _P]),
    "catre_last_launch_count": (ctypes.c_int64, [_P]),
    "catre_debug_read": (ctypes.c_int, [_P, ctypes.c_char_p, _P, ctypes.c_size_t]),
    "catre_debug_train_gemm": (ctypes.c_int, [_F, _F, _F, _F, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _F, ctypes.c_int32,
                                              ctypes.c_int32, _F, _P, _P]),
    "catre_profile_enable": (ctypes.c_int, [_P, ctypes.c_int32]),
    "catre_profile_reset": (ctypes.c_int, [_P]),
    "catre_profile_num": (ctypes.c_int32, []),
    "catre_profile_name": (ctypes.c_char_p, [ctypes.c_int32]),
    "catre_profile_get": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "catre_last_error": (ctypes.c_char_p, [_P]),
    "catre_destroy": (None, [_P]),
    "catre_version": (ctypes.c_char_p, []),
}

_LIB: Optional[ctypes.CDLL] = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the in-tree libcatre_b200.so (no compute, works without a GPU) and bind every symbol."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or _build.LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: run `python -m catre_b200.build` (nvcc, sm_100a). "
                           "catre_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


class CatreError(RuntimeError):
    pass


# order of catre_train_step's out_losses (the reference's loss_dict keys, CATRE_disR_shared.py:168-288)
TRAIN_LOSS_NAMES = ("loss_PM_R", "loss_rot", "loss_yaxis_rot", "loss_trans_xy", "loss_trans_z", "loss_scale")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """One engine per device.  All tensors fp32; device entry points take CUDA tensors on the engine's
    device and run on the current torch stream without host synchronisation."""

    def __init__(self, n_pts: int, max_batch: int, precision: str = "f16x3", device: int = 0, n_prior: Optional[int] = None):
        """n_pts = observed points per object (INPUT.NUM_PCL); n_prior = prior points per object (INPUT.NUM_KPS, default
        = n_pts).  The reference only ties conv_p to the sum (conv_out_per_rot_head.py:112)."""
        self.lib = load_library()
        self.n_pts, self.max_batch, self.device = int(n_pts), int(max_batch), int(device)
        self.n_prior = int(n_pts if n_prior is None else n_prior)
        self.precision = precision
        cfg = CatreCfg(n_obs=n_pts, n_prior=self.n_prior, max_batch=max_batch, precision=PRECISIONS[precision], device=device)
        h = ctypes.c_void_p()
        rc = self.lib.catre_create(ctypes.byref(h), ctypes.byref(cfg))
        if rc != 0:
            raise CatreError(f"catre_create failed ({rc}): {self.lib.catre_last_error(None).decode()}")
        self._h = h
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self.lib.catre_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise CatreError(f"{what} failed ({rc}): {self.lib.catre_last_error(self._h).decode()}")

    # ---- weights ---------------------------------------------------------------------------------
    @staticmethod
    def weight_names():
        lib = load_library()
        return [lib.catre_weight_name(i).decode() for i in range(lib.catre_num_weights())]

    def set_weight(self, name: str, t: torch.Tensor):
        t = t.detach().to(torch.float32).contiguous()
        shape = (ctypes.c_int64 * t.dim())(*t.shape)
        self._check(self.lib.catre_set_weight(self._h, name.encode(), t.data_ptr(), shape, t.dim()), f"set_weight({name})")

    def load_weights(self, state: Dict[str, torch.Tensor]):
        """state: {checkpoint name: tensor} (CPU or CUDA).  Extra keys are an error, like a strict load."""
        names = set(self.weight_names())
        extra = sorted(set(state) - names)
        missing = sorted(names - set(state))
        if extra or missing:
            raise CatreError(f"state dict mismatch: missing {missing[:4]}..., unexpected {extra[:4]}...")
        for k, v in state.items():
            self.set_weight(k, v)
        self.pack()

    def pack(self):
        self._check(self.lib.catre_pack(self._h, self._stream()), "catre_pack")

    # ---- compute ---------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t: torch.Tensor, shape: Tuple[int, ...], name: str) -> torch.Tensor:
        if not t.is_cuda or t.device.index != self.device:
            raise CatreError(f"{name} must be a CUDA tensor on device {self.device} (got {t.device})")
        if tuple(t.shape) != tuple(shape):
            raise CatreError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
        if t.dtype != torch.float32:
            raise CatreError(f"{name} must be float32 (got {t.dtype})")
        return t if t.is_contiguous() else t.contiguous()

    def forward_once(self, x_pm, kps_pm, pose, scale, K):
        """x_pm [B,N,3], kps_pm [B,N,3] point-major; returns (pose [B,3,4], scale [B,3])."""
        B, N = x_pm.shape[0], self.n_pts
        x_pm = self._dev(x_pm, (B, N, 3), "x")
        kps_pm = self._dev(kps_pm, (B, self.n_prior, 3), "tfd_kps")
        pose = self._dev(pose, (B, 3, 4), "init_pose")
        scale = self._dev(scale, (B, 3), "init_scale")
        K = self._dev(K, (B, 3, 3), "K")
        op = torch.empty((B, 3, 4), dtype=torch.float32, device=x_pm.device)
        os_ = torch.empty((B, 3), dtype=torch.float32, device=x_pm.device)
        self._check(self.lib.catre_forward_once(self._h, _ptr(x_pm), _ptr(kps_pm), _ptr(pose), _ptr(scale), _ptr(K), B,
                                                _ptr(op), _ptr(os_), self._stream()), "catre_forward_once")
        return op, os_

    def refine(self, pcl, prior, init_pose, init_scale, K, n_iter: int, out=None):
        """All K iterations on the device.  Returns poses [n_iter+1,B,3,4], scales [n_iter+1,B,3]."""
        B, N = pcl.shape[0], self.n_pts
        pcl = self._dev(pcl, (B, N, 3), "pcl")
        prior = self._dev(prior, (B, self.n_prior, 3), "prior")
        init_pose = self._dev(init_pose, (B, 3, 4), "init_pose")
        init_scale = self._dev(init_scale, (B, 3), "init_scale")
        K = self._dev(K, (B, 3, 3), "K")
        if out is None:
            poses = torch.empty((n_iter + 1, B, 3, 4), dtype=torch.float32, device=pcl.device)
            scales = torch.empty((n_iter + 1, B, 3), dtype=torch.float32, device=pcl.device)
        else:
            poses, scales = out
        self._check(self.lib.catre_refine(self._h, _ptr(pcl), _ptr(prior), _ptr(init_pose), _ptr(init_scale), _ptr(K), B,
                                          n_iter, _ptr(poses), _ptr(scales), self._stream()), "catre_refine")
        return poses, scales

    def pack_poses(self, poses, scales, it: int, out=None):
        """poses [K+1,B,3,4], scales [K+1,B,3] (device) -> packed [B,15] of iteration `it` (one kernel; the all-gather operand)."""
        B = poses.shape[1]
        if not poses.is_cuda or not scales.is_cuda or not poses.is_contiguous() or not scales.is_contiguous():
            raise CatreError("pack_poses: contiguous CUDA tensors expected")
        if out is None:
            out = torch.empty((B, 15), dtype=torch.float32, device=poses.device)
        rc = self.lib.catre_pack_poses(_ptr(poses), _ptr(scales), B, int(it), _ptr(out), self._stream())
        if rc != 0:
            raise CatreError(f"catre_pack_poses failed ({rc}): {self.lib.catre_last_error(None).decode()}")
        return out

    def refine_host(self, pcl, prior, init_pose, init_scale, K, n_iter: int, out=None, packed_dev=None):
        """Host tensors in (pinned for async copies), host tensors out; copies are inside the call.  With ``packed_dev``
        (a CUDA [B,15] fp32 tensor) the last iteration's packed poses additionally stay on the device for the multi-GPU
        all-gather (catre_refine_host_packed)."""
        B, N = pcl.shape[0], self.n_pts
        for name, t, shp in (("pcl", pcl, (B, N, 3)), ("prior", prior, (B, self.n_prior, 3)), ("init_pose", init_pose, (B, 3, 4)),
                             ("init_scale", init_scale, (B, 3)), ("K", K, (B, 3, 3))):
            if t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != shp or not t.is_contiguous():
                raise CatreError(f"{name}: expected contiguous float32 host tensor of shape {shp}")
        if out is None:
            poses = torch.empty((n_iter + 1, B, 3, 4), dtype=torch.float32).pin_memory()
            scales = torch.empty((n_iter + 1, B, 3), dtype=torch.float32).pin_memory()
        else:
            poses, scales = out
        if packed_dev is not None:
            if not packed_dev.is_cuda or packed_dev.dtype != torch.float32 or tuple(packed_dev.shape) != (B, 15) or not packed_dev.is_contiguous():
                raise CatreError(f"packed_dev: expected a contiguous CUDA float32 tensor of shape ({B}, 15)")
            self._check(self.lib.catre_refine_host_packed(self._h, _ptr(pcl), _ptr(prior), _ptr(init_pose), _ptr(init_scale), _ptr(K),
                                                          B, n_iter, _ptr(poses), _ptr(scales), _ptr(packed_dev), self._stream()),
                        "catre_refine_host_packed")
            return poses, scales
        self._check(self.lib.catre_refine_host(self._h, _ptr(pcl), _ptr(prior), _ptr(init_pose), _ptr(init_scale), _ptr(K),
                                               B, n_iter, _ptr(poses), _ptr(scales), self._stream()), "catre_refine_host")
        return poses, scales

    def refine_table(self, pcl, prior_table, prior_cls, init_pose, init_scale, K, n_iter: int, out=None):
        """catre_refine with a category table: prior_table [C,N,3] fp32, prior_cls [B] int32 (object b uses
        prior_table[prior_cls[b]]).  Device tensors; an out-of-range class id yields a NaN pose for that object."""
        B, N = pcl.shape[0], self.n_pts
        C = prior_table.shape[0]
        pcl = self._dev(pcl, (B, N, 3), "pcl")
        prior_table = self._dev(prior_table, (C, self.n_prior, 3), "prior_table")
        if not prior_cls.is_cuda or prior_cls.dtype != torch.int32 or tuple(prior_cls.shape) != (B,):
            raise CatreError(f"prior_cls must be a CUDA int32 tensor of shape ({B},)")
        prior_cls = prior_cls.contiguous()
        init_pose = self._dev(init_pose, (B, 3, 4), "init_pose")
        init_scale = self._dev(init_scale, (B, 3), "init_scale")
        K = self._dev(K, (B, 3, 3), "K")
        if out is None:
            poses = torch.empty((n_iter + 1, B, 3, 4), dtype=torch.float32, device=pcl.device)
            scales = torch.empty((n_iter + 1, B, 3), dtype=torch.float32, device=pcl.device)
        else:
            poses, scales = out
        self._check(self.lib.catre_refine_table(self._h, _ptr(pcl), _ptr(prior_table), _ptr(prior_cls), C, _ptr(init_pose),
                                                _ptr(init_scale), _ptr(K), B, n_iter, _ptr(poses), _ptr(scales),
                                                self._stream()), "catre_refine_table")
        return poses, scales

    def refine_table_host(self, pcl, prior_table, prior_cls, init_pose, init_scale, K, n_iter: int, out=None):
        """Host-buffer form of refine_table (class ids are validated on the host)."""
        B, N = pcl.shape[0], self.n_pts
        C = prior_table.shape[0]
        for name, t, shp, dt in (("pcl", pcl, (B, N, 3), torch.float32), ("prior_table", prior_table, (C, self.n_prior, 3), torch.float32),
                                 ("prior_cls", prior_cls, (B,), torch.int32), ("init_pose", init_pose, (B, 3, 4), torch.float32),
                                 ("init_scale", init_scale, (B, 3), torch.float32), ("K", K, (B, 3, 3), torch.float32)):
            if t.is_cuda or t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous():
                raise CatreError(f"{name}: expected contiguous {dt} host tensor of shape {shp}")
        if out is None:
            poses = torch.empty((n_iter + 1, B, 3, 4), dtype=torch.float32).pin_memory()
            scales = torch.empty((n_iter + 1, B, 3), dtype=torch.float32).pin_memory()
        else:
            poses, scales = out
        self._check(self.lib.catre_refine_table_host(self._h, _ptr(pcl), _ptr(prior_table), _ptr(prior_cls), C,
                                                     _ptr(init_pose), _ptr(init_scale), _ptr(K), B, n_iter, _ptr(poses),
                                                     _ptr(scales), self._stream()), "catre_refine_table_host")
        return poses, scales

    # ---- training step ---------------------------------------------------------
    def train_set_weight(self, name: str, t: torch.Tensor):
        """Device-to-device refresh of the engine's fp32 copy of one checkpoint tensor (after an optimiser step)."""
        if not t.is_cuda or t.device.index != self.device or t.dtype != torch.float32:
            raise CatreError(f"{name}: expected a float32 CUDA tensor on device {self.device}")
        t = t.detach().contiguous()
        self._check(self.lib.catre_train_set_weight(self._h, name.encode(), t.data_ptr(), self._stream()), f"train_set_weight({name})")

    def train_set_weights(self, tensors):
        """The same refresh for many tensors in ONE launch: {checkpoint name: contiguous float32 CUDA tensor}; names that are
        absent keep their current values."""
        if getattr(self, "_names", None) is None:
            self._names = self.weight_names()
        names = self._names
        table = (ctypes.c_void_p * len(names))()
        keep = []
        for i, name in enumerate(names):
            t = tensors.get(name)
            if t is None:
                continue
            if not t.is_cuda or t.device.index != self.device or t.dtype != torch.float32:
                raise CatreError(f"{name}: expected a float32 CUDA tensor on device {self.device}")
            t = t.detach().contiguous()
            keep.append(t)  # alive until the launch is enqueued (stream-ordered afterwards)
            table[i] = t.data_ptr()
        unknown = set(tensors) - set(names)
        if unknown:
            raise CatreError(f"unknown checkpoint tensors {sorted(unknown)[:3]}")
        self._check(self.lib.catre_train_set_weights(self._h, table, self._stream()), "train_set_weights")

    def train_set_loss_weights(self, pm_lw: float = 1.0, rot_lw: float = 1.0, trans_lw: float = 1.0, scale_lw: float = 1.0):
        """LOSS_CFG.PM_LW / ROT_LW / TRANS_LW / SCALE_LW (all > 0)."""
        self._check(self.lib.catre_train_set_loss_weights(self._h, pm_lw, rot_lw, trans_lw, scale_lw), "train_set_loss_weights")

    def train_step(self, x_pm, tfd_pm, obj_kps, pose, scale, K, gt_pose, gt_scale, is_sym, sym_rots):
        """Forward + shipped losses + backward of one refinement iteration.  Device fp32 tensors except
        is_sym (B bools / 0-1 on the host) and sym_rots ([n,3,3] fp32 numpy / CPU tensor, may be empty).
        Returns (pose [B,3,4], scale [B,3], losses [6] on the device in the order of TRAIN_LOSS_NAMES); the
        parameter gradients stay in the engine (train_grad)."""
        import numpy as np

        B, N = x_pm.shape[0], self.n_pts
        x_pm = self._dev(x_pm, (B, N, 3), "x")
        tfd_pm = self._dev(tfd_pm, (B, N, 3), "tfd_kps")
        obj_kps = self._dev(obj_kps, (B, N, 3), "obj_kps")
        pose = self._dev(pose, (B, 3, 4), "init_pose")
        scale = self._dev(scale, (B, 3), "init_scale")
        K = self._dev(K, (B, 3, 3), "K")
        gt_pose = self._dev(gt_pose, (B, 3, 4), "gt_pose")
        gt_scale = self._dev(gt_scale, (B, 3), "gt_scale")
        sym = np.ascontiguousarray(np.asarray(is_sym).astype(np.uint8).reshape(-1))
        if sym.shape[0] != B:
            raise CatreError(f"is_sym has {sym.shape[0]} entries, expected {B}")
        rots = np.ascontiguousarray(np.asarray(sym_rots, dtype=np.float32).reshape(-1, 3, 3))
        op = torch.empty((B, 3, 4), dtype=torch.float32, device=x_pm.device)
        os_ = torch.empty((B, 3), dtype=torch.float32, device=x_pm.device)
        losses = torch.empty((6,), dtype=torch.float32, device=x_pm.device)
        self._check(self.lib.catre_train_step(self._h, _ptr(x_pm), _ptr(tfd_pm), _ptr(obj_kps), _ptr(pose), _ptr(scale), _ptr(K),
                                              _ptr(gt_pose), _ptr(gt_scale), sym.ctypes.data, rots.ctypes.data if len(rots) else None,
                                              len(rots), B, _ptr(op), _ptr(os_), _ptr(losses), self._stream()), "catre_train_step")
        return op, os_, losses

    def train_grad(self, name: str, out: torch.Tensor) -> torch.Tensor:
        """Copy the gradient of checkpoint tensor `name` from the last train_step into `out` (device fp32, same numel)."""
        if not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous():
            raise CatreError(f"{name}: gradient destination must be a contiguous float32 CUDA tensor")
        self._check(self.lib.catre_train_grad(self._h, name.encode(), out.data_ptr(), self._stream()), f"train_grad({name})")
        return out

    def train_grad_layout(self):
        """({checkpoint name: offset in floats}, arena size in floats) of the engine's gradient arena."""
        n = self.lib.catre_num_weights()
        offs, total = (ctypes.c_int64 * n)(), ctypes.c_int64()
        self._check(self.lib.catre_train_grad_layout(self._h, offs, ctypes.byref(total)), "train_grad_layout")
        return {name: int(offs[i]) for i, name in enumerate(self.weight_names())}, int(total.value)

    def train_grads_flat(self, scale: float = 1.0) -> torch.Tensor:
        """scale * (all gradients of the last train_step) as one flat device tensor laid out like train_grad_layout()."""
        _, total = self._grad_layout()
        out = torch.empty((total,), dtype=torch.float32, device=torch.device("cuda", self.device))
        self._check(self.lib.catre_train_grads_flat(self._h, out.data_ptr(), float(scale), self._stream()), "train_grads_flat")
        return out

    def _grad_layout(self):
        if getattr(self, "_layout", None) is None:
            self._layout = self.train_grad_layout()
        return self._layout

    # ---- accounting ------------------------------------------------------------------------------
    def last_launch_count(self) -> int:
        return int(self.lib.catre_last_launch_count(self._h))

    def debug_read(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """Tests only: copy an internal workspace buffer of the last launch to a CPU tensor."""
        out = torch.empty(shape, dtype=dtype)
        self._check(self.lib.catre_debug_read(self._h, name.encode(), out.data_ptr(), out.numel() * out.element_size()),
                    f"debug_read({name})")
        return out

    def workspace_bytes(self, B: int) -> int:
        return int(self.lib.catre_workspace_bytes(self._h, B))

    def profile_enable(self, on: bool):
        self._check(self.lib.catre_profile_enable(self._h, 1 if on else 0), "profile_enable")

    def profile_reset(self):
        self._check(self.lib.catre_profile_reset(self._h), "profile_reset")

    def profile(self) -> Dict[str, Tuple[float, int]]:
        """{kernel group: (total device ms, launches)} since the last reset."""
        out = {}
        for i in range(self.lib.catre_profile_num()):
            ms, n = ctypes.c_double(), ctypes.c_int64()
            self._check(self.lib.catre_profile_get(self._h, i, ctypes.byref(ms), ctypes.byref(n)), "profile_get")
            if n.value:
                out[self.lib.catre_profile_name(i).decode()] = (ms.value, n.value)
        return out

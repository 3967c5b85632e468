"""Fused Ranger optimiser step: a ``torch.optim.Optimizer`` with the constructor, parameter groups
and state-dict entries (``step``, ``exp_avg``, ``exp_avg_sq``, ``slow_buffer``) of the reference's ``Ranger``
(lib/torch_utils/solver/ranger.py:31-200 -- the optimiser the shipped config names, configs/.../aug05_..._120e.py:49),
whose ``step()`` is two CUDA launches over all parameter tensors (``catre_ranger_step``, csrc/optim_kernels.cuh) instead
of ~10 small launches per tensor.  ``nan_to_num=True`` folds in the gradient guard the reference's training loop runs
before the step (core/catre/engine/engine.py:349-352).

The state lives in torch tensors, so ``state_dict()`` / ``load_state_dict()`` interoperate with checkpoints written by
the reference's Ranger.  No CPU path: parameters must be CUDA tensors.
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, Optional

import torch
from torch.optim.optimizer import Optimizer

from . import engine as _engine


class RangerArgs(ctypes.Structure):  # catre_ranger_args (include/catre_b200.h)
    _fields_ = [("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float), ("one_minus_beta1", ctypes.c_float),
                ("one_minus_beta2", ctypes.c_float), ("step_size", ctypes.c_float),
                ("rectified", ctypes.c_int32), ("alpha", ctypes.c_float), ("lookahead", ctypes.c_int32), ("nan_to_num", ctypes.c_int32)]


def radam_step_size(step: int, beta1: float, beta2: float, threshold: float):
    """(rectified, step_size) for this step count: ranger.py:160-178."""
    beta2_t = beta2 ** step
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma > threshold:
        return True, math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max / (n_sma_max - 2)) / (
            1 - beta1 ** step)
    return False, 1.0 / (1 - beta1 ** step)


class FusedRanger(Optimizer):
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(0.95, 0.999), eps=1e-5, weight_decay=0,
                 use_gc=True, gc_conv_only=False, nan_to_num=False, step_fn: Optional[Callable] = None):
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f"Invalid slow update rate: {alpha}")
        if not 1 <= k:
            raise ValueError(f"Invalid lookahead steps: {k}")
        if not lr > 0:
            raise ValueError(f"Invalid Learning Rate: {lr}")
        if not eps > 0:
            raise ValueError(f"Invalid eps: {eps}")
        defaults = dict(lr=lr, alpha=alpha, k=k, step_counter=0, betas=betas, N_sma_threshhold=N_sma_threshhold, eps=eps,
                        weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.N_sma_threshhold, self.alpha, self.k = N_sma_threshhold, alpha, k
        self.use_gc = use_gc
        self.gc_gradient_threshold = 3 if gc_conv_only else 1
        self.nan_to_num = bool(nan_to_num)
        self._step_fn = step_fn  # tests drive the CPU emulation of the kernels through this; None = libcatre_b200.so
        self._shape_cache: dict = {}

    def _launch(self, group, ps):
        dev = ps[0].device
        if self._step_fn is None and not all(p.is_cuda for p in ps):
            raise _engine.CatreError("catre_b200.optim.FusedRanger runs on CUDA only; there is no CPU path")
        step = self.state[ps[0]]["step"] + 1
        rows, table, lr_wd = [], [], []
        for p in ps:
            st = self.state[p]
            if st["step"] + 1 != step:
                raise NotImplementedError("FusedRanger: parameters of one group at different step counts")
            g = p.grad
            for name, t in (("parameter", p), ("gradient", g), ("exp_avg", st["exp_avg"]), ("exp_avg_sq", st["exp_avg_sq"]),
                            ("slow_buffer", st["slow_buffer"])):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                    raise NotImplementedError(f"FusedRanger: {name} must be a contiguous float32 tensor on {dev}")
            # like the reference's step() (lib/torch_utils/solver/ranger.py:146-148): centralise whenever the gradient has more
            # dims than the threshold -- use_gc only chooses that threshold in the constructor, it does not gate the step
            gc = g.dim() > self.gc_gradient_threshold
            row_len = p.numel() // p.shape[0] if gc else 0
            rows.append(p.shape[0] if gc else 0)
            table.append([p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["slow_buffer"].data_ptr(),
                          p.numel(), row_len, 0])
            lr_wd.append([group["lr"], group["weight_decay"]])
        key = (id(group), tuple(p.numel() for p in ps), tuple(rows), str(dev))
        cached = self._shape_cache.get(key)
        if cached is None:  # prefix sums and the row-mean scratch depend on the shapes only
            es = torch.tensor([0] + [p.numel() for p in ps], dtype=torch.int64).cumsum(0)
            rs = torch.tensor([0] + rows, dtype=torch.int64).cumsum(0)
            cached = (es.to(dev), rs.to(dev), torch.empty(max(int(rs[-1]), 1), dtype=torch.float32, device=dev), int(es[-1]), int(rs[-1]))
            self._shape_cache[key] = cached
        es_d, rs_d, scratch, total, total_rows = cached
        table_d = torch.tensor(table, dtype=torch.int64).to(dev)  # the gradient tensors are new every step
        lr_wd_d = torch.tensor(lr_wd, dtype=torch.float32).to(dev)
        beta1, beta2 = group["betas"]
        rect, step_size = radam_step_size(step, beta1, beta2, self.N_sma_threshhold)
        args = RangerArgs(beta1, beta2, group["eps"], 1 - beta1, 1 - beta2, step_size, int(rect), self.alpha, int(step % group["k"] == 0), int(self.nan_to_num))
        if self._step_fn is None:
            fn = _engine.load_library().catre_ranger_step
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        else:
            fn, stream = self._step_fn, None
        rc = fn(ctypes.c_void_p(table_d.data_ptr()), ctypes.c_void_p(lr_wd_d.data_ptr()), ctypes.c_void_p(es_d.data_ptr()),
                ctypes.c_void_p(rs_d.data_ptr()), len(ps), total, total_rows, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(args), stream)
        if rc != 0:
            raise _engine.CatreError(f"catre_ranger_step failed ({rc})")
        for p in ps:
            self.state[p]["step"] = step
        torch._foreach_add_(ps, 0.0)  # the kernel wrote the parameters behind torch's back: bump their version counters
        return table_d, lr_wd_d  # kept alive by the caller until the next step (stream-ordered use)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        keep = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if p.grad.is_sparse:
                    raise RuntimeError("Ranger optimizer does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:  # ranger.py:128-137
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["slow_buffer"] = p.detach().clone(memory_format=torch.contiguous_format)
            # parameters whose step counts differ (e.g. unfrozen later) go in separate launches
            by_step: dict = {}
            for p in ps:
                by_step.setdefault(self.state[p]["step"], []).append(p)
            for _, sub in sorted(by_step.items()):
                keep.append(self._launch(group, sub))
        self._keep = keep
        return loss

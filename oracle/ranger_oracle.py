"""CPU oracle of the Ranger step (RAdam + Lookahead + gradient centralisation) the shipped config trains with.

TEST INFRASTRUCTURE ONLY.  Restates lib/torch_utils/solver/ranger.py:102-200 (and the gradient NaN guard of
core/catre/engine/engine.py:349-352) on numpy arrays in float32; pinned to ``tests/golden/golden_ranger.npz``, which the
unmodified reference optimiser produced (``tests/golden/make_golden_ranger.py``), by ``tests/test_optim.py``."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


def radam_scalars(step: int, beta1: float, beta2: float, n_sma_threshold: float = 5.0) -> Tuple[bool, float]:
    """(rectified, step_size) of ranger.py:160-178 -- Python float arithmetic, like the reference."""
    beta2_t = beta2 ** step
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma > n_sma_threshold:
        return True, math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max / (n_sma_max - 2)) / (
            1 - beta1 ** step)
    return False, 1.0 / (1 - beta1 ** step)


class RangerOracle:
    def __init__(self, params: List[np.ndarray], lrs: List[float], wds: List[float], alpha=0.5, k=6, betas=(0.95, 0.999), eps=1e-5,
                 gc_min_dims: int = 1):
        self.p = [p.astype(np.float32).copy() for p in params]
        self.lr, self.wd = lrs, wds
        self.alpha, self.k, self.betas, self.eps, self.gc_min_dims = alpha, k, betas, eps, gc_min_dims
        self.m = [np.zeros_like(p) for p in self.p]
        self.v = [np.zeros_like(p) for p in self.p]
        self.slow = [p.copy() for p in self.p]  # ranger.py:136-137: initialised from the parameter at the first step
        self.step_count = 0

    def step(self, grads: List[np.ndarray], nan_to_num: bool = True) -> None:
        self.step_count += 1
        b1, b2 = np.float32(self.betas[0]), np.float32(self.betas[1])
        rect, step_size = radam_scalars(self.step_count, self.betas[0], self.betas[1])
        for i, g in enumerate(grads):
            g = g.astype(np.float32).copy()
            if nan_to_num:
                g = np.nan_to_num(g, nan=0.0, posinf=1e5, neginf=-1e5).astype(np.float32)
            if g.ndim > self.gc_min_dims:  # gradient centralisation over everything but the first dimension (:147-148)
                g = g - g.mean(axis=tuple(range(1, g.ndim)), keepdims=True, dtype=np.float32)
            self.v[i] = self.v[i] * b2 + np.float32(1 - self.betas[1]) * g * g  # torch rounds the Python scalar 1 - beta2 once
            self.m[i] = self.m[i] * b1 + np.float32(1 - self.betas[0]) * g
            p = self.p[i]
            if self.wd[i] != 0:
                p = p + p * np.float32(-self.wd[i] * self.lr[i])
            if rect:
                p = p + np.float32(-step_size * self.lr[i]) * (self.m[i] / (np.sqrt(self.v[i]) + np.float32(self.eps)))
            else:
                p = p + np.float32(-step_size * self.lr[i]) * self.m[i]
            if self.step_count % self.k == 0:
                self.slow[i] = self.slow[i] + np.float32(self.alpha) * (p - self.slow[i])
                p = self.slow[i].copy()
            self.p[i] = p.astype(np.float32)

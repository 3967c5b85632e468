"""Three whole training iterations on the CPU: the drop-in's training forward (real ``dropin.CatreB200`` + autograd bridge) on
the emulated kernel chain, the reference loop's NaN guard, and ``optim.FusedRanger`` on the emulated optimiser kernels --
against the weights and losses the UNMODIFIED reference model + the reference's own Ranger produce with the same loop
(tests/golden/make_golden_train_loop.py).  Integration check of SURVEY.md 8(f) N4 end to end without a GPU; the engine
adapter below exists only in this test."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from catre_b200 import dropin, engine, optim, synth
from oracle import catre_oracle as co
from tests.test_optim import emu_step  # noqa: F401  (fixture: emulated catre_ranger_step)
from tests.test_train_emu import NAMES, emu, _ptr  # noqa: F401  (fixture: emulated training chain)
from tests.test_train_gpu import sample_positions

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_train_loop.npz")
N_PTS, BATCH, SEED, N_STEPS, LR = 128, 3, 31, 3, 1e-3


class EmuEngine:
    """The Engine methods the drop-in's training path calls, on the CPU emulation of the kernel chain."""
    device = 0

    def __init__(self, lib, weights):
        self.lib, self.w = lib, {k: np.ascontiguousarray(v.numpy(), dtype=np.float32) for k, v in weights.items()}
        self.g = {k: np.zeros_like(v) for k, v in self.w.items()}
        self.loss_w = None

    def train_set_weight(self, name, t):
        self.w[name][...] = t.detach().numpy()

    def train_set_loss_weights(self, *w):
        self.loss_w = np.asarray(w, dtype=np.float32)

    def train_step(self, x_pm, tfd_pm, obj_kps, pose, scale, K, gt_pose, gt_scale, is_sym, sym_rots):
        B, N = x_pm.shape[0], x_pm.shape[1]
        f = lambda t: np.ascontiguousarray(t.detach().numpy(), dtype=np.float32)
        arrs = [f(x_pm), f(obj_kps), f(pose), f(scale), f(K), f(gt_pose), f(gt_scale)]
        x_np, tfd_np = f(x_pm), f(tfd_pm)
        sym = np.ascontiguousarray(np.asarray(is_sym).astype(np.uint8))
        rots = np.ascontiguousarray(np.asarray(sym_rots, dtype=np.float32).reshape(-1, 3, 3))
        wp = (ctypes.c_void_p * 74)(*[_ptr(self.w[k]) for k in NAMES])
        gp = (ctypes.c_void_p * 74)(*[_ptr(self.g[k]) for k in NAMES])
        po, so, lo = np.zeros((B, 3, 4), np.float32), np.zeros((B, 3), np.float32), np.zeros(6, np.float32)
        rc = self.lib.emu_train_step(wp, B, N, *[_ptr(a) for a in arrs], _ptr(sym), _ptr(rots), len(rots), _ptr(po), _ptr(so), _ptr(lo),
                                     gp, None, _ptr(x_np), _ptr(tfd_np), None, None if self.loss_w is None else _ptr(self.loss_w))
        assert rc == 0
        return torch.from_numpy(po), torch.from_numpy(so), torch.from_numpy(lo)

    def train_grad(self, name, out):
        return out.copy_(torch.from_numpy(self.g[name]).reshape(out.shape))

    def _grad_layout(self):  # an arena in checkpoint order with padding between the tensors, like the engine's
        offs, o = {}, 0
        for k in NAMES:
            offs[k] = o
            o += (self.g[k].size + 63) // 64 * 64
        return offs, o

    def train_grads_flat(self, scale=1.0):
        offs, total = self._grad_layout()
        flat = torch.full((total,), float("nan"))
        for k in NAMES:
            flat[offs[k]: offs[k] + self.g[k].size] = scale * torch.from_numpy(self.g[k]).flatten()
        return flat


def _refresh(m):
    for n, p in m.named_parameters():
        cur = (p._version, p.data_ptr())
        if m._train_versions.get(n) != cur:
            m._engine.train_set_weight(n, p.data)
            m._train_versions[n] = cur
    return m._engine


@pytest.mark.parametrize("flat_grads", ["0", "1"])
def test_three_training_iterations_match_the_reference_loop(emu, emu_step, monkeypatch, flat_grads):
    monkeypatch.setenv("CATRE_TRAIN_FLAT_GRADS", flat_grads)  # 1 = gradients as views of one copy of the arena (opt-in)
    z = np.load(GOLDEN)
    w = co.resize_conv_p(synth.load_weights(), N_PTS)
    model = dropin.CatreB200(N_PTS, N_PTS, max_batch=4)
    model.load_state_dict(w, strict=True)
    model.train()
    model._engine = EmuEngine(emu, w)
    model._train_versions = {n: (p._version, p.data_ptr()) for n, p in model.named_parameters()}
    monkeypatch.setattr(dropin.CatreB200, "_engine_for_training", lambda self, device: _refresh(self))
    groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)], "lr": LR} for pre in ("pcl_net.", "rot_head.", "ts_head.")]
    opt = optim.FusedRanger(groups, lr=LR, weight_decay=0, step_fn=emu_step)
    batch, tgt = synth.make_train_batch(BATCH, N_PTS, SEED, round_robin_cls=True)
    from oracle import train_oracle as to

    rots = to.y_symmetry_rotations()
    sym_info = [rots if s else None for s in tgt.sym_y]
    x, tfd = co.update_points(batch.pcl, batch.prior, batch.init_pose, batch.init_scale)  # what batch_updater hands the model
    for it in range(N_STEPS):
        out_dict, loss_dict = model(x, tfd, init_pose=batch.init_pose, init_scale=batch.init_scale, K_zoom=batch.K, obj_class=batch.obj_cls,
                                    gt_ego_rot=tgt.gt_pose[:, :3, :3], gt_trans=tgt.gt_pose[:, :3, 3], gt_scale=tgt.gt_scale,
                                    obj_kps=batch.prior, sym_info=sym_info, do_loss=True, cur_iter=1)
        losses = sum(loss_dict.values())
        losses.backward()
        for p in model.parameters():  # the loop's NaN guard (engine.py:349-352)
            if p.grad is not None:
                torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
        opt.step()
        opt.zero_grad(set_to_none=True)
        assert abs(float(losses.detach()) - float(z[f"loss{it + 1}"])) <= 2e-4 * abs(float(z[f"loss{it + 1}"])), (it, float(losses.detach()))
    worst = 0.0
    for name, p in model.named_parameters():
        got = p.detach().double().flatten().numpy()
        delta = float(z[f"delta/{name}"])
        if name in dropin.UNUSED_PARAMS:
            assert delta == 0.0 and np.array_equal(p.detach().numpy(), w[name].numpy())  # never touched, in the reference either
            continue
        if f"weight/{name}/full" in z.files:
            err = np.abs(got - z[f"weight/{name}/full"]).max()
        else:
            err = np.abs(got[sample_positions(name, got.size)] - z[f"weight/{name}/samples"]).max()
        assert delta > 0 and err <= 0.02 * delta + 2e-8, (name, err, delta)
        worst = max(worst, err / delta)
    print(f"worst weight error relative to the tensor's change over 3 iterations: {worst:.2e}")
    assert worst < 0.02

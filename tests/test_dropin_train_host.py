"""Host-side plumbing of the drop-in's training forward (autograd bridge, loss-dict keys, weight refresh bookkeeping) with a
stub in place of the CUDA engine.  The stub exists only in this test: the product has no CPU path (test_cabi.py checks that
it raises), the arithmetic is covered by test_train_emu.py (CPU emulation of the kernels) and test_train_gpu.py."""
import numpy as np
import pytest
import torch

from catre_b200 import dropin, engine, synth


class StubEngine:
    device = 0

    def __init__(self):
        self.refreshed, self.steps = [], 0

    def train_set_weight(self, name, t):
        self.refreshed.append(name)

    def train_step(self, x_pm, tfd_pm, obj_kps, pose, scale, K, gt_pose, gt_scale, is_sym, sym_rots):
        self.steps += 1
        self.last = dict(x=x_pm, gt_pose=gt_pose, is_sym=list(is_sym), n_rots=len(sym_rots))
        return pose + 1.0, scale + 1.0, torch.arange(1.0, 7.0)

    def train_grad(self, name, out):
        return out.fill_(float(len(name)))


@pytest.fixture()
def model(monkeypatch):
    m = dropin.CatreB200(64, 64, max_batch=4)
    stub = StubEngine()
    m._engine = stub
    m._train_versions = {n: (p._version, p.data_ptr()) for n, p in m.named_parameters()}
    monkeypatch.setattr(dropin.CatreB200, "_engine_for_training", lambda self, device: _refresh(self))
    return m, stub


def _refresh(m):  # the bookkeeping half of _engine_for_training, without the CUDA device checks
    for n, p in m.named_parameters():
        cur = (p._version, p.data_ptr())
        if m._train_versions.get(n) != cur:
            m._engine.train_set_weight(n, p.data)
            m._train_versions[n] = cur
    return m._engine


def call(m, sym_info):
    B = len(sym_info)
    x = torch.randn(B, 64, 3).permute(0, 2, 1)
    pose = torch.cat((torch.eye(3).expand(B, 3, 3), torch.ones(B, 3, 1)), 2)
    return m(x, x, init_pose=pose, init_scale=torch.ones(B, 3), K_zoom=torch.eye(3).expand(B, 3, 3), gt_ego_rot=pose[:, :, :3],
             gt_trans=pose[:, :, 3], gt_scale=torch.ones(B, 3), obj_kps=torch.randn(B, 64, 3), sym_info=sym_info, do_loss=True,
             cur_iter=2)


def test_loss_dict_keys_and_gradients(model):
    m, stub = model
    rots = np.stack([np.eye(3, dtype=np.float32)] * 5)
    out, loss = call(m, [None, rots, None])
    assert set(out) == {"pose_2", "scale_2"} and not out["pose_2"].requires_grad
    assert list(loss) == list(engine.TRAIN_LOSS_NAMES) and float(loss["loss_scale"].detach()) == 6.0
    assert stub.last["is_sym"] == [False, True, False] and stub.last["n_rots"] == 5 and stub.last["x"].shape == (3, 64, 3)
    assert stub.last["gt_pose"].shape == (3, 3, 4)
    (3.0 * sum(loss.values())).backward()  # uniform factor (AMP loss scale)
    for name, p in m.named_parameters():
        if name in dropin.UNUSED_PARAMS:
            assert p.grad is None
        else:
            assert p.grad.shape == p.shape and bool((p.grad == 3.0 * len(name)).all()), name
    # the reference omits loss_rot when every object is symmetric and loss_yaxis_rot when none is
    assert "loss_rot" not in call(m, [rots, rots])[1] and "loss_yaxis_rot" in call(m, [rots, rots])[1]
    only_asym = call(m, [None, None])[1]
    assert "loss_yaxis_rot" not in only_asym and "loss_rot" in only_asym
    sum(only_asym.values()).backward()  # grad_out of the absent entry is 0 and must not count as a per-term weight


def test_per_term_weights_and_stale_backward_are_refused(model):
    m, _ = model
    _, loss = call(m, [None])
    with pytest.raises(NotImplementedError):
        (loss["loss_PM_R"] + 2.0 * loss["loss_scale"]).backward()
    _, old = call(m, [None])
    call(m, [None])
    with pytest.raises(RuntimeError):
        sum(old.values()).backward()


def test_only_changed_tensors_are_refreshed(model):
    m, stub = model
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    _, loss = call(m, [None])
    assert stub.refreshed == []
    sum(loss.values()).backward()
    opt.step()
    call(m, [None])
    assert len(stub.refreshed) == 68 and not set(stub.refreshed) & set(dropin.UNUSED_PARAMS)

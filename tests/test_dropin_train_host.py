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

    def train_set_loss_weights(self, *w):
        self.loss_w = w

    def train_step(self, x_pm, tfd_pm, obj_kps, pose, scale, K, gt_pose, gt_scale, is_sym, sym_rots):
        self.steps += 1
        self.last = dict(x=x_pm, gt_pose=gt_pose, is_sym=list(is_sym), n_rots=len(sym_rots))
        return pose + 1.0, scale + 1.0, torch.arange(1.0, 7.0)

    def train_grad(self, name, out):
        return out.fill_(float(len(name)))

    # the default hand-off since round 2: one scaled copy of the gradient arena, handed out as views
    def _grad_layout(self):
        offs, o = {}, 0
        for name, shape in dropin.param_specs(64, 64):
            offs[name] = o
            o += int(torch.tensor(shape).prod())
        return offs, o

    def train_grads_flat(self, scale=1.0):
        offs, total = self._grad_layout()
        flat = torch.empty(total)
        for name, shape in dropin.param_specs(64, 64):
            n = int(torch.tensor(shape).prod())
            flat[offs[name]: offs[name] + n] = float(len(name)) * scale
        return flat


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


def test_loss_weights_from_the_config_reach_the_engine(model):
    m, stub = model
    call(m, [None])
    assert not hasattr(stub, "loss_w")  # shipped weights: nothing to set
    m.cfg = {"MODEL": {"CATRE": {"LOSS_CFG": {"PM_LW": 0.5, "TRANS_LW": 3.0}}}}
    call(m, [None])
    assert stub.loss_w == (0.5, 1.0, 3.0, 1.0)


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


# ---- data-parallel training: the reference wraps the model in DistributedDataParallel(find_unused_parameters=True)
#      (core/catre/main_catre.py:154-160); world_size-2 gloo run of the drop-in under that wrapper, stub engine per rank
def _ddp_worker(rank, world, port, q):
    import os

    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class RankStub(StubEngine):
        def train_grad(self, name, out):
            return out.fill_(float(len(name)) * (rank + 1))  # rank-dependent gradients: DDP must average them

        def train_grads_flat(self, scale=1.0):
            return StubEngine.train_grads_flat(self, scale * (rank + 1))

    m = dropin.CatreB200(64, 64, max_batch=4)
    m._engine = RankStub()
    m._train_versions = {n: (p._version, p.data_ptr()) for n, p in m.named_parameters()}
    dropin.CatreB200._engine_for_training = lambda self, device: _refresh(self)
    ddp = DistributedDataParallel(m, broadcast_buffers=False, find_unused_parameters=True)
    ok = True
    for it in range(2):  # two iterations: the second forward fails if a reduction of the first never finished
        B = 2
        x = torch.randn(B, 64, 3).permute(0, 2, 1)
        pose = torch.cat((torch.eye(3).expand(B, 3, 3), torch.ones(B, 3, 1)), 2)
        _, loss = ddp(x, x, init_pose=pose, init_scale=torch.ones(B, 3), K_zoom=torch.eye(3).expand(B, 3, 3), gt_ego_rot=pose[:, :, :3],
                      gt_trans=pose[:, :, 3], gt_scale=torch.ones(B, 3), obj_kps=torch.randn(B, 64, 3), sym_info=[None] * B,
                      do_loss=True, cur_iter=it + 1)
        m.zero_grad(set_to_none=True)
        sum(loss.values()).backward()
        for name, p in m.named_parameters():
            if name in dropin.UNUSED_PARAMS:
                ok = ok and p.grad is None
            else:
                ok = ok and bool((p.grad == len(name) * 1.5).all())  # mean of (1, 2) x len(name)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_dropin_under_ddp_gloo_world2():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_vis_scalars_follow_the_reference_formulas():
    """vis/* scalars (CATRE_disR_shared.py:127-146): checked against the oracle's pose update and plain formulas."""
    from oracle import catre_oracle as co

    g = torch.Generator().manual_seed(3)
    b = synth.make_batch(3, 64, 5)
    d_t = torch.tensor([[3.0, -2.0, 1.01], [0.5, 0.2, 0.99], [-1.0, 4.0, 1.0]])
    rot, t, s = co.pose_update(torch.eye(3).expand(3, 3, 3), d_t, torch.zeros(3, 3), b.init_pose[:, :, :3], b.init_pose[:, :, 3],
                               b.init_scale, b.K)
    pose = torch.cat((rot, t.reshape(3, 3, 1)), dim=2)
    gt_rot = b.init_pose[:, :, :3]  # zero rotation error
    gt_t = b.init_pose[:, :, 3] + torch.randn(3, 3, generator=g) * 0.01
    v = dropin.vis_scalars(2, pose, b.init_pose, b.K, gt_rot, gt_t)
    assert len(v) == 14 and abs(v["vis/error_R_2"]) < 0.05
    assert abs(v["vis/error_t_2"] - 100 * float((gt_t - t).norm(dim=1).mean())) < 1e-4
    for a, ax in enumerate("xyz"):
        assert abs(v[f"vis/t{ax}_delta_2"] - float(d_t[0, a])) < 1e-3  # the head's raw deltas, recovered
        assert abs(v[f"vis/t{ax}_pred_2"] - float(t[0, a])) < 1e-7 and abs(v[f"vis/t{ax}_gt_2"] - float(gt_t[0, a])) < 1e-7
        assert abs(v[f"vis/error_t{ax}_2"] - 100 * abs(float(t[0, a] - gt_t[0, a]))) < 1e-5

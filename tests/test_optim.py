"""The fused Ranger step: oracle pinned to the reference's own optimiser (golden_ranger.npz), the CUDA
kernels' source run on the CPU (tests/emu/optim_emu.cpp) through catre_b200.optim.FusedRanger against the same goldens,
and -- on the GPU -- the real library."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_ranger import N_STEPS, make_inputs  # noqa: E402  (seeded inputs only)

from catre_b200 import optim  # noqa: E402
from oracle import ranger_oracle as ro  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden_ranger.npz")
LRS, WDS = [1e-2] * 3 + [3e-3] * 2, [0.0] * 3 + [0.1] * 2


def test_oracle_matches_reference_ranger():
    z = np.load(GOLDEN)
    params, grads = make_inputs()
    o = ro.RangerOracle([p.numpy() for p in params], LRS, WDS)
    for k in range(N_STEPS):
        o.step([g.numpy() for g in grads[k]])
        for i in range(len(params)):
            assert np.allclose(o.p[i], z[f"step{k + 1}_p{i}"], rtol=2e-6, atol=2e-7), (k, i)
    for i in range(len(params)):
        assert np.allclose(o.m[i], z[f"final_exp_avg{i}"], rtol=1e-5, atol=1e-8) and np.allclose(o.v[i], z[f"final_exp_avg_sq{i}"], rtol=1e-5, atol=1e-10) and np.allclose(o.slow[i], z[f"final_slow{i}"], rtol=2e-6, atol=2e-7)


@pytest.fixture(scope="module")
def emu_step(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "liboptim_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DCATRE_HOST_EMU", "-o", out, os.path.join(HERE, "emu", "optim_emu.cpp")])
    fn = ctypes.CDLL(out).catre_ranger_step
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return fn


def run_fused(device, step_fn):
    z = np.load(GOLDEN)
    params, grads = make_inputs()
    ps = [torch.nn.Parameter(p.clone().to(device)) for p in params]
    opt = optim.FusedRanger([{"params": ps[:3], "lr": 1e-2}, {"params": ps[3:], "lr": 3e-3, "weight_decay": 0.1}], lr=1e-2,
                            nan_to_num=True, step_fn=step_fn)
    for k in range(N_STEPS):
        versions = [p._version for p in ps]
        for p, g in zip(ps, grads[k]):
            p.grad = g.clone().to(device)
        opt.step()
        assert all(p._version > v for p, v in zip(ps, versions))  # what the drop-in's weight refresh keys on
        for i, p in enumerate(ps):
            assert np.allclose(p.detach().cpu().numpy(), z[f"step{k + 1}_p{i}"], rtol=2e-6, atol=2e-7), (k, i)
    sd = opt.state_dict()["state"]
    for i in range(len(ps)):  # the reference's state-dict entries
        assert sd[i]["step"] == N_STEPS
        assert np.allclose(sd[i]["exp_avg"].cpu().numpy(), z[f"final_exp_avg{i}"], rtol=1e-5, atol=1e-8)
        assert np.allclose(sd[i]["exp_avg_sq"].cpu().numpy(), z[f"final_exp_avg_sq{i}"], rtol=1e-5, atol=1e-10)
        assert np.allclose(sd[i]["slow_buffer"].cpu().numpy(), z[f"final_slow{i}"], rtol=2e-6, atol=2e-7)
    return opt, ps


def test_fused_ranger_on_emulated_kernels_matches_reference(emu_step):
    opt, ps = run_fused("cpu", emu_step)
    ps[1].grad = None  # a parameter without gradient is skipped and keeps its step count; the others move on in their own launch
    before = ps[1].detach().clone()
    for p in (ps[0], ps[2]):
        p.grad = torch.ones_like(p)
    opt.step()
    assert torch.equal(ps[1], before) and opt.state[ps[1]]["step"] == N_STEPS and opt.state[ps[0]]["step"] == N_STEPS + 1


def test_fused_ranger_refuses_cpu_without_the_library_path():
    p = torch.nn.Parameter(torch.zeros(4, 4))
    p.grad = torch.ones(4, 4)
    with pytest.raises(Exception):
        optim.FusedRanger([p]).step()


def test_fused_ranger_resumes_from_a_state_dict(emu_step):
    """Checkpoint / resume (the reference saves optimizer.state_dict() with its PeriodicCheckpointer): stop after 5 steps,
    load the state into a fresh optimiser over fresh parameter objects, continue -- same trajectory as the uninterrupted run."""
    z = np.load(GOLDEN)
    params, grads = make_inputs()

    def build(ps):
        return optim.FusedRanger([{"params": ps[:3], "lr": 1e-2}, {"params": ps[3:], "lr": 3e-3, "weight_decay": 0.1}], lr=1e-2,
                                 nan_to_num=True, step_fn=emu_step)

    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = build(ps)
    for k in range(5):
        for p, g in zip(ps, grads[k]):
            p.grad = g.clone()
        opt.step()
    saved = opt.state_dict()
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = build(ps2)
    opt2.load_state_dict(saved)
    for k in range(5, N_STEPS):
        for p, g in zip(ps2, grads[k]):
            p.grad = g.clone()
        opt2.step()
    for i, p in enumerate(ps2):
        assert np.allclose(p.detach().numpy(), z[f"step{N_STEPS}_p{i}"], rtol=2e-6, atol=2e-7), i

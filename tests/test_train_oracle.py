"""The training-step oracle against the golden losses / gradient digests produced by the
unmodified reference (tests/golden/make_golden_train.py), and the hand-derived backward against autograd."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from catre_b200 import synth
from oracle import train_oracle as to

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_train.npz")
N_SAMPLES = 256


def sample_positions(name, numel):  # same rule as make_golden_train.sample_positions
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31)
    return np.random.RandomState(seed).randint(0, numel, size=N_SAMPLES)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def inputs(golden):
    z = golden
    batch, tgt = synth.make_train_batch(6, 1024, 11, round_robin_cls=True)
    assert np.array_equal(batch.pcl.numpy(), z["pcl"]) and np.array_equal(tgt.gt_pose.numpy(), z["gt_pose"])
    sym_rots = to.y_symmetry_rotations()
    assert sym_rots.shape == z["sym_rots"].shape and np.abs(sym_rots - z["sym_rots"]).max() < 1e-6
    sym_info = [z["sym_rots"] if s else None for s in z["sym_y"]]
    return batch, tgt, sym_info


def check_grads(z, it, grads, rtol, atol_frac):
    names = sorted({k.split("/")[1] for k in z.files if k.startswith(f"it{it}_grad/")})
    assert len(names) == 68 and set(names) == set(grads.keys())
    for name in names:
        g = grads[name].double().flatten().numpy()
        if f"it{it}_grad/{name}/full" in z.files:
            want = z[f"it{it}_grad/{name}/full"]
            scale = max(np.abs(want).max(), 1e-12)
            assert np.abs(g - want).max() <= atol_frac * scale, name
        else:
            stats, samples = z[f"it{it}_grad/{name}/stats"], z[f"it{it}_grad/{name}/samples"]
            got = np.array([g.sum(), np.abs(g).sum(), np.sqrt((g * g).sum())])
            assert np.allclose(got[1:], stats[1:], rtol=rtol), (name, got, stats)
            assert abs(got[0] - stats[0]) <= rtol * stats[1], name
            scale = max(np.abs(samples).max(), 1e-12)
            assert np.abs(g[sample_positions(name, g.size)] - samples).max() <= atol_frac * scale, name


def test_losses_and_gradients_match_reference(golden, inputs):
    z = golden
    batch, tgt, sym_info = inputs
    w = synth.load_weights()
    pose, scale = batch.init_pose, batch.init_scale
    for it in (1, 2):
        pose, scale, losses, grads = to.train_step(w, batch.pcl, batch.prior, pose, scale, batch.K, tgt.gt_pose, tgt.gt_scale,
                                                   sym_info)
        assert sorted(losses.keys()) == list(z[f"it{it}_loss_names"])
        got = np.array([losses[k] for k in sorted(losses.keys())])
        assert np.allclose(got, z[f"it{it}_loss_values"], rtol=2e-5, atol=1e-7), (got, z[f"it{it}_loss_values"])
        assert np.abs(pose.numpy() - z[f"it{it}_pose"]).max() < 2e-6 and np.abs(scale.numpy() - z[f"it{it}_scale"]).max() < 2e-6
        check_grads(z, it, grads, rtol=2e-4, atol_frac=2e-3)


def test_manual_backward_equals_autograd(inputs):
    """The stage-by-stage backward the CUDA chain follows == autograd, in fp64 (algebra check, all 68 tensors)."""
    batch, tgt, sym_info = inputs
    w = {k: v.double() for k, v in synth.load_weights().items()}
    args = [t.double() for t in (batch.pcl[:4], batch.prior[:4], batch.init_pose[:4], batch.init_scale[:4], batch.K[:4],
                                 tgt.gt_pose[:4], tgt.gt_scale[:4])]
    sym = [None if s is None else s.astype(np.float64) for s in sym_info[:4]]
    pose, scale, losses, grads = to.train_step(w, *args, sym)
    sv, mgr = to.manual_train_step(w, *args, sym)
    assert (sv["rot"] - pose[:, :3, :3]).abs().max() < 1e-12 and (sv["s"] - scale).abs().max() < 1e-12
    assert set(mgr.keys()) == set(grads.keys())
    for k, g in grads.items():
        assert mgr[k].shape == g.shape, k
        assert (mgr[k] - g).abs().max() <= 1e-9 * max(g.abs().max().item(), 1e-6), (k, (mgr[k] - g).abs().max().item(), g.abs().max().item())

"""The training-step kernel chain (catre_b200/csrc/train_kernels.cuh + train_chain.cuh), compiled for the CPU with
CATRE_HOST_EMU (every kernel functor runs as loops over its grid), against the training oracle: checks the kernels'
indexing and the host orchestration without a GPU.  The emulation library is test infrastructure, built into a
temporary directory; the product library never contains it."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from catre_b200 import synth
from oracle import catre_oracle as co
from oracle import train_oracle as to

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = list(np.load(synth.WEIGHTS_NPZ).files)  # checkpoint order


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "libtrain_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DCATRE_HOST_EMU", "-o", out,
                           os.path.join(HERE, "emu", "train_emu.cpp")])
    return ctypes.CDLL(out)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_emu(lib, w, batch, tgt, sym_rots, pose, scale, reposed=False, loss_w=None):
    B, N = batch.pcl.shape[0], batch.pcl.shape[1]
    wa = [np.ascontiguousarray(w[k].numpy(), dtype=np.float32) for k in NAMES]
    ga = [np.zeros_like(a) for a in wa]
    wp = (ctypes.c_void_p * 74)(*[_ptr(a) for a in wa])
    gp = (ctypes.c_void_p * 74)(*[_ptr(a) for a in ga])
    f = lambda t: np.ascontiguousarray(t.numpy(), dtype=np.float32)
    arrs = [f(batch.pcl), f(batch.prior), f(pose), f(scale), f(batch.K), f(tgt.gt_pose), f(tgt.gt_scale)]
    is_sym = np.ascontiguousarray(tgt.sym_y.numpy().astype(np.uint8))
    rots = np.ascontiguousarray(sym_rots, dtype=np.float32)
    pose_out, scale_out = np.zeros((B, 3, 4), np.float32), np.zeros((B, 3), np.float32)
    losses, launches, macs = np.zeros(6, np.float32), ctypes.c_long(0), ctypes.c_double(0)
    x_pm = tfd_pm = None
    if reposed:  # the reference forward's inputs: x = pcl - t, tfd_kps = R (s * kps), here point-major
        x, tfd = co.update_points(batch.pcl, batch.prior, pose, scale)
        x_pm, tfd_pm = f(x.permute(0, 2, 1)), f(tfd.permute(0, 2, 1))
        arrs[0] = np.full_like(arrs[0], np.nan)  # the raw cloud must not be read on this route
    rc = lib.emu_train_step(wp, B, N, *[_ptr(a) for a in arrs], _ptr(is_sym), _ptr(rots), len(rots), _ptr(pose_out), _ptr(scale_out),
                            _ptr(losses), gp, ctypes.byref(launches), None if x_pm is None else _ptr(x_pm),
                            None if tfd_pm is None else _ptr(tfd_pm), ctypes.byref(macs),
                            None if loss_w is None else _ptr(np.asarray(loss_w, dtype=np.float32)))
    assert rc == 0
    run_emu.last_gemm_macs = macs.value
    return pose_out, scale_out, losses, dict(zip(NAMES, ga)), launches.value


LOSS_ORDER = ("loss_PM_R", "loss_rot", "loss_yaxis_rot", "loss_trans_xy", "loss_trans_z", "loss_scale")


# the last case has real (un-resized) weights and 4096 rows, which switches the weight-gradient GEMMs to split-K
# ... and two edge shapes: a single symmetric object with a point count that is no multiple of anything (N = 100: ragged
# tiles everywhere, one GroupNorm chunk), and an all-asymmetric batch with N = 130 (64 GroupNorm chunks, some empty)
@pytest.mark.parametrize("B,N,seed,reposed,sym", [(3, 64, 21, False, None), (2, 128, 22, True, None), (2, 1024, 23, False, None),
                                                  (1, 100, 24, False, None), (2, 130, 25, True, False)])
def test_emulated_chain_matches_oracle(emu, B, N, seed, reposed, sym):
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    w = co.resize_conv_p(synth.load_weights(), N)
    batch, tgt = synth.make_train_batch(B, N, seed, round_robin_cls=True)  # classes 0, 1, 2: symmetric and asymmetric
    if sym is not None:
        tgt.sym_y[:] = sym
    sym_rots = to.y_symmetry_rotations()
    sym_info = [sym_rots if s else None for s in tgt.sym_y]
    pose, scale = batch.init_pose, batch.init_scale
    # fp64 oracle: the fp32 torch oracle itself is only good to ~3e-3 on the STN gradients (arg-max near-ties)
    args64 = [t.double() for t in (batch.pcl, batch.prior, pose, scale, batch.K, tgt.gt_pose, tgt.gt_scale)]
    loss_w = (0.5, 2.0, 3.0, 0.25) if seed == 21 else None  # one case with LOSS_CFG weights other than 1
    p_ref, s_ref, l_ref, g_ref = to.train_step({k: v.double() for k, v in w.items()}, *args64,
                                               [None if r is None else r.astype(np.float64) for r in sym_info],
                                               loss_w or (1.0, 1.0, 1.0, 1.0))
    p, s, losses, grads, launches = run_emu(emu, w, batch, tgt, sym_rots, pose, scale, reposed, loss_w)
    assert launches > 100
    if N == 1024:  # the algorithmic work DESIGN.md quotes: 1.58 M multiply-adds per point (1.05 M forward + 0.53 M backward)
        per_point = run_emu.last_gemm_macs / (2 * B * N)
        assert 1.55e6 < per_point < 1.68e6, per_point
    assert np.abs(p - p_ref.numpy()).max() < 2e-5 and np.abs(s - s_ref.numpy()).max() < 2e-5
    for i, k in enumerate(LOSS_ORDER):
        assert abs(losses[i] - l_ref.get(k, 0.0)) <= 2e-5 * max(1.0, abs(l_ref.get(k, 0.0))), (k, losses[i], l_ref.get(k))
    for k in NAMES:
        if k in to.UNUSED:
            assert not grads[k].any(), k
            continue
        want = g_ref[k].numpy()
        scale_k = max(np.abs(want).max(), 1e-8)
        err = np.abs(grads[k] - want).max() / scale_k
        assert err < 2e-4, (k, err, scale_k)


def test_emulated_chain_under_address_sanitizer(tmp_path):
    """Every workspace slice followed by a poisoned red zone, ragged shapes: no kernel of the chain reads or writes out of range
    (the CPU stand-in for compute-sanitizer's memcheck, which needs a GPU)."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "train_asan")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address", "-DCATRE_HOST_EMU", "-DCATRE_EMU_ASAN",
                        os.path.join(HERE, "emu", "train_emu_asan_main.cpp"), "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("AddressSanitizer runtime not available: " + r.stderr[-200:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "no out-of-range access" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]

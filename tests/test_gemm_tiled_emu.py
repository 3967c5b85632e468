"""tk_gemm_tiled (the shared-memory GEMM every training layer goes through) compiled for the CPU from its CUDA source
(tests/emu/gemm_tiled_emu.cpp: one OS thread per CUDA thread, a barrier for __syncthreads) against KGemmNaive and numpy,
over the stride / batch / split-K / accumulate combinations train_chain.cuh uses.  Test infrastructure only."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
I64, I32, P = ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p


class GemmP(ctypes.Structure):  # catre_train::GemmP (train_kernels.cuh)
    _fields_ = [("A", P), ("sam", I64), ("sak", I64), ("sab", I64), ("B", P), ("sbk", I64), ("sbn", I64), ("sbb", I64),
                ("C", P), ("scm", I64), ("scn", I64), ("scb", I64), ("bias", P), ("sbias_b", I64),
                ("M", I32), ("N", I32), ("K", I32), ("relu", I32), ("accumulate", I32), ("splits", I32), ("k_per", I32),
                ("partial", P), ("f16", I32), ("bias_grad_partial", P)]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "libgemm_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DCATRE_HOST_EMU", "-o", out,
                           os.path.join(HERE, "emu", "gemm_tiled_emu.cpp")])
    so = ctypes.CDLL(out)
    assert so.emu_gemm_param_bytes() == ctypes.sizeof(GemmP)
    return so


def ptr(a, off=0):
    return ctypes.c_void_p(a.ctypes.data + 4 * off)


def run(lib, tiled, A, sa, B, sb, C, sc, M, N, K, bias=None, sbias_b=0, relu=0, acc=0, batch=1, splits=1, k_per=0, partial=None,
        a_off=0, b_off=0, c_off=0):
    p = GemmP(ptr(A, a_off), sa[0], sa[1], sa[2], ptr(B, b_off), sb[0], sb[1], sb[2], ptr(C, c_off), sc[0], sc[1], sc[2],
              None if bias is None else ptr(bias), sbias_b, M, N, K, relu, acc, splits, k_per or K,
              None if partial is None else ptr(partial))
    lib.emu_gemm(ctypes.byref(p), splits if splits > 1 else batch, tiled)


def both(lib, make, **kw):
    outs = []
    for tiled in (0, 1, 2):  # naive, tk_gemm_tiled, tk_gemm_tiled2
        arrs = make()
        run(lib, tiled, *arrs["args"], **kw, **arrs.get("kw", {}))
        outs.append(arrs["out"]().copy())
    assert np.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-5) and np.allclose(outs[0], outs[2], rtol=1e-5, atol=1e-5)
    return outs[2]


def test_layer_bias_relu_edges(lib):
    rng = np.random.RandomState(0)
    rows, C, K = 130, 70, 37
    x, W, b = (rng.randn(rows, K).astype(np.float32), rng.randn(C, K).astype(np.float32), rng.randn(C).astype(np.float32))

    def make():
        out = np.full((rows, C), 7.0, np.float32)
        return dict(args=(x, (K, 1, 0), W, (1, K, 0), out, (C, 1, 0), rows, C, K), kw=dict(bias=b, relu=1), out=lambda: out)

    got = both(lib, make)
    assert np.allclose(got, np.maximum(x @ W.T + b, 0), rtol=1e-4, atol=1e-4)


def test_weight_gradient_accumulate_offset_and_split_k(lib):
    rng = np.random.RandomState(1)
    rows, C, Kin, ldy, ldw, woff = 300, 50, 20, 53, 33, 9
    dy, x = rng.randn(rows, ldy).astype(np.float32), rng.randn(rows, Kin).astype(np.float32)
    base = rng.randn(C, ldw).astype(np.float32)

    def make():
        g = base.copy()
        return dict(args=(dy, (1, ldy, 0), x, (Kin, 1, 0), g, (ldw, 1, 0), C, Kin, rows), kw=dict(acc=1, c_off=woff), out=lambda: g)

    got = both(lib, make)
    want = base.copy()
    want[:, woff:woff + Kin] += dy[:, :C].T @ x
    assert np.allclose(got, want, rtol=1e-4, atol=1e-3)

    def make_split():
        part = np.zeros((3, C, Kin), np.float32)
        dummy = np.zeros((C, Kin), np.float32)
        return dict(args=(dy, (1, ldy, 0), x, (Kin, 1, 0), dummy, (Kin, 1, 0), C, Kin, rows),
                    kw=dict(splits=3, k_per=112, partial=part), out=lambda: part)

    part = both(lib, make_split)
    assert np.allclose(part.sum(0), dy[:, :C].T @ x, rtol=1e-4, atol=1e-3)
    assert np.allclose(part[2], dy[224:, :C].T @ x[224:], rtol=1e-4, atol=1e-3)  # the last slab is the short one


def test_input_gradient_with_column_offset(lib):
    rng = np.random.RandomState(2)
    rows, C, Kin, ldw, woff = 75, 40, 64, 90, 26
    dy, W = rng.randn(rows, C).astype(np.float32), rng.randn(C, ldw).astype(np.float32)
    prev = rng.randn(rows, Kin).astype(np.float32)

    def make():
        dx = prev.copy()
        return dict(args=(dy, (C, 1, 0), W, (ldw, 1, 0), dx, (Kin, 1, 0), rows, Kin, C), kw=dict(acc=1, b_off=woff), out=lambda: dx)

    assert np.allclose(both(lib, make), prev + dy @ W[:, woff:woff + Kin], rtol=1e-4, atol=1e-3)


def test_batched_per_set_operands(lib):
    rng = np.random.RandomState(3)
    S, n, K, C = 3, 70, 64, 100
    pf, W, cset = rng.randn(S, n, K).astype(np.float32), rng.randn(C, 150).astype(np.float32), rng.randn(S, C).astype(np.float32)

    def make():  # rotation layer 0: shared weight (column offset), per-set bias row
        out = np.zeros((S, n, C), np.float32)
        return dict(args=(pf, (K, 1, n * K), W, (1, 150, 0), out, (C, 1, n * C), n, C, K),
                    kw=dict(bias=cset, sbias_b=C, batch=S, b_off=50), out=lambda: out)

    assert np.allclose(both(lib, make), pf @ W[:, 50:50 + K].T + cset[:, None, :], rtol=1e-4, atol=1e-3)
    t = rng.randn(S, K, K).astype(np.float32)

    def make_t():  # dh1 = dpf . T64^T per set (B read transposed)
        out = np.zeros((S, n, K), np.float32)
        return dict(args=(pf, (K, 1, n * K), t, (1, K, K * K), out, (K, 1, n * K), n, K, K), kw=dict(batch=S), out=lambda: out)

    assert np.allclose(both(lib, make_t), pf @ t.transpose(0, 2, 1), rtol=1e-4, atol=1e-3)
    d = rng.randn(S, n, K).astype(np.float32)

    def make_g():  # dT64 = h1^T . dpf per set (A read transposed, reduction over the points)
        out = np.zeros((S, K, K), np.float32)
        return dict(args=(pf, (1, K, n * K), d, (K, 1, n * K), out, (K, 1, K * K), K, K, n), kw=dict(batch=S), out=lambda: out)

    assert np.allclose(both(lib, make_g), pf.transpose(0, 2, 1) @ d, rtol=1e-4, atol=1e-3)
    q, dq = rng.randn(S, n, 3).astype(np.float32), rng.randn(S, n, 3).astype(np.float32)

    def make_3():  # dT3 = q^T . dq' per set (3 x 3 outputs)
        out = np.zeros((S, 3, 3), np.float32)
        return dict(args=(q, (1, 3, n * 3), dq, (3, 1, n * 3), out, (3, 1, 9), 3, 3, n), kw=dict(batch=S), out=lambda: out)

    assert np.allclose(both(lib, make_3), q.transpose(0, 2, 1) @ dq, rtol=1e-4, atol=1e-3)

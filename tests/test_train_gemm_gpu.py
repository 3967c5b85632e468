"""The tensor-core GEMM of the training step (catre_b200/csrc/train_gemm_tc.cuh) against an fp64 product of the same operands,
through the C ABI's debug entry, over the stride / batch / split-K / ragged-edge combinations the training chain produces
(train_chain.cuh: forward layers, weight gradients, data gradients, per-set products) -- and against the CUDA-core tile kernel.

Tolerances (relative to the largest |entry| of the exact product): fp16 hi/lo operands 1e-5 (22 significand bits per operand,
fp32 accumulation; measured 5e-6 at K = 1091), bf16 hi/lo operands 5e-5 (16 bits).  The tensor core's fp32 accumulator
truncates when it aligns an addend, so the error of ONE unsplit reduction grows linearly with its length (measured 1.4e-5 at
K = 4096): the bound is scaled by K / 2048 above 2048 terms.  The training chain never issues such a reduction -- it cuts every
K >= 4096 product into slices of at most 1024 terms (train_chain.cuh, Chain::gemm) whose partial sums are added in rounded
fp32 by KSplitReduce, which is what test_split_k covers at the unscaled bound."""
import ctypes

import pytest
import torch

from catre_b200 import engine

pytestmark = pytest.mark.gpu
TOL = {0: 1e-5, 1: 1e-5, 2: 5e-5}


def run_gemm(A, sa, B, sb, C, sc, M, N, K, batch=1, bias=None, relu=0, acc=0, splits=1, kernel=1):
    """A, B, C: flat CUDA tensors; sa = (sam, sak, sab), sb = (sbk, sbn, sbb), sc = (scm, scn, scb)."""
    lib = engine.load_library()
    st = (ctypes.c_int64 * 9)(*sa, *sb, *sc)
    partial = torch.empty(max(splits, 1) * M * N, device="cuda") if splits > 1 else None
    rc = lib.catre_debug_train_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), None if bias is None else bias.data_ptr(), st, M, N, K,
                                    batch, relu, acc, splits, None if partial is None else partial.data_ptr(), kernel, 0, None, None,
                                    torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()


def strided(flat, off, shape, strides):
    return torch.as_strided(flat, shape, strides, off)


CASES = [
    # name,                    M,    N,    K,  a_kfast, b_kfast, batch
    ("layer_128_1024",         4096, 1024, 128, True,  True,  1),   # forward layer: x [rows, K] . W[C, K]^T
    ("layer_ragged",           300,  200,  100, True,  True,  1),
    ("layer_k1091",            64,   256,  1091, True, True,  1),   # ts-head layer 0
    ("dx_256_64",              2048, 64,   256, True,  False, 1),   # dx = dy . W   (B(k, n) = W[k, n]: n fast)
    ("dw_rowfast_both",        512,  128,  4096, False, False, 1),  # dW = dy^T x  (A(m, k) = dy[k, m]: m fast)
    ("per_set_64",             1024, 64,   64,  True,  False, 8),   # pf = h1 . T64 per set
    ("per_set_t",              64,   64,   1024, False, False, 4),  # dT64 = h1^T . dpf per set
    ("small_m_fc",             32,   512,  1024, True,  True,  1),  # T-Net FC1 on 32 sets
    ("tiny",                   16,   16,   16,  True,  True,  1),
    ("one_and_a_bit_tiles",    129,  130,  65,  True,  False, 2),
]


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_gemm_matches_fp64(case, kernel):
    name, M, N, K, a_kfast, b_kfast, batch = case
    g = torch.Generator(device="cuda").manual_seed(sum(name.encode()))
    pad = 4  # leading dimensions a little larger than the extent, and an offset base pointer that stays 16-byte aligned
    if a_kfast:
        lda = K + pad; sa = (lda, 1, M * lda)
    else:
        lda = M + pad; sa = (1, lda, K * lda)
    if b_kfast:
        ldb = K + pad; sb = (1, ldb, N * ldb)
    else:
        ldb = N + pad; sb = (ldb, 1, K * ldb)
    ldc = N + pad; sc = (ldc, 1, M * ldc)
    A = torch.randn(batch * sa[2] + 8, device="cuda", generator=g)
    B = torch.randn(batch * sb[2] + 8, device="cuda", generator=g) * 0.1
    if kernel == 2:  # backward operands: gradients many decades below 1
        A = A * 1e-7
    bias = torch.randn(N, device="cuda", generator=g) * (1e-7 if kernel == 2 else 1.0)
    C0 = torch.randn(batch * sc[2] + 8, device="cuda", generator=g) * (1e-7 if kernel == 2 else 1.0)
    Av = strided(A, 4, (batch, M, K), (sa[2], sa[0], sa[1])).double()
    Bv = strided(B, 4, (batch, K, N), (sb[2], sb[0], sb[1])).double()
    exact = Av @ Bv
    scale = exact.abs().max().item()
    for relu, acc, use_bias in ((0, 0, False), (1, 0, True), (0, 1, True)):
        C = C0.clone()
        run_gemm(A[4:], sa, B[4:], sb, C[4:], sc, M, N, K, batch, bias if use_bias else None, relu, acc, 1, kernel)
        want = exact + (bias.double() if use_bias else 0.0)
        if relu:
            want = want.clamp_min(0.0)
        if acc:
            want = want + strided(C0, 4, (batch, M, N), (sc[2], sc[0], sc[1])).double()
        got = strided(C, 4, (batch, M, N), (sc[2], sc[0], sc[1])).double()
        err = (got - want).abs().max().item() / scale
        assert err <= TOL[kernel] * max(1.0, K / 2048), (name, kernel, relu, acc, err)
        # nothing outside the [M, N] window of each batch was touched
        mask = torch.ones_like(C, dtype=torch.bool)
        strided(mask, 4, (batch, M, N), (sc[2], sc[0], sc[1])).fill_(False)
        assert torch.equal(C[mask], C0[mask]), name


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("splits", [2, 7, 32])
def test_split_k(kernel, splits):
    """Weight-gradient shape: a long reduction cut into `splits` slices (the last ones may be short or empty)."""
    M, N, K = 256, 192, 4000
    g = torch.Generator(device="cuda").manual_seed(5 + splits)
    A = torch.randn(K, M, device="cuda", generator=g)   # A(m, k) = dy[k, m]
    B = torch.randn(K, N, device="cuda", generator=g)   # B(k, n) = x[k, n]
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    run_gemm(A, (1, M, 0), B, (N, 1, 0), C, (N, 1, 0), M, N, K, 1, None, 0, 1, splits, kernel)
    want = A.double().t() @ B.double() + C0.double()
    err = (C.double() - want).abs().max().item() / want.abs().max().item()
    assert err <= TOL[kernel], (kernel, splits, err)
    C2 = C0.clone()
    run_gemm(A, (1, M, 0), B, (N, 1, 0), C2, (N, 1, 0), M, N, K, 1, None, 0, 1, splits, kernel)
    assert torch.equal(C, C2)  # fixed summation order


def test_unaligned_operands_take_the_scalar_path():
    """Base pointers and leading dimensions that are not multiples of 4 floats (e.g. the K = 1091 ts-head input)."""
    M, N, K = 200, 136, 1091
    g = torch.Generator(device="cuda").manual_seed(77)
    A = torch.randn(M * K + 3, device="cuda", generator=g)
    B = torch.randn(N * K + 1, device="cuda", generator=g)
    C = torch.zeros(M * N + 2, device="cuda")
    run_gemm(A[3:], (K, 1, 0), B[1:], (1, K, 0), C[2:], (N, 1, 0), M, N, K, kernel=1)
    want = A[3:].view(M, K).double() @ B[1:].view(N, K).double().t()
    err = (C[2:].view(M, N).double() - want).abs().max().item() / want.abs().max().item()
    assert err <= TOL[1], err


@pytest.mark.parametrize("kernel,relu", [(1, 1), (1, 0), (2, 0)])
def test_fused_column_max_and_argmax(kernel, relu):
    """The max-pool over each set's points fused into the epilogue: values equal the pooled stored output of the same kernel
    bit for bit, the arg-max follows torch.max's first-index rule (also among the exact ties a ReLU produces at 0)."""
    sets, rows_per_set, N, K = 6, 256, 1024, 128
    M = sets * rows_per_set
    g = torch.Generator(device="cuda").manual_seed(31 + kernel)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g) - (1.5 if relu else 0.0)  # with the ReLU: many all-zero columns
    C = torch.empty(M, N, device="cuda")
    run_gemm(A, (K, 1, 0), W, (1, K, 0), C, (N, 1, 0), M, N, K, 1, bias, relu, 0, 1, kernel)
    want_v, want_i = C.view(sets, rows_per_set, N).max(dim=1)
    # torch.max on CUDA does not promise the first index among ties: take it explicitly
    rows = torch.arange(rows_per_set, device="cuda").view(1, -1, 1).expand(sets, rows_per_set, N)
    first = torch.where(C.view(sets, rows_per_set, N) == want_v.unsqueeze(1), rows, rows_per_set).min(dim=1).values
    lib = engine.load_library()
    st = (ctypes.c_int64 * 9)(K, 1, 0, 1, K, 0, N, 1, 0)
    vmax = torch.empty(sets, N, device="cuda")
    arg = torch.empty(sets, N, device="cuda", dtype=torch.int32)
    scratch = torch.empty(2 * sets * N, device="cuda")
    rc = lib.catre_debug_train_gemm(A.data_ptr(), W.data_ptr(), None, bias.data_ptr(), st, M, N, K, 1, relu, 0, 1, scratch.data_ptr(),
                                    kernel, rows_per_set, vmax.data_ptr(), arg.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(vmax, want_v)
    assert torch.equal(arg.long(), first)
    if relu:
        assert int((want_v == 0).sum()) > 0  # the tie case was exercised

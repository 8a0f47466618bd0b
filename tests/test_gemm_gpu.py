"""GPU parity tests of the tcgen05 GEMM through the C-ABI (pi05_gemm_bf16) against torch fp32 matmul of the same bf16
inputs.  Tolerance: the output is rounded to bf16 once, so rel-L2 <= 2.5e-3 (measured 1.66e-3) against the unrounded
fp32 product and ~1e-4 against the bf16-rounded expectation of each fused epilogue."""
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _ref(a, b, a_major, b_major):
    af = a.float().transpose(-1, -2) if a_major else a.float()
    bf = b.float().transpose(-1, -2) if b_major else b.float()
    return af @ bf.transpose(-1, -2)


CASES = [
    # M, N, K, a_major, b_major, batch, block_n
    (128, 256, 64, 0, 0, None, 0),
    (512, 1024, 512, 0, 0, None, 0),
    (200, 328, 136, 0, 0, None, 0),      # ragged in every dimension
    (4736, 2304, 192, 0, 0, None, 0),    # many tiles, persistent loop + tile swizzle
    (100, 72, 72, 0, 0, None, 128),
    (50, 1024, 1024, 0, 0, None, 0),     # decode-shaped (M = action horizon)
    (1, 256, 128, 0, 0, None, 0),
    (256, 8, 64, 0, 0, None, 0),
    (250, 200, 72, 0, 0, 5, 0),          # batched
    (200, 328, 136, 0, 1, None, 0),      # dgrad form (N-major B)
    (256, 384, 192, 0, 1, None, 128),
    (200, 328, 136, 1, 0, None, 0),
    (512, 768, 1000, 1, 1, None, 0),     # wgrad form (both MN-major)
    (256, 256, 200, 1, 1, 3, 0),
]


@pytest.mark.parametrize("M,N,K,a_major,b_major,batch,block_n", CASES)
def test_gemm_layouts(M, N, K, a_major, b_major, batch, block_n):
    from kai0_b200 import gemm as G

    sa = (M, K) if a_major == 0 else (K, M)
    sb = (N, K) if b_major == 0 else (K, N)
    if batch:
        sa, sb = (batch,) + sa, (batch,) + sb
    a, b = _mk(sa, 1), _mk(sb, 2)
    d = G.gemm(a, b, a_major=a_major, b_major=b_major, block_n=block_n)
    torch.cuda.synchronize()
    assert H.rel_err(d, _ref(a, b, a_major, b_major)) < 2.5e-3


def test_gemm_two_level_batch_strided_heads():
    """The SigLIP attention view: per (image, head) slices of a fused [tokens, 3W] qkv buffer, head_dim 72."""
    from kai0_b200 import _lib
    import ctypes as C

    nimg, VH, T, vhd = 3, 4, 64, 72
    W = VH * vhd
    qkv = _mk((nimg * T, 3 * W), 3)
    out = torch.zeros(nimg * VH, T, T, device="cuda", dtype=torch.bfloat16)
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch, d.batch_inner = T, T, vhd, nimg * VH, VH
    d.A, d.B = qkv.data_ptr(), qkv.data_ptr() + W * 2
    d.lda = d.ldb = 3 * W
    d.a_batch_stride = d.b_batch_stride = vhd
    d.a_batch_stride1 = d.b_batch_stride1 = T * 3 * W
    d.epilogue = _lib.EPI_SCALE
    d.scale = vhd ** -0.5
    d.D, d.ldd, d.d_batch_stride, d.d_batch_stride1 = out.data_ptr(), T, T * T, VH * T * T
    d.block_n = 128
    _lib.check(_lib.lib().pi05_gemm_bf16(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "gemm")
    torch.cuda.synchronize()
    q = qkv[:, :W].view(nimg, T, VH, vhd).permute(0, 2, 1, 3).float()
    k = qkv[:, W:2 * W].view(nimg, T, VH, vhd).permute(0, 2, 1, 3).float()
    ref = ((q @ k.transpose(-1, -2)).to(torch.bfloat16).float() * (vhd ** -0.5)).to(torch.bfloat16)
    assert H.rel_err(out.view(nimg, VH, T, T), ref) < 2e-4


def test_gemm_epilogues():
    from kai0_b200 import gemm as G

    M, N, K = 300, 520, 264
    a, w = _mk((M, K), 1), _mk((N, K), 2, 0.2)
    acc = a.float() @ w.float().t()
    bias, res, gate = _mk((N,), 3), _mk((M, N), 4), _mk((6, N), 5)
    bf = lambda x: x.to(torch.bfloat16).float()  # noqa: E731
    gelu = lambda x: torch.nn.functional.gelu(x, approximate="tanh")  # noqa: E731
    tol = 2e-4
    assert H.rel_err(G.gemm(a, w, epilogue=G.EPI_SCALE, scale=0.1178511), bf(bf(acc) * 0.1178511)) < tol
    assert H.rel_err(G.gemm(a, w, epilogue=G.EPI_BIAS, bias=bias), bf(acc + bias.float())) < tol
    pre, act = G.gemm(a, w, epilogue=G.EPI_BIAS_GELU, bias=bias)
    assert H.rel_err(pre, bf(acc + bias.float())) < tol and H.rel_err(act, bf(gelu(bf(acc + bias.float())))) < tol
    assert H.rel_err(G.gemm(a, w, epilogue=G.EPI_RES, res=res), bf(res.float() + bf(acc))) < tol
    assert H.rel_err(G.gemm(a, w, epilogue=G.EPI_RES, res=res, bias=bias), bf(res.float() + bf(acc + bias.float()))) < tol
    gexp = gate.float().repeat_interleave(50, dim=0)[:M]
    assert H.rel_err(G.gemm(a, w, epilogue=G.EPI_RES, res=res, gate=gate, gate_rows=50),
                     bf(res.float() + bf(bf(acc) * gexp))) < tol
    d = G.gemm(a, w, epilogue=G.EPI_F32)
    assert H.rel_err(d, acc) < 1e-5
    G.gemm(a, w, epilogue=G.EPI_F32, out=d, accumulate=True)
    assert H.rel_err(d, 2 * acc) < 1e-5
    Nh = 264
    wgu = _mk((2 * Nh, K), 7, 0.2)
    gu, h = G.gemm(a, wgu, epilogue=G.EPI_GEGLU, n_out=Nh)
    g, u = bf(a.float() @ wgu[:Nh].float().t()), bf(a.float() @ wgu[Nh:].float().t())
    assert H.rel_err(gu[:, :Nh], g) < tol and H.rel_err(gu[:, Nh:], u) < tol
    assert H.rel_err(h, bf(bf(gelu(g)) * u)) < tol


def test_gemm_fused_backward_epilogues():
    """EPI_GEGLU_BWD / EPI_GELU_BWD (used for the small expert stream): dgrad with the activation backward fused."""
    import ctypes as C

    from kai0_b200 import _lib

    M, Nh, K = 300, 264, 136
    dy, wd = _mk((M, K), 1), _mk((K, Nh), 2, 0.2)       # dH = dY @ Wd, Wd stored [K, Nh] (N-major B)
    gu = _mk((M, 2 * Nh), 3)
    bf = lambda x: x.to(torch.bfloat16).float()  # noqa: E731
    dh = bf(dy.float() @ wd.float())
    g, u = gu[:, :Nh].float().requires_grad_(True), gu[:, Nh:].float()
    a = torch.nn.functional.gelu(g, approximate="tanh")
    (dgelu,) = torch.autograd.grad(a.sum(), g)
    du_ref = bf(dh * bf(a.detach()))
    dg_ref = bf(bf(dh * u) * dgelu)

    def run(epi, res, ldres, out, ldd, N):
        d = _lib.GemmDesc()
        d.M, d.N, d.K, d.batch = M, N, K, 1
        d.A, d.B, d.lda, d.ldb, d.b_major = dy.data_ptr(), wd.data_ptr(), K, Nh, 1
        d.epilogue, d.D, d.ldd, d.res, d.ldres = epi, out.data_ptr(), ldd, res.data_ptr(), ldres
        _lib.check(_lib.lib().pi05_gemm_bf16(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "gemm")
        torch.cuda.synchronize()

    dgu = torch.zeros(M, 2 * Nh, device="cuda", dtype=torch.bfloat16)
    run(7, gu, 2 * Nh, dgu, 2 * Nh, Nh)
    assert H.rel_err(dgu[:, :Nh], dg_ref) < 2e-3 and H.rel_err(dgu[:, Nh:], du_ref) < 2e-3
    pre = _mk((M, Nh), 4)
    x = pre.float().requires_grad_(True)
    (dpre_gelu,) = torch.autograd.grad(torch.nn.functional.gelu(x, approximate="tanh").sum(), x)
    dpre = torch.zeros(M, Nh, device="cuda", dtype=torch.bfloat16)
    run(8, pre, Nh, dpre, Nh, Nh)
    assert H.rel_err(dpre, bf(dh * dpre_gelu)) < 2e-3


def test_gemm_linearity_and_determinism_at_full_size():
    """Size-independent properties at the real MLP shape: D(a1 + a2) ~= D(a1) + D(a2) in fp32 accumulation, and two
    launches of the same problem are bit-identical."""
    from kai0_b200 import gemm as G

    M, N, K = 4096, 2048, 16384
    a1, a2, w = _mk((M, K), 1, 0.5), _mk((M, K), 2, 0.5), _mk((N, K), 3, 0.02)
    s = (a1.float() + a2.float()).to(torch.bfloat16)
    d1 = G.gemm(a1, w, epilogue=G.EPI_F32)
    d2 = G.gemm(a2, w, epilogue=G.EPI_F32)
    ds = G.gemm(s, w, epilogue=G.EPI_F32)
    # (a1+a2) is itself rounded to bf16, so allow that rounding (2^-9 relative per element, averaged out)
    assert H.rel_err(ds, d1 + d2) < 3e-3
    again = G.gemm(a1, w, epilogue=G.EPI_F32)
    assert torch.equal(d1, again)


def test_gemm_rejects_bad_arguments():
    from kai0_b200 import gemm as G

    a, b = _mk((64, 60), 1), _mk((64, 60), 2)  # K = 60 -> row pitch not 16-byte aligned
    with pytest.raises(RuntimeError, match="16B aligned"):
        G.gemm(a, b)


@pytest.mark.parametrize("M,N,K,a_major,b_major", [
    (1152, 4304, 24576, 1, 1),   # SigLIP fc1 weight gradient: 85 pair-tiles on 74 clusters -> K split
    (2560, 2048, 15488, 1, 1),   # Gemma qkv weight gradient shape (half the bench's token count)
    (1152, 1152, 12288, 1, 1),   # under-filled 128-wide grid
    (50, 2048, 16384, 0, 0),     # decode down-projection: small-M split-K
])
def test_gemm_split_k_paths_match_unsplit(M, N, K, a_major, b_major):
    """With a workspace the library may split K (fp32 partials summed in a fixed order).  The result must agree with
    the un-split kernel to fp32-summation-order noise and be bit-identical run to run."""
    from kai0_b200 import gemm as G

    sa = (M, K) if a_major == 0 else (K, M)
    sb = (N, K) if b_major == 0 else (K, N)
    a, b = _mk(sa, 11, 0.5), _mk(sb, 12, 0.05)
    ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    plain = G.gemm(a, b, a_major=a_major, b_major=b_major)
    split = G.gemm(a, b, a_major=a_major, b_major=b_major, workspace=ws)
    again = G.gemm(a, b, a_major=a_major, b_major=b_major, workspace=ws)
    torch.cuda.synchronize()
    assert torch.equal(split, again)
    assert H.rel_err(split, plain) < 1e-3  # both are bf16 roundings of the same fp32 sum up to summation order
    assert H.rel_err(split, _ref(a, b, a_major, b_major)) < 2.5e-3

"""Python handle on the stand-alone tcgen05 GEMM operator (pi05_gemm_bf16). Used by tests and tools."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_F32, EPI_GEGLU, EPI_RES, EPI_SCALE, EPI_STORE  # noqa: F401


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    a_major: int = 0,
    b_major: int = 0,
    epilogue: int = EPI_STORE,
    out: torch.Tensor | None = None,
    out2: torch.Tensor | None = None,
    bias: torch.Tensor | None = None,
    res: torch.Tensor | None = None,
    gate: torch.Tensor | None = None,
    gate_rows: int = 1,
    scale: float = 1.0,
    accumulate: bool = False,
    block_n: int = 0,
    n_out: int | None = None,
    workspace: torch.Tensor | None = None,
):
    """D[z] = A[z] @ B[z]^T with the fused epilogues of csrc/gemm.h.

    a: [M,K] / [Z,M,K] (a_major=0) or [K,M] / [Z,K,M] (a_major=1); b likewise with N.  2-D b with 3-D a is
    shared across the batch.  For EPI_GEGLU b is the fused [2N,K] gate|up weight and n_out=N.
    """
    assert a.is_cuda and b.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a3 = a if a.dim() == 3 else a[None]
    b3 = b if b.dim() == 3 else b[None]
    assert a3.stride(-1) == 1 and b3.stride(-1) == 1
    Z = max(a3.shape[0], b3.shape[0])
    if a_major == 0:
        M, K = a3.shape[1], a3.shape[2]
    else:
        K, M = a3.shape[1], a3.shape[2]
    if b_major == 0:
        Nb, Kb = b3.shape[1], b3.shape[2]
    else:
        Kb, Nb = b3.shape[1], b3.shape[2]
    assert K == Kb, (K, Kb)
    N = n_out if n_out is not None else Nb
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, Z
    d.A, d.B = _ptr(a3), _ptr(b3)
    d.a_major, d.b_major = a_major, b_major
    d.lda, d.ldb = a3.stride(1), b3.stride(1)
    d.a_batch_stride = a3.stride(0) if a3.shape[0] > 1 else 0
    d.b_batch_stride = b3.stride(0) if b3.shape[0] > 1 else 0
    d.epilogue = epilogue
    if out is None:
        cols = 2 * N if epilogue == EPI_GEGLU else N
        out = torch.empty(
            (Z, M, cols), device=a.device, dtype=torch.float32 if epilogue == EPI_F32 else torch.bfloat16
        )
    o3 = out if out.dim() == 3 else out[None]
    d.D, d.ldd, d.d_batch_stride = _ptr(o3), o3.stride(1), (o3.stride(0) if o3.shape[0] > 1 else 0)
    if epilogue in (EPI_GEGLU, EPI_BIAS_GELU) and out2 is None:
        out2 = torch.empty((Z, M, N), device=a.device, dtype=torch.bfloat16)
    if out2 is not None:
        o23 = out2 if out2.dim() == 3 else out2[None]
        d.D2, d.ldd2, d.d2_batch_stride = _ptr(o23), o23.stride(1), (o23.stride(0) if o23.shape[0] > 1 else 0)
    d.bias = _ptr(bias)
    if res is not None:
        r3 = res if res.dim() == 3 else res[None]
        d.res, d.ldres, d.res_batch_stride = _ptr(r3), r3.stride(1), (r3.stride(0) if r3.shape[0] > 1 else 0)
    if gate is not None:
        d.gate, d.gate_rows, d.ldgate = _ptr(gate), gate_rows, gate.stride(0)
    d.scale = scale
    d.accumulate = 1 if accumulate else 0
    d.block_n = block_n
    if workspace is not None:  # fp32 scratch: lets the library pick a split-K schedule (see include/pi05.h)
        d.workspace, d.workspace_bytes = _ptr(workspace), workspace.numel() * workspace.element_size()
    stream = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(_lib.lib().pi05_gemm_bf16(C.byref(d), stream), "pi05_gemm_bf16")
    if a.dim() == 2 and b.dim() == 2:
        out = out[0] if out.dim() == 3 else out
        if out2 is not None and out2.dim() == 3:
            out2 = out2[0]
    if epilogue in (EPI_GEGLU, EPI_BIAS_GELU):
        return out, out2
    return out

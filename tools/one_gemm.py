"""Launch one GEMM class a few times (for `ncu --set full -k regex:gemm -c N python tools/one_gemm.py <case>`).

cases: geglu (M=30976, N=2x16384, K=2048, fused GeGLU epilogue), dgrad (M=30976, N=2048, K=32768, N-major B),
       wgrad (M=32768, N=2048, K=30976, both MN-major), down (M=30976, N=2048, K=16384, residual epilogue)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kai0_b200 import gemm as G  # noqa: E402


def mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "geglu"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    M = 30976
    if case == "geglu":
        a, w = mk((M, 2048), 1), mk((32768, 2048), 2, 0.02)
        f = lambda: G.gemm(a, w, epilogue=G.EPI_GEGLU, n_out=16384, out=out, out2=out2)  # noqa: E731
        out = torch.empty(M, 32768, device="cuda", dtype=torch.bfloat16)
        out2 = torch.empty(M, 16384, device="cuda", dtype=torch.bfloat16)
        flop = 2.0 * M * 32768 * 2048
    elif case == "dgrad":
        a, w = mk((M, 32768), 1, 0.1), mk((32768, 2048), 2, 0.02)
        out = torch.empty(M, 2048, device="cuda", dtype=torch.bfloat16)
        f = lambda: G.gemm(a, w, b_major=1, out=out)  # noqa: E731
        flop = 2.0 * M * 32768 * 2048
    elif case == "wgrad":
        a, x = mk((M, 32768), 1, 0.1), mk((M, 2048), 2)
        out = torch.empty(32768, 2048, device="cuda", dtype=torch.bfloat16)
        f = lambda: G.gemm(a, x, a_major=1, b_major=1, out=out)  # noqa: E731
        flop = 2.0 * M * 32768 * 2048
    else:
        a, w, res = mk((M, 16384), 1, 0.1), mk((2048, 16384), 2, 0.02), mk((M, 2048), 3)
        out = torch.empty(M, 2048, device="cuda", dtype=torch.bfloat16)
        f = lambda: G.gemm(a, w, epilogue=G.EPI_RES, res=res, out=out)  # noqa: E731
        flop = 2.0 * M * 16384 * 2048
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{case}: {ms:.3f} ms {flop / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()

"""Bring-up probe for the tcgen05 GEMM: runs each case group in its own subprocess (a device trap poisons the
CUDA context) and prints max-abs / relative errors against torch fp32 matmul of the same bf16 inputs, plus timings.

    python tools/gemm_probe.py            # all groups
    python tools/gemm_probe.py kk         # one group (child mode)
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["kk", "kk_bn128", "small", "batch", "b_mn", "a_mn", "ab_mn", "epi", "perf"]


def rel(a, b):
    import torch

    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def mk(shape, seed, scale=1.0):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(torch.bfloat16)


def run_group(name):
    import torch

    from kai0_b200 import gemm as G

    out = []

    def ref_mm(a, b, a_major, b_major):
        af = a.float().transpose(-1, -2) if a_major else a.float()
        bf = b.float().transpose(-1, -2) if b_major else b.float()
        return af @ bf.transpose(-1, -2)

    def case(tag, M, N, K, a_major=0, b_major=0, batch=None, block_n=0, pad=0):
        shp_a = (M, K) if a_major == 0 else (K, M)
        shp_b = (N, K) if b_major == 0 else (K, N)
        if batch:
            shp_a = (batch,) + shp_a
            shp_b = (batch,) + shp_b
        a = mk(shp_a, 1)
        b = mk(shp_b, 2)
        d = G.gemm(a, b, a_major=a_major, b_major=b_major, block_n=block_n)
        torch.cuda.synchronize()
        r = ref_mm(a, b, a_major, b_major)
        e = rel(d, r)
        out.append({"case": tag, "M": M, "N": N, "K": K, "rel": e[0], "maxabs": e[1]})
        print(f"[{name}] {tag}: M={M} N={N} K={K} a_major={a_major} b_major={b_major} batch={batch} bn={block_n} "
              f"rel={e[0]:.3e} maxabs={e[1]:.3e}", flush=True)

    if name == "kk":
        case("1tile_1kb", 128, 256, 64)
        case("1tile_4kb", 128, 256, 256)
        case("multi", 512, 1024, 512)
        case("ragged", 200, 328, 136)
        case("many_tiles", 128 * 37, 256 * 9, 192)
    elif name == "kk_bn128":
        case("bn128_1tile", 128, 128, 64, block_n=128)
        case("bn128_multi", 384, 640, 320, block_n=128)
        case("bn128_ragged", 100, 72, 72, block_n=128)
    elif name == "small":
        case("m50", 50, 1024, 1024)
        case("m1", 1, 256, 128)
        case("n8", 256, 8, 64)
    elif name == "batch":
        case("batch3", 256, 256, 128, batch=3)
        case("batch5_ragged", 250, 200, 72, batch=5)
    elif name == "b_mn":
        case("b_mn_1tile", 128, 256, 64, b_major=1)
        case("b_mn_multi", 256, 512, 256, b_major=1)
        case("b_mn_bn128", 256, 384, 192, b_major=1, block_n=128)
        case("b_mn_ragged", 200, 328, 136, b_major=1)
    elif name == "a_mn":
        case("a_mn_1tile", 128, 256, 64, a_major=1)
        case("a_mn_multi", 256, 512, 256, a_major=1)
        case("a_mn_ragged", 200, 328, 136, a_major=1)
    elif name == "ab_mn":
        case("ab_mn_1tile", 128, 256, 64, a_major=1, b_major=1)
        case("ab_mn_multi", 512, 768, 1000, a_major=1, b_major=1)
        case("ab_mn_batch", 256, 256, 200, a_major=1, b_major=1, batch=3)
    elif name == "epi":
        M, N, K = 300, 520, 264
        a = mk((M, K), 1)
        w = mk((N, K), 2, 0.2)
        acc = a.float() @ w.float().t()
        bias = mk((N,), 3)
        res = mk((M, N), 4)
        gate = mk((6, N), 5)
        bf = lambda x: x.to(torch.bfloat16).float()  # noqa: E731
        # SCALE
        d = G.gemm(a, w, epilogue=G.EPI_SCALE, scale=0.117851130)
        print("[epi] scale", rel(d, bf(bf(acc) * 0.117851130)), flush=True)
        d = G.gemm(a, w, epilogue=G.EPI_BIAS, bias=bias)
        print("[epi] bias", rel(d, bf(acc + bias.float())), flush=True)
        d, d2 = G.gemm(a, w, epilogue=G.EPI_BIAS_GELU, bias=bias)
        pre = bf(acc + bias.float())
        print("[epi] bias_gelu pre", rel(d, pre), "act", rel(d2, bf(torch.nn.functional.gelu(pre, approximate="tanh"))),
              flush=True)
        d = G.gemm(a, w, epilogue=G.EPI_RES, res=res)
        print("[epi] res", rel(d, bf(res.float() + bf(acc))), flush=True)
        d = G.gemm(a, w, epilogue=G.EPI_RES, res=res, bias=bias)
        print("[epi] res+bias", rel(d, bf(res.float() + bf(acc + bias.float()))), flush=True)
        d = G.gemm(a, w, epilogue=G.EPI_RES, res=res, gate=gate, gate_rows=50)
        gexp = gate.float().repeat_interleave(50, dim=0)[:M]
        print("[epi] gated res", rel(d, bf(res.float() + bf(bf(acc) * gexp))), flush=True)
        d = G.gemm(a, w, epilogue=G.EPI_F32)
        print("[epi] f32", rel(d, acc), flush=True)
        d2 = d.clone()
        G.gemm(a, w, epilogue=G.EPI_F32, out=d2, accumulate=True)
        print("[epi] f32 acc", rel(d2, 2 * acc), flush=True)
        # GEGLU: fused [2N,K] weight
        Nh = 264
        wgu = mk((2 * Nh, K), 7, 0.2)
        gu, h = G.gemm(a, wgu, epilogue=G.EPI_GEGLU, n_out=Nh)
        g = bf(a.float() @ wgu[:Nh].float().t())
        u = bf(a.float() @ wgu[Nh:].float().t())
        hh = bf(bf(torch.nn.functional.gelu(g, approximate="tanh")) * u)
        print("[epi] geglu g", rel(gu[:, :Nh], g), "u", rel(gu[:, Nh:], u), "h", rel(h, hh), flush=True)
    elif name == "perf":
        def bench(tag, M, N, K, a_major=0, b_major=0, epilogue=0, batch=None, n_out=None, iters=10):
            shp_a = (M, K) if a_major == 0 else (K, M)
            rows_b = N if n_out is None else 2 * N
            shp_b = (rows_b, K) if b_major == 0 else (K, rows_b)
            if batch:
                shp_a = (batch,) + shp_a
                shp_b = (batch,) + shp_b
            a = mk(shp_a, 1)
            b = mk(shp_b, 2, 0.05)
            kw = dict(a_major=a_major, b_major=b_major, epilogue=epilogue)
            if n_out:
                kw["n_out"] = n_out
            outs = G.gemm(a, b, **kw)
            o1 = outs[0] if isinstance(outs, tuple) else outs
            o2 = outs[1] if isinstance(outs, tuple) else None
            for _ in range(3):
                G.gemm(a, b, out=o1, out2=o2, **kw)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                G.gemm(a, b, out=o1, out2=o2, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = 2.0 * M * (rows_b) * K * (batch or 1)
            # cuBLAS reference timing for the same math
            af = a.transpose(-1, -2) if a_major else a
            bfm = b if b_major else b.transpose(-1, -2)
            for _ in range(3):
                torch.matmul(af, bfm)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                torch.matmul(af, bfm)
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / iters
            print(f"[perf] {tag}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s   (cuBLAS {ms_ref:.3f} ms "
                  f"{fl / ms_ref / 1e9:.1f} TFLOP/s)", flush=True)
            out.append({"case": tag, "ms": ms, "tflops": fl / ms / 1e9, "cublas_tflops": fl / ms_ref / 1e9})

        bench("sq4096", 4096, 4096, 4096)
        bench("sq8192", 8192, 8192, 8192)
        bench("mlp_down M30976 N2048 K16384", 30976, 2048, 16384)
        bench("mlp_geglu M30976 N16384x2 K2048", 30976, 16384, 2048, epilogue=G.EPI_GEGLU, n_out=16384)
        bench("qkv M30976 N2560 K2048", 30976, 2560, 2048)
        bench("vit_fc1 M24576 N4304 K1152", 24576, 4304, 1152)
        bench("dgrad b_mn M30976 N2048 K16384", 30976, 2048, 16384, b_major=1)
        bench("wgrad ab_mn M2048 N16384 K30976", 2048, 16384, 30976, a_major=1, b_major=1)
        bench("attn_scores batch32 M7744 N968 K256", 7744, 968, 256, batch=32)
        bench("decode_m50 N8192 K1024", 50, 8192, 1024)
    print("GROUP_JSON " + json.dumps({"group": name, "results": out}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        run_group(sys.argv[1])
        return
    os.makedirs("gpurun_out", exist_ok=True)
    log = open("gpurun_out/gemm_probe.log", "w")
    for g in GROUPS:
        envs = [dict()]
        if g in ("b_mn", "a_mn", "ab_mn"):
            envs.append({"PI05_DBG_MN_LBO": "1024", "PI05_DBG_MN_SBO": "8192"})
        for extra in envs:
            env = dict(os.environ)
            env.update(extra)
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, g], env=env, capture_output=True, text=True, timeout=240)
                txt = p.stdout + ("\nSTDERR:\n" + p.stderr[-3000:] if p.returncode != 0 else "")
                rc = p.returncode
            except subprocess.TimeoutExpired as e:
                txt = f"TIMEOUT\n{(e.stdout or b'')[-2000:]}"
                rc = -9
            hdr = f"===== group {g} env={extra} rc={rc} ({time.time() - t0:.1f}s) ====="
            print(hdr)
            print(txt, flush=True)
            log.write(hdr + "\n" + txt + "\n")
            log.flush()


if __name__ == "__main__":
    main()

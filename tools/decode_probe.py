"""Decode-latency probe (B=1, full architecture): p50 of sample_actions through the public API, plus a
torch.profiler (CUPTI) kernel breakdown of one call: kernel-time sum vs wall time tells launch-bound from GPU-bound."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as Hh  # noqa: E402
from oracle import pi05_oracle as O  # noqa: E402
from kai0_b200.pi0_pytorch import PI0Pytorch, Pi05EngineConfig  # noqa: E402


def main():
    model = PI0Pytorch(Pi05EngineConfig(), init_weights=False).to("cuda")
    model.reset_parameters(seed=7)
    model.check_inputs = False
    model.eval()
    oc = O.OracleConfig()
    batch = O.synthetic_batch(oc, 1)
    obs = Hh.Obs(batch, "cuda")
    noise = batch["noise"].cuda()
    for _ in range(3):
        model.sample_actions("cuda", obs, noise=noise)
    torch.cuda.synchronize()
    lat = []
    for _ in range(20):
        t0 = time.perf_counter()
        a = model.sample_actions("cuda", obs, noise=noise)
        a.cpu()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    print(f"sample_actions B=1 10 steps: p50 {lat[10]:.2f} ms  min {lat[0]:.2f}  max {lat[-1]:.2f}")
    # host-side enqueue time only (no sync)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a = model.sample_actions("cuda", obs, noise=noise)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host enqueue time {1e3 * (t1 - t0):.2f} ms")
    # prefill vs denoise split (eager launches with programmatic dependent launch on, CUDA events on the stream)
    import ctypes as C

    from kai0_b200 import _lib as L

    images, img_masks, toks, tmask, _ = model._preprocess_observation(obs, train=False, rows=True, engine_train=False)
    b, keep = model._make_batch(images, img_masks, toks, tmask)
    out = torch.empty_like(noise)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):
        ev[0].record()
        L.check(L.lib().pi05_prefill(model._engine, C.byref(b), st), "prefill")
        ev[1].record()
        L.check(L.lib().pi05_denoise(model._engine, C.c_void_p(noise.data_ptr()), 10, C.c_void_p(out.data_ptr()), st), "denoise")
        ev[2].record()
        torch.cuda.synchronize()
    print(f"eager (no graph): prefill {ev[0].elapsed_time(ev[1]):.2f} ms, 10-step denoise {ev[1].elapsed_time(ev[2]):.2f} ms")
    from torch.profiler import ProfilerActivity, profile

    from kai0_b200 import _lib

    _lib.lib().pi05_debug_set_pdl(0)  # per-kernel times are only meaningful without PDL staging
    model._graphs = {}  # re-capture without programmatic edges
    model.sample_actions("cuda", obs, noise=noise)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.sample_actions("cuda", obs, noise=noise)
        torch.cuda.synchronize()
    agg = {}
    for ev in prof.events():
        if str(ev.device_type).endswith("CUDA"):
            a_ = agg.setdefault(ev.name[:80], [0, 0.0])
            a_[0] += 1
            a_[1] += ev.device_time_total
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"kernel time sum {tot / 1e3:.2f} ms over {n} launches")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"  {us / 1e3:8.3f} ms  x{cnt:5d}  avg {us / cnt:7.1f} us  {name}")


if __name__ == "__main__":
    main()

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

"""One training step with ONE joint layer, ONE ViT layer (forward + backward) and the optimiser bracketed by
cudaProfilerStart/Stop, for `ncu --set full --profile-from-start off` captures of every kernel class of the path at
the benchmark's architecture (pi05_debug_profile_layer, include/pi05.h).

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_layer \
        python tools/ncu_layer.py [batch=8] [layer=9]

ncu saves/restores device memory around every replayed kernel, so the batch is kept below the bench's 32 (the
kernels stay far larger than L2; the GEMM shapes at batch 32 are captured by tools/one_gemm.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from kai0_b200 import _lib  # noqa: E402
from kai0_b200.model import Observation  # noqa: E402
from kai0_b200.optim import FusedClipAdamW  # noqa: E402
from kai0_b200.pi0_pytorch import PI0Pytorch, Pi05EngineConfig  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    layer = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    cfg = Pi05EngineConfig()
    model = PI0Pytorch(cfg, max_batch=B, init_weights=False).to(dev)
    model.reset_parameters()
    model.check_inputs = False
    model.direct_grads = True
    model.train()
    model.flat_parameters()
    opt = FusedClipAdamW(model, lr=2.5e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10, max_norm=1.0)
    host_d, host_a = bench.make_host_batch(B, 0, pin=False)
    d, a = bench.to_device(host_d, host_a, dev)

    def step(profile_opt=False):
        obs = Observation.from_dict({"image": dict(d["image"]), "image_mask": d["image_mask"], "state": d["state"],
                                     "tokenized_prompt": d["tokenized_prompt"],
                                     "tokenized_prompt_mask": d["tokenized_prompt_mask"]})
        model(obs, a).mean().backward()
        if profile_opt:
            torch.cuda.profiler.start()
        opt.step()
        if profile_opt:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        opt.zero_grad()

    step()
    torch.cuda.synchronize()
    _lib.lib().pi05_debug_profile_layer(model._engine, layer)
    step(profile_opt=True)
    torch.cuda.synchronize()
    _lib.lib().pi05_debug_profile_layer(model._engine, -1)
    print("profiled layer", layer, "batch", B)


if __name__ == "__main__":
    main()

"""Full-size AdvantageEstimator train step on one B200 (reference config ADVANTAGE_TORCH_KAI0_FLATTEN_FOLD,
training/config.py:1220-1271: batch 16 per GPU, 6 images = 2 timesteps x 3 cameras, loss_value_weight 1,
loss_action_weight 0) and `sample_values` latency.  Prints ms/step and samples/s (device events, after warm-up)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kai0_b200.optim import FusedClipAdamW  # noqa: E402
from kai0_b200.pi0_pytorch import AdvantageEstimator, Pi05EngineConfig  # noqa: E402


class Obs:
    pass


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = torch.device("cuda:0")
    cfg = Pi05EngineConfig()
    cfg.loss_value_weight, cfg.loss_action_weight = 1.0, 0.0
    torch.manual_seed(0)
    model = AdvantageEstimator(cfg, max_batch=B, init_weights=False).to(dev)
    model.reset_parameters()
    model.check_inputs = False
    model.direct_grads = True
    model.train()
    model.flat_parameters()
    opt = FusedClipAdamW(model, lr=2.5e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10, max_norm=1.0)
    g = torch.Generator(device=dev).manual_seed(1)
    keys = [f"{p}_{t}_rgb" for t in (-100, 0) for p in ("base", "left_wrist", "right_wrist")]
    obs = Obs()
    obs.images = {k: torch.rand(B, 3, 224, 224, device=dev, generator=g) * 2 - 1 for k in keys}
    obs.image_masks = {k: torch.ones(B, dtype=torch.bool, device=dev) for k in keys}
    obs.state = torch.zeros(B, 32, device=dev)
    obs.tokenized_prompt = torch.randint(1, 257152, (B, 200), device=dev, generator=g)
    obs.tokenized_prompt_mask = torch.arange(200, device=dev)[None, :].expand(B, 200) < 96
    obs.progress = torch.rand(B, device=dev, generator=g) * 2 - 1
    actions = torch.randn(B, 50, 32, device=dev, generator=g)

    def step():
        loss, aux = model(obs, actions, return_loss_dict=True)
        loss.mean().backward()
        opt.step()
        opt.zero_grad()
        return loss, aux

    for _ in range(3):
        loss, aux = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, aux = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"advantage train step B={B} 6 images (P=1736): {ms:.1f} ms/step  {B / ms * 1e3:.1f} samples/s  "
          f"loss {float(loss.mean()):.4f} loss_value {float(aux['loss_value']):.4f}  "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    model.eval()
    lat = []
    for i in range(8):
        t0 = time.perf_counter()
        v = model.sample_values(dev, obs)
        v.cpu()
        if i >= 3:
            lat.append((time.perf_counter() - t0) * 1e3)
    print(f"sample_values B={B}: p50 {sorted(lat)[len(lat) // 2]:.1f} ms, finite {bool(torch.isfinite(v).all())}")


if __name__ == "__main__":
    main()

"""NCCL all-reduce bandwidth probe for the gradient exchange payload (6.47 GB bf16 + 0.48 GB fp32), run under torchrun:
prints algorithmic / bus bandwidth for the environment's NCCL settings (NCCL_ALGO, NCCL_MIN/MAX_CTAS, ...), so tunings can be
compared cheaply before touching the engine."""
import os

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    gb = torch.ones(3_233_000_000, dtype=torch.bfloat16, device="cuda")
    gf = torch.ones(120_000_000, dtype=torch.float32, device="cuda")
    for _ in range(2):
        dist.all_reduce(gb)
        dist.all_reduce(gf)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(gb)
        dist.all_reduce(gf)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ts)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        nbytes = gb.numel() * 2 + gf.numel() * 4
        ms = float(t)
        tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NCCL_") and k != "NCCL_DEBUG")
        print(f"[allreduce N={world}] {tag or 'defaults'}: {ms:.2f} ms, algbw {nbytes / 1e9 / (ms / 1e3):.0f} GB/s, "
              f"busbw {2 * (world - 1) / world * nbytes / 1e9 / (ms / 1e3):.0f} GB/s", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

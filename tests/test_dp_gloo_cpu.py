"""world_size-2 gloo test of the data-parallel exchange step (the only collective on the path):
flat per-dtype gradient arenas are SUM all-reduced and averaged; replicas start from identical weights."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import pi05_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(1234)  # bench.py seeds every rank identically before building the model
        model, _ = H.build_pair(O.tiny_config(), seed=0, device=None)
        model.enable_flat_allreduce()
        for dt, flat in model._flat.items():
            model._flat_grad[dt] = torch.full_like(flat, float(rank + 1))
        model._allreduce_flat_grads()
        # mean of 1 and 2 everywhere except the never-used expert lm_head (last in the bf16 arena), which is not exchanged
        used = model._offsets["paligemma_with_expert.gemma_expert.lm_head.weight"][1]
        gb, gf = model._flat_grad[torch.bfloat16], model._flat_grad[torch.float32]
        ok = bool(torch.all(gb[:used] == 1.5)) and bool(torch.all(gf == 1.5)) and bool(torch.all(gb[used:] == rank + 1))
        # replicas hold identical weights (checksum exchange)
        chk = torch.tensor([float(model._flat[torch.bfloat16].float().sum()), float(model._flat[torch.float32].sum())],
                           dtype=torch.float64)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        same = bool(torch.equal(both[0], both[1]))
        # gradient views handed to autograd alias the reduced arena
        name = "action_in_proj.weight"
        dt, o, n, shape = model._offsets[name]
        view = model._flat_grad[dt][o:o + n].view(shape)
        alias = bool(torch.all(view == 1.5))
        # average="optimizer": the arenas keep the SUM and the optimiser is told to scale by 1 / world
        model.enable_flat_allreduce(average="optimizer")
        for dt, flat in model._flat.items():
            model._flat_grad[dt] = torch.full_like(flat, float(rank + 1))
        model._allreduce_flat_grads()
        ok = ok and bool(torch.all(model._flat_grad[torch.float32] == 3.0)) and model.grad_scale == 0.5
        ok = ok and not model._dp_engine  # gloo / CPU: the torch.distributed path, never the NCCL engine path
        q.put((rank, ok, same, alias))
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, same, alias in res:
        assert ok and same and alias, (rank, ok, same, alias)

"""Two-GPU NCCL tests of the data-parallel path (skipped on a single-GPU box): the engine's overlapped gradient exchange
against the stock DistributedDataParallel wrapper of the unchanged script, N = 2 against N = 1 on the concatenated batch,
the optimiser-folded average, and the one-shot C-ABI entry point.  The checks live in tests/dp/dp_worker.py; this file
launches it under torch.distributed.run exactly as the driver launches bench.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_overlapped_exchange_matches_stock_ddp_and_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and "DP_WORKER_OK" in r.stdout

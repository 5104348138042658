"""Multi-GPU correctness of the data-parallel train step (needs >= 2 GPUs; skipped otherwise).

Two ranks (one process per GPU, torchrun), three steps on shards with different pair counts -- so that the ranks'
CUDA-graph caches miss at different steps -- with the NCCL all-reduce captured in the step graph and with the
peer-memory exchange fused into the optimiser (scann_b200/csrc/p2p.cu): parameters must follow the single-GPU
full-batch run (tools/dp_check.py) and stay bit-identical across ranks."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_train_steps_match_single_gpu_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "DP_CHECK_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]

"""Multi-GPU (NCCL) check of the sharded retrieval plumbing with the real CUDA model: needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`), skipped on a one-GPU box.  The CPU-side
logic of the same code is covered by tests/test_dist_cpu.py over gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_topk_and_rerank_over_nccl_equal_one_gpu(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "gpu_dist_worker.py"),
           str(tmp_path)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        res = torch.load(tmp_path / f"res.{rank}")
        print(rank, res)
        assert res["rows_equal"] and res["scores_equal"] and res["subset_equal"], (rank, res)
        assert res["fetch_equal"], (rank, res)
        assert res["rerank_equal"] and res["rerank_reorders"], (rank, res)

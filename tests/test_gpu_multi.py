"""Multi-GPU parity on real GPUs (NCCL): index-range sharded MSM and column-sharded commit.
Runs tools/multi_gpu_check.py under torchrun; skipped on boxes with a single GPU (the host-side
logic is covered at world_size 2 on CPU by tests/test_dist_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_msm_and_commit_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py"),
           "--log-n", "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["all_ranks_ok"] and res["sharded_msm_ok"] and res["sharded_commit_ok"]

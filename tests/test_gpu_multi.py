"""Multi-GPU expansion over NCCL: identical stores on every rank, equal to the single-GPU
result (needs >= 2 visible GPUs; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_expansion_matches_single_gpu(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "mg.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tools", "multigpu_expand_check.py"), "--seeds", "1500", "--levels", "2",
           "--views", "6", "--width", "320", "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    assert res["ok_all_ranks"] and res["patches"] == res["patches_single"] and res["inserted"] > 0
    assert 0 < res["local_records_rank0"] < res["passed"]

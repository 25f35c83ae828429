"""Multi-GPU path (NCCL, one process per GPU); skipped on a single-GPU box.  The same logic is
covered on CPU by test_host_cpu.py::test_shard_merge_over_gloo_world_size_2."""

import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_chains_equal_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tools", "multi_gpu_check.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

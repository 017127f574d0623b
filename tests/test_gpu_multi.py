"""GPU, more than one device: the sharded E-step with the in-kernel all-reduce (tests/multi_gpu_check.py under torchrun).
Skipped on a single-GPU box; run it with `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload", ["small", "cfg2"])
def test_sharded_estep_matches_single_gpu(workload):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--workload", workload]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])

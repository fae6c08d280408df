"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): partitioned device path vs the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["classic", "feec"])
def test_two_gpus_partitioned_assembly_and_halo_spmv(family):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613" if family == "classic" else "29615", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, DCP_CHECK_FAMILY=family))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "-> OK" in r.stdout


@pytest.mark.gpu
def test_two_ranks_through_the_c_abi():
    """tests/cpp/halo_test.cpp: communicator, ghost exchange, all-reduced dot and max through include/dcp.h only."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(ROOT, "tests", "cpp", "halo_test")
    assert os.path.exists(exe), "tests/cpp/halo_test is not built (make)"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "halo_test: OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

"""World-size-2 checks of the sharded step.  With two GPUs visible the ranks run one per GPU over NCCL; on a one-GPU
box they run as two processes sharing the GPU (CUDA IPC peer memory works between processes on one device; the
all-reduce route then goes through gloo), so the check is never skipped."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(port, env_extra, timeout=900):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "peer_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.gpu
def test_peer_memory_exchange_equals_allreduce_route_and_single_handle():
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    res = _run(29541, {"PEER_CHECK_BACKEND": backend})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "peer_check OK" in res.stdout


@pytest.mark.gpu
def test_grid_curriculum_histograms_are_global_over_the_shards():
    """north_star: "curriculum histograms NCCL-allreduced".  With the pitch x yaw grid curriculum every shard keeps the
    histograms of ALL envs (the step's outcomes are summed over the ranks when the step is closed: by the peer-memory
    exchange kernel or by the caller's all-reduce), so bins, stones and histograms equal those of one handle."""
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    res = _run(29545, {"PEER_CHECK_BACKEND": backend, "PEER_CHECK_GRID": "5", "PEER_CHECK_ENVS": "3000",
                       "PEER_CHECK_STEPS": "10"})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "peer_check OK" in res.stdout and "grid 5x5" in res.stdout


@pytest.mark.gpu
def test_peer_exchange_timeout_is_fatal_and_sticky():
    res = _run(29543, {"PEER_CHECK_BACKEND": "gloo", "PEER_CHECK_MODE": "timeout", "ALLSTEPS_PEER_TIMEOUT_MS": "300"})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "peer_check timeout OK" in res.stdout

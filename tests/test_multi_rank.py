"""N > 1 paths: host logic under gloo on CPU (world_size 2 and 3), NCCL slabs on >= 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp_worker.py")


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def launch(nproc, mode, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), WORKER, "--mode", mode]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 3])
def test_host_logic_gloo(world):
    r = launch(world, "host", 300)
    assert r.returncode == 0 and "HOST-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_nccl_slabs_match_single_rank_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = int(os.environ.get("MACROC_TEST_WORLD", min(n, 4)))
    r = launch(min(world, n), "gpu", 1500)
    assert r.returncode == 0 and "GPU-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

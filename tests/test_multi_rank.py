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


@pytest.mark.gpu
@pytest.mark.parametrize("procs", ["-da_processors_x 1 -da_processors_y 1 -da_processors_z 2", ""])
def test_c_host_driver_two_ranks(tmp_path, procs):
    """macroc_b200/lib/macroc on 2 GPUs: ranks from RANK/WORLD_SIZE/LOCAL_RANK, NCCL id through
    MACROC_ID_FILE (no MPI in the image); z-slabs and the PETSC_DECIDE processor grid.  The log
    must match the reference binary's single-rank run (decomposition independence)."""
    import re
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import load_golden
    import macroc_b200 as M
    z, kv = load_golden("beam_16x6x6_bending")
    exe = os.path.join(os.path.dirname(M.capi.LIB_PATH), "macroc")
    args = [a for kvp in kv.items() for a in kvp] + procs.split()
    procs_ = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MACROC_ID_FILE=str(tmp_path / "nccl_id"))
        procs_.append(subprocess.Popen([exe] + args, cwd=tmp_path, env=env, stdout=subprocess.PIPE,
                                       stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs_]
    assert all(p.returncode == 0 for p in procs_), outs[0][1][-2000:] + outs[1][1][-2000:]
    out = outs[0][0]
    ksp = [int(x) for x in re.findall(r"Its = (\d+)", out)]
    newton = [int(x) for x in re.findall(r"Newton Iteration = (\d+)", out)]
    res = [float(x) for x in re.findall(r"\|RES\| = (\S+)", out)]
    assert newton == list(z["newton_lines"])
    assert len(ksp) == len(z["ksp_its"]) and all(abs(a - int(b)) <= 1 for a, b in zip(ksp, z["ksp_its"]))
    assert res[1] == pytest.approx(float(z["res_norms"][1]), rel=1e-5)
    info = np.loadtxt(tmp_path / "info.dat", ndmin=2)
    assert np.allclose(info[:, 3], z["force"], rtol=1e-4, atol=1e-9)

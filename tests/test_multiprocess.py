"""Multi-process tests: world_size-2 on gloo here (host logic of the N>1 path), and the real
one-process-per-GPU run over NVLink peer memory when the box has >= 2 GPUs (gpurun --gpus 2)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER] + list(args)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    ok = [ln for ln in p.stdout.splitlines() if ln.startswith("MP_WORKER_OK ")]
    assert p.returncode == 0 and ok, "worker failed:\n" + p.stdout[-3000:] + "\n" + p.stderr[-6000:]
    return json.loads(ok[-1][len("MP_WORKER_OK "):])


@pytest.mark.parametrize("blocks,bc", [("1,1,2", "duct"), ("2,1,1", "periodic")])
def test_gloo_world2_host_path(blocks, bc):
    out = _run(2, "--mode", "host", "--cells", "16,12,20", "--blocks", blocks, "--bc", bc)
    assert out["world"] == 2 and out["bb"] > 0


def test_gloo_world4_host_path():
    out = _run(4, "--mode", "host", "--cells", "16,12,20", "--blocks", "2,2,1", "--bc", "cavity")
    assert out["world"] == 4


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("blocks,bc,nparts", [("1,1,2", "duct", 0), ("2,1,1", "channel", 0), ("1,2,1", "sedimentation", 3)])
def test_two_gpus_match_single_block_oracle(blocks, bc, nparts):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = _run(2, "--mode", "gpu", "--cells", "32,32,32", "--blocks", blocks, "--bc", bc, "--nparts", str(nparts))
    assert abs(out["niter"] - out["oracle_niter"]) <= 1


@pytest.mark.gpu
def test_four_and_eight_gpus_match_single_block_oracle():
    n = _ngpu()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    out = _run(4, "--mode", "gpu", "--cells", "32,32,32", "--blocks", "2,2,1", "--bc", "cavity")
    assert abs(out["niter"] - out["oracle_niter"]) <= 1
    if n >= 8:
        out = _run(8, "--mode", "gpu", "--cells", "32,32,32", "--blocks", "2,2,2", "--bc", "periodic")
        assert abs(out["niter"] - out["oracle_niter"]) <= 1

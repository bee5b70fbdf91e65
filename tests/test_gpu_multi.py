"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/dp_worker.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_data_parallel_training_parity(nproc):
    if _gpus() < nproc:
        pytest.skip("needs %d GPUs, the box has %d" % (nproc, _gpus()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + nproc), os.path.join(HERE, "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "dp parity OK on %d GPUs" % nproc in r.stdout


def test_sharded_fitness_parity():
    """cfg5 on 2 GPUs: the individuals to evaluate are sharded over ranks, only fitness values are gathered
    (SURVEY 8e row 2); every rank ends with the grades one rank computes alone."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs, the box has %d" % _gpus())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(HERE, "fitness_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "sharded fitness OK" in r.stdout

"""Brick-decomposed runs (one process per GPU, NCCL halo + migration inside libsphbvf.so) against
the same golden fixtures as the single-GPU parity tests.  Needs >= 2 GPUs; tests/mgpu_worker.py does
the work under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_bricks_match_reference(nranks):
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + nranks), os.path.join(HERE, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("mgpu ")))   # kept in the committed logs (pytest -s)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]

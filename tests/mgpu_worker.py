"""Multi-GPU parity worker (one process per GPU; launched by tests/test_multi_gpu.py or by hand):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/mgpu_worker.py [fixture ...]

Every rank takes the atoms of its brick from a golden fixture, the library migrates / halo-exchanges
over NCCL, and rank 0 reassembles the fields by atom tag and compares them with what the UNMODIFIED
reference produced on one CPU rank (tests/golden/*.npz): same bar as tests/test_gpu_parity.py --
pair sets bit-exact, consumed fields within 1e-10 of the field's max-norm.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from brick_parity import run_fixture as _run_fixture  # noqa: E402
from conftest import load_package  # noqa: E402


def run_fixture(pkg, name, rank, world, dev):
    errors, info, _, _ = _run_fixture(pkg, name, rank, world, dev, dist)
    return errors or [], info


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    pkg = load_package()
    names = sys.argv[1:] or ["cavity_n50", "synth3d_n14", "fsi_nx20", "yeast_nx40", "natconv_n40", "solid3d_mech_n10"]
    bad = []
    for name in names:
        errors, info = run_fixture(pkg, name, rank, world, dev)
        if rank == 0:
            print("mgpu %s: ranks %d grid %s nlocal[0] %d nghost[0] %d builds %d -> %s" % (info + ("OK" if not errors else errors[:6],)), flush=True)
        bad += errors
    flag = [len(bad)]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(1 if flag[0] else 0)


if __name__ == "__main__":
    main()

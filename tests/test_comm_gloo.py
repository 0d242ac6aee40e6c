"""World-size-2 (and 4) CPU test of the brick halo protocol, `gloo` backend.

The product exchanges over NCCL inside libsphbvf.so (csrc/comm_nccl.cu); what can be checked without
a GPU is the host-side logic it relies on: sphbvf_proc_grid / sphbvf_brick_bounds / sphbvf_comm_plan
(peers and periodic shifts for the 27 directions) and the message protocol -- per-direction counts
all-gathered, sends posted in ascending direction order and receives in descending order so that
several messages between the same two ranks pair up (two bricks in a periodic dimension are each
other's neighbour twice).  Each rank rebuilds its ghost shell with that protocol from a golden
fixture's atoms, and the neighbour pairs found on owned+ghost atoms must equal the reference's
pair list (bit-exact, as sorted tag pairs)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def worker(rank, world, port, name, q):
    from common import load_fixture
    from conftest import load_package
    from refsnap import canonical_pairs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    L = pkg.lib()
    meta, z = load_fixture(name)
    dim = meta["dim"]
    prd = np.array(meta["boxhi"]) - np.array(meta["boxlo"])
    grid = (C.c_int * 3)()
    L.sphbvf_proc_grid(world, dim, (C.c_double * 3)(*prd), C.byref(grid))
    cfg = pkg.config_from_meta(meta, procgrid=tuple(grid), rank=rank, nranks=world)
    sublo, subhi = (C.c_double * 3)(), (C.c_double * 3)()
    L.sphbvf_brick_bounds(C.byref(cfg), rank, C.byref(sublo), C.byref(subhi))
    peer, shift = (C.c_int * 27)(), (C.c_double * 81)()
    L.sphbvf_comm_plan(C.byref(cfg), rank, C.byref(peer), C.byref(shift))
    shift = np.array(shift).reshape(27, 3)
    x, tag, typ = z["init_x"], z["init_tag"], z["init_type"]
    mine = np.ones(len(x), bool)
    for k in range(dim):
        mine &= (x[:, k] >= sublo[k]) & (x[:, k] < subhi[k])
    xo, to, ty = x[mine], tag[mine], typ[mine]
    # cutneigh per type pair as Neighbor::init builds it (neighbor.cpp:296-310)
    nt = meta["ntypes"]
    cut = np.zeros((nt + 1, nt + 1))
    for p in meta["pairs"]:
        cut[p["i"], p["j"]] = cut[p["j"], p["i"]] = p["h"] + meta["skin"]
    cutghost = cut.max()
    # border lists (CommBrick::borders slab rule, all 26 directions at once)
    send = {}
    for d in range(27):
        if peer[d] < 0:
            continue
        s = (d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1)
        m = np.ones(len(xo), bool)
        for k in range(3):
            if s[k] == 1:
                m &= xo[:, k] >= subhi[k] - cutghost
            elif s[k] == -1:
                m &= xo[:, k] <= sublo[k] + cutghost
        send[d] = np.nonzero(m)[0]
    counts = torch.zeros(27, dtype=torch.int64)
    for d, idx in send.items():
        counts[d] = len(idx)
    table = [torch.zeros(27, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(table, counts)
    reqs, recv = [], {}
    for d in range(27):                      # sends ascending
        if d in send and len(send[d]) and peer[d] != rank:
            buf = np.concatenate([xo[send[d]] + shift[d], to[send[d], None].astype(float), ty[send[d], None].astype(float)], axis=1)
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(buf)), dst=peer[d]))
    for d in range(26, -1, -1):              # receives descending
        if peer[d] < 0:
            continue
        n = int(table[peer[d]][26 - d])
        if not n:
            continue
        if peer[d] == rank:                  # own periodic image: device copy in the product
            src = 26 - d
            recv[d] = np.concatenate([xo[send[src]] + shift[src], to[send[src], None].astype(float), ty[send[src], None].astype(float)], axis=1)
        else:
            t = torch.zeros((n, 5), dtype=torch.float64)
            dist.recv(t, src=peer[d])
            recv[d] = t.numpy()
    for r in reqs:
        r.wait()
    ghosts = np.concatenate([recv[d] for d in sorted(recv)]) if recv else np.zeros((0, 5))
    xa = np.concatenate([xo, ghosts[:, :3]])
    ta = np.concatenate([to, ghosts[:, 3].astype(np.int64)])
    tya = np.concatenate([ty, ghosts[:, 4].astype(np.int64)])
    # pairs (owned i, any j) with rsq <= cutneighsq, each unordered tag pair once
    pairs = []
    for i in range(len(xo)):
        dx = xa - xo[i]
        if dim == 2:
            dx[:, 2] = 0.0
        rsq = dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1] + dx[:, 2] * dx[:, 2]
        c = cut[ty[i], tya]
        hit = np.nonzero((rsq <= c * c) & (np.arange(len(xa)) != i))[0]
        for j in hit:
            if j < len(xo):
                if i < j:
                    pairs.append((to[i], ta[j]))
            elif to[i] < ta[j] or (to[i] == ta[j] and tuple(xa[j] - xo[i])[::-1] > (0, 0, 0)):
                pairs.append((to[i], ta[j]))
    allp = [None] * world
    dist.all_gather_object(allp, np.array(pairs, dtype=np.int64).reshape(-1, 2))
    if rank == 0:
        got = canonical_pairs(np.concatenate(allp))
        ref = z["p0"]
        if meta["variant"] == 2:
            ref = np.unique(ref, axis=0)
        q.put((name, world, tuple(grid), bool(got.shape == ref.shape and np.array_equal(got, ref)), got.shape[0], ref.shape[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("cavity_n20", 2), ("fsi_nx20", 2), ("yeast_nx40", 2), ("synth3d_n14_lattice", 2),
                                        ("fsi_nx20", 4)])
def test_halo_protocol_rebuilds_reference_pairs(name, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (hash((name, world)) % 200)
    procs = [ctx.Process(target=worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = q.get(timeout=10)
    assert res[3], res

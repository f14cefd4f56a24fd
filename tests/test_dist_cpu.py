"""CPU tests of the multi-GPU host logic: block-cyclic layout helpers and the rendezvous path with a
world_size-2 gloo group (the NCCL data path itself is exercised by tests/dist_check.py on GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def test_block_cyclic_roundtrip():
    from svdsolver_b200 import distributed as D
    rng = np.random.default_rng(0)
    for n, band, P in ((64, 4, 3), (128, 32, 2), (96, 8, 5), (64, 64, 2)):
        a = rng.normal(size=(n, n))
        parts = [D.scatter_block_cyclic(a, band, r, P) for r in range(P)]
        assert sum(p.shape[1] for p in parts) == n
        for r, p in enumerate(parts):
            assert p.shape[1] == D.local_cols(n, band, r, P)
        assert np.array_equal(D.gather_block_cyclic(parts, band, n), a)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from svdsolver_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, band = 96, 8
    a = np.arange(n * n, dtype=np.float64).reshape(n, n)
    loc = D.scatter_block_cyclic(a, band, rank, world)
    # the same gather the GPU check performs, on CPU tensors over gloo
    wmax = max(D.local_cols(n, band, r, world) for r in range(world))
    pad = torch.zeros(n, wmax, dtype=torch.float64)
    pad[:, : loc.shape[1]] = torch.from_numpy(loc)
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    full = D.gather_block_cyclic([o[:, : D.local_cols(n, band, r, world)].numpy() for r, o in enumerate(outs)], band, n)
    # unique-id exchange path: without NCCL devices rank 0 cannot create an id; the broadcast of the
    # 128-byte payload itself is what is covered here
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    q.put((rank, bool(np.array_equal(full, a)), obj[0] == bytes(range(128))))
    dist.destroy_process_group()


def test_gloo_world2_layout_and_rendezvous():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res


def test_unpack_band_layout():
    """packed band storage of the stage-2 hand-off: packed[gc, t] = A[gc - band + t, gc]"""
    from svdsolver_b200 import distributed as D
    rng = np.random.default_rng(1)
    for n, b in ((50, 4), (64, 32), (33, 1)):
        a = np.triu(np.tril(rng.normal(size=(n, n)), b))
        packed = np.zeros((n, b + 1))
        for gc in range(n):
            for t in range(b + 1):
                r = gc - b + t
                if r >= 0:
                    packed[gc, t] = a[r, gc]
        assert np.array_equal(D.unpack_band(packed, n, b), a)

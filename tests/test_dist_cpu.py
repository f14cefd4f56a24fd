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


# ---- the distributed row panel of csrc/dist.cu, restated in numpy over a world_size-2 gloo group ------------------------------
def _lq_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import torch
    import torch.distributed as dist
    from panel_chol_model import chol_panel
    from svdsolver_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, b, k = 160, 8, 2                                           # block step k: the row panel is rows k*b .. k*b+b-1, columns (k+1)*b ..
    a = np.random.default_rng(5).random((n, n)) * 5
    o = k * b
    loc = D.scatter_block_cyclic(a, b, rank, world)
    # local trailing columns of this rank = local blocks with global index > k; panel coordinates: rows <-> local columns
    blocks = [j for j in D.owned_blocks(n, b, rank, world) if j > k]
    cols = np.concatenate([np.arange(j * b, (j + 1) * b) for j in blocks])
    first = D.owned_blocks(n, b, rank, world).index(blocks[0]) * b
    p_loc = loc[o:o + b, first:].T.copy()                          # (ncl x b): my rows of the panel
    own_top = (k + 1) % world == rank                              # my first b rows are the panel's top block
    # every rank: Gram matrix of its rows; the owner also contributes the top block; ONE all-reduce
    msg = np.zeros((2 * b, b))
    msg[:b] = p_loc.T @ p_loc
    if own_top:
        msg[b:] = p_loc[:b]
    t = torch.from_numpy(msg)
    dist.all_reduce(t)
    g, a1 = t.numpy()[:b], t.numpy()[b:]
    # the same b x b algebra on every rank (tools/panel_chol_model.py restates it on a full panel; here from G and A1 only)
    r = np.linalg.cholesky(g).T
    w = a1.copy(); s = np.zeros(b)
    for i in range(b):
        s[i] = -np.copysign(1.0, w[i, i]); w[i, i:] -= s[i] * r[i, i:]
        w[i + 1:, i] /= w[i, i]; w[i + 1:, i + 1:] -= np.outer(w[i + 1:, i], w[i, i + 1:])
    l = np.tril(w, -1) + np.eye(b); ut = np.triu(w)
    m1 = np.linalg.inv(ut)
    tmat = np.linalg.inv(-(l.T * s) @ (r @ m1))
    m2 = -m1 @ tmat.T
    # local second pass: my rows of U^T (= Y) and of V2
    y_loc = p_loc @ m1
    v2_loc = p_loc @ m2
    if own_top:
        y_loc[:b] = l
        v2_loc[:b] = -l @ tmat.T
    # reference: the whole panel factorised in one place
    full = a[o:o + b, (k + 1) * b:].T
    rhh, y, tau, tfull, v2, ratio = chol_panel(full, np.float64, False)
    rows = cols - (k + 1) * b
    ok = (np.abs(y_loc - y[rows]).max() < 1e-12 and np.abs(v2_loc - v2[rows]).max() < 1e-12 and np.abs(s[:, None] * r - rhh).max() < 1e-11
          and ratio > 1e-3)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_row_panel_by_local_gram_matrices_world2():
    """LQ half-step of the multi-GPU stage 1 (csrc/dist.cu): local Gram matrices + the owner's top block through ONE all-reduce,
    identical b x b algebra on every rank, local second pass == rows of the panel factorised in one place."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lq_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]

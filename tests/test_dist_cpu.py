"""CPU suite, part 3: the N > 1 path over gloo, world_size 2.

The product's exchange steps are (a) the all-gather of every partition's feature
slab before ScatterGather and (b) the all-reduce of dW.  Here two CPU processes
play two partitions: bounds come from the product's own roc_partition (host code,
no GPU needed), the data moves over torch.distributed (gloo) with exactly the
offsets / counts the C++ host uses, and the per-partition math is the oracle's.
The stitched result must equal the single-partition oracle bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from roc_b200 import _lib, datasets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    re_t, col_t = datasets.rmat_graph(9, 3000, seed=11)
    row_end = re_t.numpy().astype(np.uint64)
    col = col_t.numpy().astype(np.uint32)
    n, e = row_end.shape[0], int(row_end[-1])
    vb = np.zeros((world, 2), dtype=np.uint32)
    eb = np.zeros((world, 2), dtype=np.uint64)
    nr = C.c_int(0)
    assert _lib.lib.roc_partition(n, e, world, row_end.ctypes.data, vb.ctypes.data, eb.ctypes.data,
                                  C.cast(C.byref(nr), C.c_void_p)) == 0
    rl, rr, cl, cr = int(vb[rank, 0]), int(vb[rank, 1]), int(eb[rank, 0]), int(eb[rank, 1])
    h = 12
    x_full = np.random.RandomState(5).randn(n, h).astype(np.float32)
    w = np.random.RandomState(6).randn(3, h).astype(np.float32)
    mine = torch.from_numpy(x_full[rl:rr + 1].copy())
    # (a) all-gather-v by one broadcast per owner into the [N][H] matrix (Comm::allgatherv)
    gathered = torch.zeros((n, h), dtype=torch.float32)
    for r in range(world):
        a, b = int(vb[r, 0]), int(vb[r, 1])
        buf = mine if r == rank else torch.empty((b - a + 1, h), dtype=torch.float32)
        dist.broadcast(buf, src=r)
        gathered[a:b + 1] = buf
    assert np.array_equal(gathered.numpy(), x_full)
    y = oracle.scatter_gather(rl, rr, cl, row_end[rl:rr + 1], col[cl:cr + 1], gathered.numpy())
    # (b) dW replica all-reduce
    dy = np.random.RandomState(7).randn(n, 3).astype(np.float32)[rl:rr + 1]
    dw = np.zeros_like(w)
    oracle.linear_bwd(x_full[rl:rr + 1], w, None, dy.copy(), dw, need_dx=False)
    t = torch.from_numpy(dw)
    dist.all_reduce(t)
    np.save(os.path.join(tmpdir, "y%d.npy" % rank), y)
    np.save(os.path.join(tmpdir, "dw%d.npy" % rank), t.numpy())
    dist.destroy_process_group()


def test_two_partitions_over_gloo(tmp_path):
    from oracle import oracle
    from roc_b200 import datasets
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    re_t, col_t = datasets.rmat_graph(9, 3000, seed=11)
    row_end = re_t.numpy().astype(np.uint64)
    col = col_t.numpy().astype(np.uint32)
    n = row_end.shape[0]
    x_full = np.random.RandomState(5).randn(n, 12).astype(np.float32)
    want = oracle.scatter_gather(0, n - 1, 0, row_end, col, x_full)
    got = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")])
    assert np.array_equal(got, want)
    w = np.random.RandomState(6).randn(3, 12).astype(np.float32)
    dy = np.random.RandomState(7).randn(n, 3).astype(np.float32)
    dw = np.zeros_like(w)
    oracle.linear_bwd(x_full, w, None, dy.copy(), dw, need_dx=False)
    d0, d1 = np.load(tmp_path / "dw0.npy"), np.load(tmp_path / "dw1.npy")
    assert np.array_equal(d0, d1)
    assert np.allclose(d0, dw, rtol=1e-5, atol=1e-5)

"""CPU suite, part 3: the N > 1 path over gloo, world_size 2 and 4.

The product's exchange steps are (a) the halo exchange before every ScatterGather — each partition
receives the distinct remote rows its edges read, packed by their owners — and (b) the all-reduce of
dW.  The GPU builds the halo id list (roc_halo_create); everything after that is HOST bookkeeping in
the product's own C ABI: roc_partition (bounds), roc_halo_recv_layout (which owner holds which slab of
the sorted halo), roc_halo_send_layout (what this rank packs for whom, from the all-gathered P x P
request matrix).  Here CPU processes play the partitions: the id list is the definition restated in
numpy (sorted distinct sources outside the own range — what tests/test_kernels_gpu.py checks the GPU
builder against), the counts / offsets come from the PRODUCT's functions, the ids and rows travel over
torch.distributed (gloo) with exactly those counts / offsets, and the per-partition math is the
oracle's on [own rows | halo rows] with the remapped col.  The stitched result must equal the
single-partition oracle bit for bit, and every rank's dW all-reduce must agree."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

H = 12


def _graph():
    from roc_b200 import datasets
    re_t, col_t = datasets.rmat_graph(10, 7000, seed=11)
    return re_t.numpy().astype(np.uint64), col_t.numpy().astype(np.uint32)


def _all_to_all(outs, ins, rank, world):
    """all-to-all-v by point-to-point pairs (gloo has no alltoall): what Comm::alltoallv does with grouped
    ncclSend / ncclRecv — no self transfer, empty messages skipped on both sides."""
    reqs = []
    for q in range(world):
        if q == rank:
            continue
        if outs[q].numel():
            reqs.append(dist.isend(outs[q], dst=q))
        if ins[q].numel():
            reqs.append(dist.irecv(ins[q], src=q))
    for r in reqs:
        r.wait()


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from roc_b200 import _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    row_end, col = _graph()
    n, e = row_end.shape[0], int(row_end[-1])
    vb = np.zeros((world, 2), dtype=np.uint32)
    eb = np.zeros((world, 2), dtype=np.uint64)
    nr = C.c_int(0)
    assert _lib.lib.roc_partition(n, e, world, row_end.ctypes.data, vb.ctypes.data, eb.ctypes.data,
                                  C.cast(C.byref(nr), C.c_void_p)) == 0
    rl, rr, cl, cr = int(vb[rank, 0]), int(vb[rank, 1]), int(eb[rank, 0]), int(eb[rank, 1])
    nloc = rr - rl + 1
    my_col = col[cl:cr + 1]
    # the halo: sorted distinct sources outside [rl, rr] (roc_halo_create's definition)
    remote = (my_col < rl) | (my_col > rr)
    ids = np.unique(my_col[remote]).astype(np.uint32)
    # ---- product host bookkeeping: who owns which slab of the halo
    rc_, ro_ = np.zeros(world, dtype=np.uint64), np.zeros(world, dtype=np.uint64)
    assert _lib.lib.roc_halo_recv_layout(ids.shape[0], ids.ctypes.data, world, rank, vb.ctypes.data,
                                         rc_.ctypes.data, ro_.ctypes.data) == 0
    assert int(rc_.sum()) == ids.shape[0] and rc_[rank] == 0
    # everyone learns everyone's request counts (Graph::build's allgather_i32)
    mine = torch.from_numpy(rc_.astype(np.int32))
    allc = [torch.zeros(world, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(allc, mine)
    allc = np.stack([a.numpy() for a in allc]).astype(np.int32)           # [q][r] = rows q requests from r
    sc_, so_ = np.zeros(world, dtype=np.uint64), np.zeros(world, dtype=np.uint64)
    tot = C.c_uint64(0)
    assert _lib.lib.roc_halo_send_layout(world, rank, allc.ctypes.data, sc_.ctypes.data, so_.ctypes.data,
                                         C.cast(C.byref(tot), C.c_void_p)) == 0
    # ---- the id lists travel to their owners (Comm::alltoallv of u32), owners turn them into local rows
    send_ids = [torch.from_numpy(ids[int(ro_[q]):int(ro_[q] + rc_[q])].astype(np.int32)) for q in range(world)]
    recv_ids = [torch.zeros(int(sc_[q]), dtype=torch.int32) for q in range(world)]
    _all_to_all(send_ids, recv_ids, rank, world)
    send_rows = np.concatenate([r.numpy() for r in recv_ids]).astype(np.int64) if tot.value else np.zeros(0, np.int64)
    assert send_rows.shape[0] == tot.value
    assert ((send_rows >= rl) & (send_rows <= rr)).all()                   # Graph::build asserts the same
    send_rows -= rl
    for q in range(world):                                                  # packed per requester, in order
        assert int(so_[q]) == int(sc_[:q].sum())
    # ---- the exchange itself: pack (roc_pack_rows) -> all-to-all-v -> halo slab behind the own rows
    x_full = np.random.RandomState(5).randn(n, H).astype(np.float32)
    x_loc = x_full[rl:rr + 1]
    packed = x_loc[send_rows]
    out_chunks = [torch.from_numpy(packed[int(so_[q]):int(so_[q] + sc_[q])].copy()) for q in range(world)]
    in_chunks = [torch.zeros((int(rc_[q]), H), dtype=torch.float32) for q in range(world)]
    _all_to_all(out_chunks, in_chunks, rank, world)
    halo = np.concatenate([c.numpy() for c in in_chunks]) if ids.shape[0] else np.zeros((0, H), np.float32)
    assert np.array_equal(halo, x_full[ids])                               # the slab holds exactly the rows the edges read
    # ---- ScatterGather on [own rows | halo rows] with the remapped col (roc_halo_col_local's definition)
    col_local = np.where(remote, nloc + np.searchsorted(ids, my_col), my_col.astype(np.int64) - rl).astype(np.uint32)
    slab = np.concatenate([x_loc, halo])
    y = oracle.scatter_gather(0, nloc - 1, 0, row_end[rl:rr + 1] - np.uint64(cl), col_local, slab)
    # ---- (b) dW replica all-reduce
    w = np.random.RandomState(6).randn(3, H).astype(np.float32)
    dy = np.random.RandomState(7).randn(n, 3).astype(np.float32)[rl:rr + 1]
    dw = np.zeros_like(w)
    oracle.linear_bwd(x_loc, w, None, dy.copy(), dw, need_dx=False)
    t = torch.from_numpy(dw)
    dist.all_reduce(t)
    np.save(os.path.join(tmpdir, "y%d.npy" % rank), y)
    np.save(os.path.join(tmpdir, "dw%d.npy" % rank), t.numpy())
    np.save(os.path.join(tmpdir, "halo%d.npy" % rank), np.array([ids.shape[0], int(tot.value)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_over_gloo(tmp_path, world):
    from oracle import oracle
    port = 29500 + (os.getpid() * 7 + world) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    row_end, col = _graph()
    n = row_end.shape[0]
    x_full = np.random.RandomState(5).randn(n, H).astype(np.float32)
    want = oracle.scatter_gather(0, n - 1, 0, row_end, col, x_full)
    got = np.concatenate([np.load(tmp_path / ("y%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, want)
    w = np.random.RandomState(6).randn(3, H).astype(np.float32)
    dy = np.random.RandomState(7).randn(n, 3).astype(np.float32)
    dw = np.zeros_like(w)
    oracle.linear_bwd(x_full, w, None, dy.copy(), dw, need_dx=False)
    ds = [np.load(tmp_path / ("dw%d.npy" % r)) for r in range(world)]
    for d in ds[1:]:
        assert np.array_equal(ds[0], d)
    assert np.allclose(ds[0], dw, rtol=1e-5, atol=1e-5)
    halos = [np.load(tmp_path / ("halo%d.npy" % r)) for r in range(world)]
    assert sum(int(h[0]) for h in halos) == sum(int(h[1]) for h in halos) > 0     # every requested row is sent once


def test_halo_layout_rejects_bad_lists():
    from roc_b200 import _lib
    vb = np.array([[0, 9], [10, 19]], dtype=np.uint32)
    rc_, ro_ = np.zeros(2, dtype=np.uint64), np.zeros(2, dtype=np.uint64)

    def call(ids, me):
        a = np.asarray(ids, dtype=np.uint32)
        return _lib.lib.roc_halo_recv_layout(a.shape[0], a.ctypes.data, 2, me, vb.ctypes.data, rc_.ctypes.data, ro_.ctypes.data)
    assert call([10, 12, 19], 0) == 0 and list(rc_) == [0, 3] and list(ro_) == [0, 0]
    assert call([], 1) == 0 and list(rc_) == [0, 0]
    assert call([3, 12], 0) == _lib.ROC_ERR_INVALID          # names an own row
    assert call([12, 11], 0) == _lib.ROC_ERR_INVALID         # not sorted
    assert call([12, 12], 0) == _lib.ROC_ERR_INVALID         # duplicate
    assert call([12, 25], 0) == _lib.ROC_ERR_INVALID         # outside every range
    allc = np.array([[0, 2], [5, 0]], dtype=np.int32)
    sc_, so_ = np.zeros(2, dtype=np.uint64), np.zeros(2, dtype=np.uint64)
    assert _lib.lib.roc_halo_send_layout(2, 0, allc.ctypes.data, sc_.ctypes.data, so_.ctypes.data, None) == 0
    assert list(sc_) == [0, 5] and list(so_) == [0, 0]
    bad = np.array([[1, 2], [5, 0]], dtype=np.int32)
    assert _lib.lib.roc_halo_send_layout(2, 0, bad.ctypes.data, sc_.ctypes.data, so_.ctypes.data, None) == _lib.ROC_ERR_INVALID

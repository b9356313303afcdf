"""GPU suite, part 3 (needs >= 2 GPUs; skipped where the box has fewer than the case wants): the vertex-range
partitioned engine — one process per GPU, the halo exchange before every ScatterGather (peer writes over
NVLink pipelined with the producer, or staged rows + NCCL with ROC_B200_HALO=nccl) and the dW all-reduce over
NCCL — must reproduce the single-partition oracle at 2, 4 and 8 ranks: logits of every partition, the summed
dW, the weights after several steps.  (bench.py --gpus N runs the same kind of check on the driver's N-GPU
box and prints it as `parity_check`.)"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_close

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

LAYERS = (24, 32, 16, 7)      # more than 3 dims: exercises the residual branch too
EPOCHS = 3


def _case():
    from roc_b200 import datasets
    re_t, col_t = datasets.rmat_graph(11, 12000, seed=21)
    row_end, col = re_t.numpy().astype(np.uint64), col_t.numpy().astype(np.uint32)
    feats, labels, mask = datasets.node_data(row_end.shape[0], LAYERS[0], LAYERS[-1], seed=5)
    return row_end, col, feats.numpy(), labels.numpy(), mask.numpy()


def _worker(rank, world, port, tmpdir, dropout):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from roc_b200.model import Host, Model, build_gcn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    row_end, col, feats, labels, mask = _case()
    host = Host(rank, rank, world)
    uid = [Host.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    host.nccl_init(uid[0])
    host.graph_from_arrays(row_end, col)
    info = host.graph_info()
    rl, rr = info["rowLeft"], info["rowRight"]
    m = Model(host, seed=1)
    h = build_gcn(m, list(LAYERS), dropout)
    m.set_tensor(h["input"], feats[rl:rr + 1])
    m.set_labels(h["label"], labels[rl:rr + 1])
    m.set_tensor(h["mask"], mask[rl:rr + 1].astype(np.int32))
    out = {"rl": rl, "rr": rr, "w0": [m.get_parameter(p) for p in range(m.num_parameters())]}
    for ep in range(EPOCHS):
        m.train_mode(); m.zero_gradients(); m.forward()
        if ep == 0:
            out["logits"] = m.get_tensor(h["logits"])
        out.setdefault("relu", []).append([m.get_tensor(t) > 0 for t in h["relu_outs"]])
        m.backward()
        if ep == 0:
            out["dW_local"] = [m.get_parameter(p, "grad") for p in range(m.num_parameters())]
        out.setdefault("perf", []).append(m.metrics())
        m.update()
        if ep == 0:
            out["dW_reduced"] = [m.get_parameter(p, "grad") for p in range(m.num_parameters())]
    out["w"] = [m.get_parameter(p) for p in range(m.num_parameters())]
    np.save(os.path.join(tmpdir, "rank%d.npy" % rank), np.array([out], dtype=object), allow_pickle=True)
    host.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
@pytest.mark.parametrize("dropout", [0.0, 0.5])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_training_matches_oracle(tmp_path, world, dropout, halo, monkeypatch):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    if halo == "nccl" and (world != 2 or dropout != 0.5):
        pytest.skip("the NCCL exchange is cross-checked once")
    import torch.multiprocessing as mp
    from oracle import oracle
    from roc_b200 import datasets
    from test_model_gpu import sync_relu_masks
    if halo == "nccl":
        monkeypatch.setenv("ROC_B200_HALO", "nccl")
    else:
        monkeypatch.delenv("ROC_B200_HALO", raising=False)
    port = 29700 + (os.getpid() * 3 + world) % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path), dropout), nprocs=world, join=True)
    r = [np.load(tmp_path / ("rank%d.npy" % k), allow_pickle=True)[0] for k in range(world)]
    row_end, col, feats, labels, mask = _case()
    n = row_end.shape[0]
    k, vb, _ = oracle.partition(row_end, world)
    assert k == world
    for q in range(world):
        assert (r[q]["rl"], r[q]["rr"]) == (int(vb[q, 0]), int(vb[q, 1]))      # bit-exact partition bounds
    assert r[world - 1]["rr"] == n - 1
    for q in range(1, world):
        for a, b in zip(r[0]["w0"], r[q]["w0"]):
            assert np.array_equal(a, b)               # same srand seed -> identical Glorot weights on every rank
    o = oracle.GcnOracle(row_end, col, LAYERS, r[0]["w0"], lr=0.01, weight_decay=0.05, dropout=dropout)
    oh = datasets.onehot(labels, LAYERS[-1])
    for ep in range(EPOCHS):
        o.forward(feats, train=True)
        masks = [np.concatenate([r[q]["relu"][ep][i] for q in range(world)]) for i in range(len(r[0]["relu"][ep]))]
        sync_relu_masks(o, masks)
        if ep == 0:
            rel_close(np.concatenate([r[q]["logits"] for q in range(world)]), o.logits, what="stitched logits")
        o.backward(oh, mask)
        if ep == 0:
            for p in range(len(o.dW)):
                rel_close(sum(r[q]["dW_local"][p] for q in range(world)), o.dW[p], rtol=2e-4, what="sum of dW replicas")
                for q in range(1, world):
                    assert np.array_equal(r[0]["dW_reduced"][p], r[q]["dW_reduced"][p])   # all-reduce: same bits everywhere
                rel_close(r[0]["dW_reduced"][p], o.dW[p], rtol=2e-4, what="all-reduced dW")
        tot = sum(r[q]["perf"][ep]["trainAll"] for q in range(world))
        assert tot == o.perf["trainAll"]
        loss = sum(r[q]["perf"][ep]["trainLoss"] for q in range(world))
        assert abs(loss - o.perf["trainLoss"]) <= 2e-4 * abs(o.perf["trainLoss"])
        o.update()
    for p in range(len(o.W)):
        for q in range(1, world):
            assert np.array_equal(r[0]["w"][p], r[q]["w"][p])
        rel_close(r[0]["w"][p], o.W[p], rtol=1e-3, atol_scale=1e-4, what="W[%d] after %d epochs" % (p, EPOCHS))

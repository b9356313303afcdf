"""make_golden.py — mint golden vectors from the REFERENCE's own kernels.

The reference (jiazhihao/ROC) has no tests and no golden vectors, and
its full binary cannot be built (Legion absent).  Its CUDA kernels can: oracle/
Makefile cuts them out of /root/reference into oracle/_ref/libroc_ref.so.  This
script runs those kernels (plus the cuBLAS / cuDNN / cuRAND calls the reference
makes, same arguments) on a B200 over small seeded inputs and stores inputs and
outputs in tests/golden/ref_golden.npz.  The CPU oracle and the product kernels
are then both checked against these files.

Run on a GPU box (needs the prebuilt oracle/_ref/libroc_ref.so):
    python tests/golden/make_golden.py [outdir]       # default: gpurun_out/golden
and copy the .npz into tests/golden/.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from roc_b200 import datasets  # noqa: E402


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    dev = "cuda"
    rng = np.random.RandomState(20260921)
    G = {}

    # ---- graph A: cfg-1 shaped (1 000 vertices, ~10 K edges), whole graph = one partition
    re_t, col_t = datasets.uniform_graph(1000, 4500, seed=1)
    row_end = re_t.numpy().astype(np.uint64)
    col = col_t.numpy().astype(np.uint32)
    N, E = row_end.shape[0], col.shape[0]
    G["A_row_end"], G["A_col"] = row_end, col
    d_rows = torch.from_numpy(row_end.astype(np.int64)).to(dev)
    d_cols = torch.from_numpy(col.astype(np.int32)).to(dev)
    rp, es = ref.edge_structs(d_cols, d_rows, 0, 0)
    G["A_rowptrs"] = rp.cpu().numpy().astype(np.uint64)
    G["A_edgestructs"] = es.cpu().numpy().astype(np.uint32)
    for H in (16, 41, 64):
        x = (rng.rand(N, H).astype(np.float32) * 2 - 1)
        dx = torch.from_numpy(x).to(dev)
        y = ref.scatter_gather(0, N - 1, 0, rp, es, dx)
        G["A_sg_in_%d" % H] = x
        G["A_sg_out_%d" % H] = y.cpu().numpy()
        z = ref.indegree_norm(0, N - 1, 0, rp, dx)
        G["A_norm_out_%d" % H] = z.cpu().numpy()

    # ---- graph A split in two partitions: pins rowLeft / colLeft conventions
    from oracle import oracle
    k, vb, eb = oracle.partition(row_end, 2)
    assert k == 2
    G["A_vb2"], G["A_eb2"] = vb, eb
    x = G["A_sg_in_16"]
    dx = torch.from_numpy(x).to(dev)
    for c in range(2):
        rl, rr, cl, cr = int(vb[c, 0]), int(vb[c, 1]), int(eb[c, 0]), int(eb[c, 1])
        rows_c = d_rows[rl:rr + 1].contiguous()
        cols_c = d_cols[cl:cr + 1].contiguous()
        rp_c, es_c = ref.edge_structs(cols_c, rows_c, rl, cl)
        G["A_p%d_edgestructs" % c] = es_c.cpu().numpy().astype(np.uint32)
        y = ref.scatter_gather(rl, rr, cl, rp_c, es_c, dx)
        G["A_p%d_sg_out_16" % c] = y.cpu().numpy()
        z = ref.indegree_norm(rl, rr, cl, rp_c, dx[rl:rr + 1].contiguous())
        G["A_p%d_norm_out_16" % c] = z.cpu().numpy()

    # ---- graph B: skewed (R-MAT scale 10) with hub rows
    re_t, col_t = datasets.rmat_graph(10, 8192, seed=3)
    row_end = re_t.numpy().astype(np.uint64)
    col = col_t.numpy().astype(np.uint32)
    N = row_end.shape[0]
    G["B_row_end"], G["B_col"] = row_end, col
    d_rows = torch.from_numpy(row_end.astype(np.int64)).to(dev)
    d_cols = torch.from_numpy(col.astype(np.int32)).to(dev)
    rp, es = ref.edge_structs(d_cols, d_rows, 0, 0)
    x = (rng.rand(N, 32).astype(np.float32) * 2 - 1)
    y = ref.scatter_gather(0, N - 1, 0, rp, es, torch.from_numpy(x).to(dev))
    G["B_sg_in_32"], G["B_sg_out_32"] = x, y.cpu().numpy()

    # ---- linear (cuBLAS sgemm with the reference's arguments)
    n, i_dim, o_dim = 200, 33, 9
    X = (rng.rand(n, i_dim).astype(np.float32) * 2 - 1)
    W = (rng.rand(o_dim, i_dim).astype(np.float32) * 2 - 1)   # W_mem[o*in + i]
    dY = (rng.rand(n, o_dim).astype(np.float32) * 2 - 1)
    dX, dW_, dYt = torch.from_numpy(X).to(dev), torch.from_numpy(W).to(dev), torch.from_numpy(dY).to(dev)
    for relu in (0, 1):
        Y = ref.linear_fwd(dX, dW_, relu=bool(relu))
        G["lin_Y_relu%d" % relu] = Y.cpu().numpy()
        gw = torch.zeros_like(dW_)
        gx = torch.zeros_like(dX)
        gy = dYt.clone()
        ref.linear_bwd(dX, dW_, Y, gy, gw, gx, relu=bool(relu))
        torch.cuda.synchronize()
        G["lin_dW_relu%d" % relu], G["lin_dX_relu%d" % relu] = gw.cpu().numpy(), gx.cpu().numpy()
        G["lin_dY_after_relu%d" % relu] = gy.cpu().numpy()
    G["lin_X"], G["lin_W"], G["lin_dY"] = X, W, dY

    # ---- activations (cuDNN)
    a = (rng.randn(64, 24).astype(np.float32))
    da = torch.from_numpy(a).to(dev)
    gy = (rng.randn(64, 24).astype(np.float32))
    G["act_x"], G["act_dy"] = a, gy
    for mode, nm in ((1, "relu"), (2, "sigmoid")):
        y = ref.activation_fwd(da, mode)
        G["act_%s_y" % nm] = y.cpu().numpy()
        dx0 = torch.full_like(da, 0.25)   # beta = 1: accumulates onto what is there
        ref.activation_bwd(da, y, torch.from_numpy(gy).to(dev), dx0, mode)
        torch.cuda.synchronize()
        G["act_%s_dx_acc" % nm] = dx0.cpu().numpy()

    # ---- softmax cross entropy backward + metrics
    n, c = 500, 7
    logits = (rng.randn(n, c).astype(np.float32) * 2)
    lab = rng.randint(0, c, size=n).astype(np.int32)
    mask = rng.randint(0, 4, size=n).astype(np.int32)
    oh = datasets.onehot(lab, c)
    g, perf = ref.softmax_xent_bwd(torch.from_numpy(logits).to(dev), torch.from_numpy(oh).to(dev),
                                   torch.from_numpy(mask).to(dev))
    G["sm_logits"], G["sm_labels"], G["sm_mask"], G["sm_grad"] = logits, lab, mask, g.cpu().numpy()
    G["sm_perf"] = np.array([perf["trainLoss"], perf["trainAll"], perf["testAll"], perf["valAll"],
                             perf["trainCorrect"], perf["testCorrect"], perf["valCorrect"]], dtype=np.float64)

    # ---- adam with 3 gradient replicas
    cnt = 1000
    w = rng.randn(cnt).astype(np.float32)
    gr = rng.randn(3, cnt).astype(np.float32)
    m = (rng.randn(cnt).astype(np.float32) * 0.1)
    v = np.abs(rng.randn(cnt).astype(np.float32) * 0.1)
    dw, dg, dm, dv = (torch.from_numpy(t.copy()).to(dev) for t in (w, gr, m, v))
    ref.adam_update(dw, dg, dm, dv, 0.01, 0.9, 0.999, 0.05, 1e-8)
    torch.cuda.synchronize()
    G["adam_w"], G["adam_g"], G["adam_m"], G["adam_v"] = w, gr, m, v
    G["adam_w_out"], G["adam_m_out"], G["adam_v_out"] = dw.cpu().numpy(), dm.cpu().numpy(), dv.cpu().numpy()
    G["adam_gsum"] = dg[0].cpu().numpy()

    # ---- element add
    G["add_out"] = ref.add_fwd(torch.from_numpy(a).to(dev), torch.from_numpy(gy).to(dev)).cpu().numpy()

    # ---- Glorot init (cuRAND XORWOW seeded like initializer_kernel.cu:40-48)
    # glibc: srand(1); rand() -> 1804289383, 846930886 (gnn.cc:56 + initializer.cc:38)
    for seed, (i_dim, o_dim) in ((1804289383, (16, 16)), (846930886, (16, 5)), (1804289383, (602, 64))):
        G["glorot_%d_%dx%d" % (seed, i_dim, o_dim)] = ref.glorot(i_dim, o_dim, seed).cpu().numpy()

    path = os.path.join(outdir, "ref_golden.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, "with", len(G), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

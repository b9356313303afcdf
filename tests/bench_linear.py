"""bench_linear.py — micro-benchmark of the Linear GEMMs alone (not a pytest file).

Forward (Y = X W^T) and dW (= dY^T X) at the bench shapes, CUDA-event timed on HBM-resident
operands larger than L2, with a sampled fp64 accuracy check so a fast wrong kernel is caught here.
    python tests/bench_linear.py [--rows 4194304] [--shapes 602x64,64x41] [--iters 5]
Prints one JSON line per (shape, op): ms, achieved GB/s over the algorithmic bytes, max rel error.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roc_b200 import kernels as K  # noqa: E402


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 22)
    ap.add_argument("--shapes", default="602x64,64x41")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    for shp in a.shapes.split(","):
        i, o = (int(t) for t in shp.split("x"))
        n = a.rows
        x = K.padded(n, i, dev)
        x[:, :i] = torch.randn(n, i, device=dev, generator=g)
        w = torch.randn(o, i, device=dev, generator=g) * 0.1
        dy = K.padded(n, o, dev)
        dy[:, :o] = torch.randn(n, o, device=dev, generator=g)
        y = K.padded(n, o, dev)
        # ---- forward
        ms = timeit(lambda: K.linear_fwd(x[:, :i], w, out=y[:, :o]), a.iters)
        rows = torch.randint(0, n, (4096,), device=dev, generator=g)
        rows[0], rows[1] = 0, n - 1
        ref = x[rows, :i].double() @ w.double().t()
        got = y[rows, :o].double()
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        byt = n * (i + o) * 4
        print(json.dumps({"op": "linear_fwd", "shape": shp, "rows": n, "ms": round(ms, 4),
                          "GBps": round(byt / ms / 1e6, 1), "max_rel_err": err}), flush=True)
        # ---- forward with the indegree-norm epilogue, and the fused dX (dropout bwd + relu mask + norm)
        re = torch.cumsum(torch.randint(1, 30, (n,), device=dev, generator=g), 0)
        ms = timeit(lambda: K.linear_fwd(x[:, :i], w, norm_row_end=re, out=y[:, :o]), a.iters)
        print(json.dumps({"op": "linear_fwd+norm", "shape": shp, "rows": n, "ms": round(ms, 4),
                          "GBps": round(byt / ms / 1e6, 1)}), flush=True)
        if i <= 256:
            mask = K.dropout_mask(n, i, 0, 0.5, 5, 1, dev)
            dxb, dwb = K.padded(n, i, dev), torch.zeros(o, i, device=dev)
            gy = K.padded(n, o, dev, fill=dy[:, :o])
            t0 = timeit(lambda: K.linear_bwd_fused(x[:, :i], w, None, gy, dwb, dxb, mask=mask, rate=0.5,
                                                   relu_of=x[:, :i], norm_row_end=re), a.iters)
            t1 = timeit(lambda: K.linear_bwd(x[:, :i], w, None, gy, dwb), a.iters)
            print(json.dumps({"op": "linear_dx_fused(bwd - dW-only bwd)", "shape": shp, "rows": n,
                              "ms": round(t0 - t1, 4)}), flush=True)
            del mask, dxb, gy
        # ---- dW (dX not requested)
        dw = torch.zeros(o, i, device=dev)

        def run_dw():
            dw.zero_()
            K.linear_bwd(x[:, :i], w, None, dy[:, :o], dw)
        ms = timeit(run_dw, a.iters)
        m = min(n, 1 << 16)        # fp64 check on a prefix (exact function of the same kernel path)
        dw2 = torch.zeros(o, i, device=dev)
        K.linear_bwd(x[:m, :i], w, None, dy[:m, :o], dw2)
        ref = dy[:m, :o].double().t() @ x[:m, :i].double()
        err = ((dw2.double() - ref).abs().max() / ref.abs().max()).item()
        # whole-range sanity through linearity: dW over all rows with dY = const column of ones
        print(json.dumps({"op": "linear_dw", "shape": shp, "rows": n, "ms": round(ms, 4),
                          "GBps": round(byt / ms / 1e6, 1), "max_rel_err_prefix": err}), flush=True)
        del x, dy, y


if __name__ == "__main__":
    main()

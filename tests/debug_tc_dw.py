"""debug_tc_dw.py — probe the tcgen05 dW kernel with structured inputs (not a pytest file)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roc_b200 import _lib  # noqa: E402
from roc_b200 import kernels as K  # noqa: E402

lib = _lib.lib


def run(rows, i_dim, o_dim, xf, gf, tag):
    dev = "cuda"
    v = torch.arange(rows, device=dev, dtype=torch.float32)[:, None]
    ii = torch.arange(i_dim, device=dev, dtype=torch.float32)[None, :]
    oo = torch.arange(o_dim, device=dev, dtype=torch.float32)[None, :]
    x = K.padded(rows, i_dim, dev, fill=xf(v, ii).expand(rows, i_dim))
    g = K.padded(rows, o_dim, dev, fill=gf(v, oo).expand(rows, o_dim))
    w = torch.zeros((o_dim, i_dim), device=dev)
    dw = torch.zeros((o_dim, i_dim), device=dev)
    nbytes = lib.roc_linear_bwd_workspace_bytes(rows, i_dim, o_dim)
    ws = torch.full((max(nbytes // 4, 4),), float("nan"), device=dev)
    rc = lib.roc_linear_bwd(rows, i_dim, o_dim, x.data_ptr(), x.stride(0), w.data_ptr(), None, 0, g.data_ptr(),
                            g.stride(0), dw.data_ptr(), None, 0, 0, 0, ws.data_ptr(), nbytes, None)
    torch.cuda.synchronize()
    want = (g.contiguous().double().T @ x.contiguous().double()).float()
    err = (dw - want).abs().max().item()
    print("%-28s rows=%d in=%d out=%d rc=%d maxerr=%.3g  ws nan=%d zero=%d of %d" %
          (tag, rows, i_dim, o_dim, rc, err, int(torch.isnan(ws).sum()), int((ws == 0).sum()), ws.numel()))
    if err > 1e-3 * max(1.0, want.abs().max().item()):
        print("  want[0:3,0:8]", want[:3, :8].cpu().numpy())
        print("  got [0:3,0:8]", dw[:3, :8].cpu().numpy())


for (rows, i_dim, o_dim) in [(64, 128, 64), (64, 32, 32), (129, 16, 16), (1000, 602, 64)]:
    run(rows, i_dim, o_dim, lambda v, i: torch.ones_like(v + i), lambda v, o: torch.ones_like(v + o), "ones x ones")
    run(rows, i_dim, o_dim, lambda v, i: i + 0 * v, lambda v, o: torch.ones_like(v + o), "x=i, g=1")
    run(rows, i_dim, o_dim, lambda v, i: torch.ones_like(v + i), lambda v, o: o + 0 * v, "x=1, g=o")
    run(rows, i_dim, o_dim, lambda v, i: (v == 3).float() + 0 * i, lambda v, o: (v == 3).float() + 0 * o, "delta v=3")
    run(rows, i_dim, o_dim, lambda v, i: (v % 7) + 0 * i, lambda v, o: (v % 5) + 0 * o, "x=v%7, g=v%5")

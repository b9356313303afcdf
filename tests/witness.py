"""An independent derivation of the composed models: torch fp64 + autograd (no code shared with the product or
with oracle.GcnOracle's hand-written backward).  Used by the GPU suite (product vs witness) and by the CPU suite
(oracle vs witness)."""
import numpy as np
import torch


def torch_witness(kind, row_end, col, feats, labels, mask, layers, weights):
    """The composed model in torch fp64 with autograd: logits and dL/dW for L = sum over Train vertices of the
    cross entropy (the reference's un-averaged gradient, softmax_kernel.cu:19-33).  An independent derivation:
    it shares no code with the product or with oracle.GcnOracle's hand-written backward."""
    n = row_end.shape[0]
    deg = np.diff(np.concatenate([[0], row_end.astype(np.int64)]))
    dst = np.repeat(np.arange(n), deg)
    a = torch.zeros((n, n), dtype=torch.float64)
    a[torch.from_numpy(dst), torch.from_numpy(col.astype(np.int64))] = 1.0
    dinv = torch.from_numpy(1.0 / np.sqrt(deg.astype(np.float64))).unsqueeze(1)
    ws = [torch.from_numpy(w.astype(np.float64)).requires_grad_(True) for w in weights]
    t = torch.from_numpy(feats.astype(np.float64))
    L = len(layers) - 1
    wi = 0
    for i in range(1, L + 1):
        d = t                                                   # dropout rate 0
        if kind == "gcn":
            z = (a @ ((d @ ws[wi].T) * dinv)) * dinv; wi += 1
            if i != L:
                z = torch.relu(z)
            if len(layers) > 3:
                z = z + d @ ws[wi].T; wi += 1
            t = z
        else:                                                   # GraphSAGE-mean: D^-1 A (d W_nb) + d W_root
            nb = ((a @ (d @ ws[wi].T)) * dinv) * dinv; wi += 1
            z = nb + d @ ws[wi].T; wi += 1
            t = torch.relu(z) if i != L else z
    logits = t
    train = torch.from_numpy(mask == 0)
    loss = torch.nn.functional.cross_entropy(logits[train], torch.from_numpy(labels.astype(np.int64))[train], reduction="sum")
    loss.backward()
    return logits.detach().numpy(), [w.grad.numpy() for w in ws]



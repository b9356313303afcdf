"""oracle.py — ctypes front end of oracle/libroc_oracle.so plus the GCN training
epoch composed from its ops.

TEST INFRASTRUCTURE ONLY (see oracle/roc_oracle.c header): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under roc_b200/ imports it.

The model composition mirrors the reference driver: the GCN stack of
gnn.cc:78-92, Model::forward/backward/update (gnn.cc:696-724) including the
resetInputGrads first-writer rule (gnn.cc:704-713), AdamOptimizer::next
(optimizer.cc:79-85) and the lr decay of gnn.cc:100-101.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libroc_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.roc_oracle_num_threads.restype = C.c_int
        _LIB.roc_oracle_partition.restype = C.c_int
    return _LIB


def _p(a, t=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle wants contiguous arrays"
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().roc_oracle_num_threads()


def set_num_threads(n):
    lib().roc_oracle_set_num_threads(C.c_int(n))


# ---------------------------------------------------------------- graph ------
def partition(row_end, num_parts):
    """gnn.cc:806-829, 852-870. Returns (num_ranges, vbounds[P,2] u32, ebounds[P,2] u64)."""
    row_end = np.ascontiguousarray(row_end, dtype=np.uint64)
    n = row_end.shape[0]
    e = int(row_end[-1]) if n else 0
    cap = max(num_parts, 1) * 4 + 8
    vb = np.zeros((cap, 2), dtype=np.uint32)
    eb = np.zeros((cap, 2), dtype=np.uint64)
    k = lib().roc_oracle_partition(C.c_uint32(n), C.c_uint64(e), C.c_int(num_parts), _p(row_end), _p(vb), _p(eb),
                                   C.c_int(cap))
    return k, vb[: min(k, cap)].copy(), eb[: min(k, cap)].copy()


def build_csr(row_left, row_right, col_left, raw_rows, raw_cols):
    """init_graph_kernel, load_task.cu:271-294 -> (rowPtrs u64[Nloc], colIdxs u32[Eloc,2])."""
    raw_rows = np.ascontiguousarray(raw_rows, dtype=np.uint64)
    raw_cols = np.ascontiguousarray(raw_cols, dtype=np.uint32)
    nloc = row_right - row_left + 1
    eloc = int(raw_rows[-1]) - col_left if nloc else 0
    rp = np.zeros(nloc, dtype=np.uint64)
    ci = np.zeros((max(eloc, 1), 2), dtype=np.uint32)
    lib().roc_oracle_build_csr(C.c_uint32(row_left), C.c_uint32(row_right), C.c_uint64(col_left), _p(raw_rows),
                               _p(raw_cols), _p(rp), _p(ci))
    return rp, ci[:eloc]


def scatter_gather(row_left, row_right, col_left, row_end, col_src, x, acc64=True):
    """aggre_coop_kernel, scattergather_kernel.cu:20-76. x: [N][H] whole matrix -> [Nloc][H]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    h = x.shape[1]
    nloc = row_right - row_left + 1
    out = np.empty((nloc, h), dtype=np.float32)
    lib().roc_oracle_scatter_gather(C.c_uint32(row_left), C.c_uint32(row_right), C.c_uint64(col_left), C.c_int(h),
                                    _p(np.ascontiguousarray(row_end, dtype=np.uint64)),
                                    _p(np.ascontiguousarray(col_src, dtype=np.uint32)), _p(x), _p(out),
                                    C.c_int(1 if acc64 else 0))
    return out


def indegree_norm(row_left, row_right, col_left, row_end, x):
    """norm_coop_kernel, graphnorm_kernel.cu:19-57. x: the partition's [Nloc][H]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().roc_oracle_indegree_norm(C.c_uint32(row_left), C.c_uint32(row_right), C.c_uint64(col_left),
                                   C.c_int(x.shape[1]), _p(np.ascontiguousarray(row_end, dtype=np.uint64)), _p(x),
                                   _p(out))
    return out


def linear_fwd(x, w, relu=False, acc64=True):
    """linear_kernel.cu:76-104. x [N][in], w [out][in] (= W_mem[o*in+i]) -> [N][out]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    y = np.empty((x.shape[0], w.shape[0]), dtype=np.float32)
    lib().roc_oracle_linear_fwd(C.c_int64(x.shape[0]), C.c_int(x.shape[1]), C.c_int(w.shape[0]), _p(x), _p(w), _p(y),
                                C.c_int(int(relu)), C.c_int(int(acc64)))
    return y


def linear_bwd(x, w, y, dy, dw, need_dx=True, dx=None, relu=False, acc64=True):
    """linear_kernel.cu:129-245. dy is modified in place when relu; dw accumulated in place.
    Returns dx (new array, or accumulated into the given dx) or None."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    assert dy.flags["C_CONTIGUOUS"] and dw.flags["C_CONTIGUOUS"]
    accumulate = dx is not None
    if need_dx and dx is None:
        dx = np.zeros_like(x)
    lib().roc_oracle_linear_bwd(C.c_int64(x.shape[0]), C.c_int(x.shape[1]), C.c_int(w.shape[0]), _p(x), _p(w),
                                _p(y) if y is not None else None, _p(dy), _p(dw), _p(dx) if need_dx else None,
                                C.c_int(int(relu)), C.c_int(int(accumulate)), C.c_int(int(acc64)))
    return dx if need_dx else None


def activation_fwd(x, mode):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().roc_oracle_activation_fwd(C.c_int64(x.size), C.c_int(mode), _p(x), _p(y))
    return y


def activation_bwd(y, dy, mode, dx=None):
    acc = dx is not None
    if dx is None:
        dx = np.zeros_like(y)
    lib().roc_oracle_activation_bwd(C.c_int64(y.size), C.c_int(mode), _p(np.ascontiguousarray(y)),
                                    _p(np.ascontiguousarray(dy)), _p(dx), C.c_int(int(acc)))
    return dx


def dropout_mask(first_row, rows, h, rate, seed, step):
    """keep[rows][h] (uint8) of rows first_row.. of a width-h tensor (roc_oracle_dropout_mask)."""
    keep = np.empty((rows, h), dtype=np.uint8)
    lib().roc_oracle_dropout_mask(C.c_int64(first_row), C.c_int64(rows), C.c_int(h), C.c_float(rate),
                                  C.c_uint64(seed), C.c_uint32(step), _p(keep))
    return keep


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.empty(4, dtype=np.uint32)
    lib().roc_oracle_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def dropout_apply(x, keep, rate):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().roc_oracle_dropout_apply(C.c_int64(x.size), C.c_float(rate), _p(np.ascontiguousarray(keep)), _p(x), _p(y))
    return y


class Perf(C.Structure):
    _fields_ = [("trainLoss", C.c_float), ("trainAll", C.c_int), ("testAll", C.c_int), ("valAll", C.c_int),
                ("trainCorrect", C.c_int), ("testCorrect", C.c_int), ("valCorrect", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def softmax_xent_bwd(logits, onehot, mask):
    """softmax_kernel.cu:81-171 -> (grad [N][C], perf dict)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    onehot = np.ascontiguousarray(onehot, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.int32)
    g = np.empty_like(logits)
    perf = Perf()
    lib().roc_oracle_softmax_xent_bwd(C.c_int64(logits.shape[0]), C.c_int(logits.shape[1]), _p(logits), _p(onehot),
                                      _p(mask), _p(g), C.byref(perf))
    return g, perf.as_dict()


def adam_update(w, g, m, v, alpha_t, beta1, beta2, wd, eps):
    lib().roc_oracle_adam_update(C.c_int64(w.size), C.c_float(alpha_t), C.c_float(beta1), C.c_float(beta2),
                                 C.c_float(wd), C.c_float(eps), _p(g), _p(m), _p(v), _p(w))


def glorot_scale(u, in_dim, out_dim):
    w = np.ascontiguousarray(u, dtype=np.float32).copy()
    lib().roc_oracle_glorot_scale(C.c_int64(w.size), C.c_int(in_dim), C.c_int(out_dim), _p(w))
    return w


def lux_read(path, row_left=None, row_right=None):
    n = C.c_uint32()
    e = C.c_uint64()
    rc = lib().roc_oracle_lux_header(path.encode(), C.byref(n), C.byref(e))
    if rc != 0:
        raise IOError("cannot read %s (%d)" % (path, rc))
    n, e = n.value, e.value
    rows = np.empty(n, dtype=np.uint64)
    cols = np.empty(max(e, 1), dtype=np.uint32)
    rc = lib().roc_oracle_lux_read(path.encode(), C.c_uint32(0), C.c_uint32(n - 1), C.c_uint64(0), C.c_uint64(e - 1),
                                   _p(rows), _p(cols))
    if rc != 0:
        raise IOError("cannot read %s (%d)" % (path, rc))
    return n, e, rows, cols[:e]


# ------------------------------------------------------------- the model -----
class GcnOracle:
    """The reference GCN of gnn.cc:78-92 on one partition covering the whole graph
    (numParts = 1), or on `parts` partitions executed one after another the way
    Legion would run the index launch (dW replicas summed in order,
    optimizer_kernel.cu:88-94).  Weights are given (Glorot draws come from cuRAND
    in the reference); dropout uses the documented Philox mask or rate 0."""

    def __init__(self, row_end, col_src, layers, weights, lr=0.01, weight_decay=0.05, dropout=0.0,
                 beta1=0.9, beta2=0.999, eps=1e-8, acc64=True, dropout_seed=0, parts=1):
        self.row_end = np.ascontiguousarray(row_end, dtype=np.uint64)
        self.col_src = np.ascontiguousarray(col_src, dtype=np.uint32)
        self.N = self.row_end.shape[0]
        self.layers = list(layers)
        self.W = [np.ascontiguousarray(w, dtype=np.float32).copy() for w in weights]
        self.M = [np.zeros_like(w) for w in self.W]
        self.V = [np.zeros_like(w) for w in self.W]
        self.alpha, self.beta1, self.beta2, self.eps, self.wd = float(lr), beta1, beta2, eps, float(weight_decay)
        self.beta1_t, self.beta2_t, self.alpha_t = 1.0, 1.0, float(lr)
        self.dropout, self.acc64, self.dropout_seed = float(dropout), acc64, dropout_seed
        self.step = 0
        self.residual = len(self.layers) > 3          # gnn.cc:86
        self.parts = parts
        k, vb, eb = partition(self.row_end, parts)
        assert k == parts, "partitioner produced %d ranges for %d parts" % (k, parts)
        self.vb, self.eb = vb, eb

    # op indices of the dropout layers, matching Model::dropout's position in `layers`
    def _dropout_op_index(self, layer_i):
        # per GCN layer the script appends: dropout, linear, norm, sg, norm, [relu], [linear, add]
        idx = 0
        L = len(self.layers) - 1
        for i in range(1, layer_i):
            idx += 5 + (1 if i != L else 0) + (2 if self.residual else 0)
        return idx

    def _sg(self, x):
        outs = []
        for c in range(self.parts):
            rl, rr = int(self.vb[c, 0]), int(self.vb[c, 1])
            cl = int(self.eb[c, 0])
            outs.append(scatter_gather(rl, rr, cl, self.row_end[rl:rr + 1], self.col_src[cl:int(self.eb[c, 1]) + 1],
                                       x, self.acc64))
        return np.concatenate(outs, axis=0)

    def _norm(self, x):
        return indegree_norm(0, self.N - 1, 0, self.row_end, x)

    def forward(self, feats, train=True):
        if train:
            self.step += 1
        L = len(self.layers) - 1
        t = np.ascontiguousarray(feats, dtype=np.float32)
        self.saved = []
        wi = 0
        for i in range(1, L + 1):
            rate = self.dropout if train else 0.0
            if rate > 0.0:
                key = (self.dropout_seed << 32) | self._dropout_op_index(i)
                keep = dropout_mask(0, t.shape[0], t.shape[1], rate, key, self.step)
                d = dropout_apply(t, keep, rate)
            else:
                keep = None
                d = t.copy()
            w_main = wi
            lin = linear_fwd(d, self.W[wi], False, self.acc64); wi += 1
            n1 = self._norm(lin)
            sg = self._sg(n1)
            n2 = self._norm(sg)
            a = activation_fwd(n2, 1) if i != L else n2
            rec = dict(d=d, keep=keep, rate=rate, w_main=w_main, a=a, relu=(i != L), pre=n2)
            if self.residual:
                rec["w_res"] = wi
                r = linear_fwd(d, self.W[wi], False, self.acc64); wi += 1
                t = a + r
            else:
                t = a
            self.saved.append(rec)
        self.logits = t
        return t

    def backward(self, onehot, mask):
        grad, perf = softmax_xent_bwd(self.logits, onehot, mask)
        self.perf = perf
        dW = [np.zeros_like(w) for w in self.W]
        g = grad
        for li in range(len(self.saved) - 1, -1, -1):
            rec = self.saved[li]
            first_layer = li == 0
            d_d = None   # gradient wrt the dropout output (accumulated from both linears)
            if self.residual:
                # add backward: both inputs take dOut (element_kernel.cu:93-101); residual
                # linear was built later, so its backward runs first (reverse layer order)
                d_d = linear_bwd(rec["d"], self.W[rec["w_res"]], None, g.copy(), dW[rec["w_res"]],
                                 need_dx=not first_layer, acc64=self.acc64)
            ga = g
            if rec["relu"]:
                # "a_override": relu output whose sign pattern the backward should use instead of the
                # oracle's own (a test may inject the product's mask where pre-activations are ~0)
                ga = activation_bwd(rec.get("a_override", rec["a"]), ga, 1)
            gn2 = self._norm(ga)
            gsg = self._sg(gn2)          # A, not A^T (scattergather_kernel.cu:160-170)
            gn1 = self._norm(gsg)
            d_d = linear_bwd(rec["d"], self.W[rec["w_main"]], None, gn1, dW[rec["w_main"]],
                             need_dx=not first_layer, dx=d_d, acc64=self.acc64)
            if not first_layer:
                if rec["rate"] > 0.0:
                    g = dropout_apply(d_d, rec["keep"], rec["rate"])
                else:
                    g = d_d
        self.dW = dW
        return dW

    def update(self):
        self.beta1_t *= self.beta1
        self.beta2_t *= self.beta2
        self.alpha_t = self.alpha * np.sqrt(1 - self.beta2_t) / (1 - self.beta1_t)
        for p in range(len(self.W) - 1, -1, -1):   # gnn.cc:721-723 reverse order
            adam_update(self.W[p], self.dW[p], self.M[p], self.V[p], np.float32(self.alpha_t), np.float32(self.beta1),
                        np.float32(self.beta2), np.float32(self.wd), np.float32(self.eps))

    def train_epoch(self, feats, onehot, mask):
        self.forward(feats, train=True)
        self.backward(onehot, mask)
        self.update()
        return self.perf

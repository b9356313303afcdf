#!/usr/bin/env python
"""bench.py — GCN training edges/sec on synthetic R-MAT graphs (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] at N = 1 — 2-layer GCN
602 -> 64 -> 41 on an R-MAT scale-22 graph (2^22 vertices, 2^25 undirected pairs
before symmetrisation + dedup + self loops, ~64M edges), dropout 0.5, Adam.
For N > 1 the graph grows with N (scale 22 + log2 N, pairs 2^25 * N: weak
scaling, the N = 8 point is configs[3]'s R-MAT scale-25 graph) and is
vertex-range partitioned by the reference's own partitioner.

A "step" is one training epoch: zero_gradients + forward + backward + update
(gnn.cc:103-106), everything resident in HBM.  `value` = total edges * steps /
max-over-ranks device time (CUDA events on the engine's stream).  `e2e` = the same
metric through the public API with HOST buffers: every step re-uploads the
features / labels / mask from pinned host memory (H2D inside the timed region)
and reads the loss metrics back (D2H).  `roofline` is the ScatterGather launch
at H = 64 timed live with CUDA events inside the timed steps.  `cpu_baseline` /
`--impl reference` time the CPU oracle (OpenMP, all host cores) on a bounded
sample of the same workload — ROC ships no CPU kernels.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS = [602, 64, 41]
DROPOUT = 0.5
LR, WD = 0.01, 0.0001          # example_run.sh: -lr 0.01 -decay 0.0001
BASE_SCALE = 22
BASE_PAIRS = 1 << 25
CPU_SAMPLE_SCALE = 19          # cpu_baseline / reference arm: same generator, 1/8 of the vertices


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_graph(scale, pairs, device):
    import torch
    from roc_b200 import datasets
    t0 = time.time()
    row_end, col = datasets.rmat_graph(scale, pairs, seed=1, device=device)
    if device != "cpu":
        torch.cuda.synchronize()
    re_h = row_end.cpu().numpy().astype(np.uint64)
    col_h = col.cpu().numpy().astype(np.uint32)
    del row_end, col
    if device != "cpu":
        torch.cuda.empty_cache()
    log("[bench] R-MAT scale %d: N=%d E=%d (%.1fs)" % (scale, re_h.shape[0], col_h.shape[0], time.time() - t0))
    return re_h, col_h


def sg_bytes(n, e, h):
    """Algorithmic bytes of one ScatterGather launch: E*(4H+4) + N*(4H+8)."""
    return e * (4 * h + 4) + n * (4 * h + 8)


# --------------------------------------------------------------------- CPU arm ---
def cpu_epoch_rate(steps, warmup, device_for_gen):
    """Oracle GCN epoch (fp32, OpenMP) on the bounded sample; returns (edges/s, ms/step, info)."""
    from oracle import oracle
    from roc_b200 import datasets
    pairs = BASE_PAIRS >> (BASE_SCALE - CPU_SAMPLE_SCALE)
    re_h, col_h = make_graph(CPU_SAMPLE_SCALE, pairs, device_for_gen)
    n, e = re_h.shape[0], col_h.shape[0]
    feats, labels, mask = datasets.node_data(n, LAYERS[0], LAYERS[-1], seed=1)
    rng = np.random.RandomState(1)
    ws = [((rng.rand(LAYERS[i + 1], LAYERS[i]).astype(np.float32) * 2 - 1) *
           np.float32(math.sqrt(6.0 / (LAYERS[i] + LAYERS[i + 1])))) for i in range(len(LAYERS) - 1)]
    m = oracle.GcnOracle(re_h, col_h, LAYERS, ws, lr=LR, weight_decay=WD, dropout=DROPOUT, acc64=False)
    oh = datasets.onehot(labels.numpy(), LAYERS[-1])
    f, mk = feats.numpy(), mask.numpy()
    for _ in range(warmup):
        m.train_epoch(f, oh, mk)
    t0 = time.time()
    for _ in range(steps):
        m.train_epoch(f, oh, mk)
    dt = time.time() - t0
    info = {"cores": oracle.num_threads(), "kind": "port",
            "sample": "oracle GCN %s epoch (fp32, OpenMP) on R-MAT scale-%d (N=%d, E=%d), %d steps" %
                      ("-".join(map(str, LAYERS)), CPU_SAMPLE_SCALE, n, e, steps)}
    return e * steps / dt, 1e3 * dt / steps, info


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    rate, ms, info = cpu_epoch_rate(steps, warm, "cpu")
    info["value"] = rate
    line = {"impl": "reference", "metric": "gcn_training_edges_per_sec", "value": rate, "unit": "edges/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "2-layer GCN 602-64-41, R-MAT (bounded sample: scale %d)" % CPU_SAMPLE_SCALE,
                       "note": "ROC ships CUDA-only kernels and cannot be built here (Legion absent); the reference "
                               "arm is the CPU restatement of its algorithm (oracle/) on all host cores"},
            "cpu_baseline": info,
            "e2e": {"value": rate, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------- GPU arm ---
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from roc_b200 import _lib, datasets
    from roc_b200.model import Host, Model, build_gcn

    _lib.require_device()
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    scale = BASE_SCALE + int(round(math.log2(world)))
    pairs = BASE_PAIRS * world
    if args.scale:
        scale, pairs = args.scale, (1 << (args.scale + 3)) * 1
    re_h, col_h = make_graph(scale, pairs, dev)
    n, e = re_h.shape[0], col_h.shape[0]

    host = Host(local_rank, rank, world)
    if world > 1:
        uid = [Host.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        host.nccl_init(uid[0])
    host.graph_from_arrays(re_h, col_h)
    info = host.graph_info()
    rl, rr = info["rowLeft"], info["rowRight"]
    nloc = rr - rl + 1
    m = Model(host, seed=1)
    hnd = build_gcn(m, LAYERS, DROPOUT, lr=LR, weight_decay=WD)

    # synthetic node data: pinned host copies (e2e uploads them every step) of this rank's rows
    g = torch.Generator(device="cpu"); g.manual_seed(1000 + rank)
    feats = torch.empty((nloc, LAYERS[0]), dtype=torch.float32, pin_memory=True)
    chunk = 1 << 18
    dg = torch.Generator(device=dev); dg.manual_seed(1000 + rank)
    for a in range(0, nloc, chunk):
        b = min(nloc, a + chunk)
        feats[a:b].copy_(torch.rand((b - a, LAYERS[0]), device=dev, generator=dg) * 2 - 1)
    labels = torch.randint(0, LAYERS[-1], (nloc,), generator=g, dtype=torch.int32).pin_memory()
    u = torch.rand(nloc, generator=g)
    mask = torch.full((nloc,), 2, dtype=torch.int32)
    mask[u < 0.76] = 1
    mask[u < 0.66] = 0
    mask = mask.pin_memory()
    torch.cuda.synchronize()

    def upload():
        m.set_tensor_from_host_ptr(hnd["input"], feats.data_ptr())
        m.set_labels(hnd["label"], labels.numpy())
        m.set_tensor_from_host_ptr(hnd["mask"], mask.data_ptr())
    upload()
    h2d = feats.numel() * 4 + labels.numel() * 4 + mask.numel() * 4
    d2h = 28   # sizeof(PerfMetrics)

    stream = torch.cuda.ExternalStream(host.stream, device=dev)

    def barrier():
        host.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    # ---- resident-in-HBM arm
    for _ in range(max(args.warmup, 3)):
        m.train_epoch()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m.profile_sg(True)
    l0 = _lib.lib.roc_launch_count()
    ms_total = timed(m.train_epoch, args.steps)
    launches = _lib.lib.roc_launch_count() - l0
    sg_times = m.profile_sg_read()
    m.profile_sg(False)
    clocks = sampler.stop() if rank == 0 else None
    perf = m.metrics()
    value = e * args.steps / (ms_total * 1e-3)

    # ---- end-to-end arm: host buffers in, metrics out, every step
    def e2e_step():
        upload()
        m.train_epoch()
        m.metrics()
    e2e_steps = max(1, min(args.steps, 3))
    if args.no_e2e:
        ms_e2e, e2e_value = float("nan"), None
    else:
        e2e_step()
        ms_e2e = timed(e2e_step, e2e_steps)
        e2e_value = e * e2e_steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: ScatterGather at H = 64 on this rank's partition
    eloc = info["colRight"] - info["colLeft"] + 1
    sg64 = [t for (h, t) in sg_times if h == LAYERS[1]]
    peak, peak_src = peaks()
    roof = None
    if sg64:
        t_avg = sum(sg64) / len(sg64)
        ach = sg_bytes(nloc, eloc, LAYERS[1]) / (t_avg * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "sg_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "sg_chunk_kernel<float4,L=16> (+fix-up) H=%d" % LAYERS[1], "launch_ms": t_avg,
                "launches_timed": len(sg64), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sg_bytes(nloc, eloc, LAYERS[1]),
                # the algorithmic bytes are mostly neighbour-row gathers and 59 % of those sectors hit
                # in L2 (hub rows), so `frac` can exceed 1; this one uses the ncu DRAM traffic instead
                "frac_dram": (traffic / (t_avg * 1e-3) / 1e9 / peak) if (traffic and world == 1) else None}
    sg_share = sum(t for _, t in sg_times) / ms_total if sg_times else None

    # ---- side columns (rank 0, N = 1 only): CPU oracle + the reference's own kernel on this GPU
    cpu = None
    refk = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            rate, _, cinfo = cpu_epoch_rate(1, 0, dev)
            cinfo["value"] = rate
            cinfo["unit"] = "edges/s"
            cpu = cinfo
        except Exception as ex:   # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"error": repr(ex)}
        try:
            refk = reference_kernel_column(re_h, col_h, dev)
        except Exception as ex:
            refk = {"error": repr(ex)}

    if rank == 0:
        line = {"metric": "gcn_training_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "2-layer GCN %s, R-MAT scale-%d (N=%d, E=%d incl. self loops), dropout %.1f, Adam"
                                       % ("-".join(map(str, LAYERS)), scale, n, e, DROPOUT),
                           "parallelism": "vertex-range dp%d" % world,
                           "l2": "inputs larger than L2 (features %.1f GB/GPU, graph %.2f GB)" %
                                 (nloc * 604 * 4 / 1e9, (e * 4 + n * 8) / 1e9)},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
                "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu,
                "sg_share_of_step": sg_share, "reference_kernel": refk,
                "train_loss": perf["trainLoss"], "plan": host.plan_info()}
        print(json.dumps(line), flush=True)
    host.close()
    if world > 1:
        dist.destroy_process_group()


def reference_kernel_column(row_end_h, col_h, dev, h=64, iters=5):
    """The reference's own aggre_coop_kernel (cut from scattergather_kernel.cu:20-76 into oracle/_ref) on
    the same graph and GPU, H = 64, HBM-resident buffers, the reference's grid.  It omits the reference's
    per-call PCIe staging (types.cu:28, scattergather_kernel.cu:145-157): an upper bound for ROC."""
    import torch
    from oracle import ref
    if not ref.available():
        return {"unavailable": "oracle/_ref/libroc_ref.so not built"}
    n, e = row_end_h.shape[0], col_h.shape[0]
    d_re = torch.from_numpy(row_end_h.astype(np.int64)).to(dev)
    d_col = torch.from_numpy(col_h.astype(np.int32)).to(dev)
    rp, es = ref.edge_structs(d_col, d_re, 0, 0)
    x = torch.rand((n, h), device=dev) - 0.5
    out = torch.empty((n, h), device=dev)
    for _ in range(2):
        ref.scatter_gather(0, n - 1, 0, rp, es, x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ref.scatter_gather(0, n - 1, 0, rp, es, x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"kernel": "aggre_coop_kernel (reference, sm_100a build)", "H": h, "launch_ms": ms,
            "algorithmic_GBps": sg_bytes(n, e, h) / (ms * 1e-3) / 1e9, "edges_per_s": e / (ms * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=0, help="override the R-MAT scale (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        log("[bench] WORLD_SIZE=%d but --gpus %d: launch with torch.distributed.run --nproc-per-node %d"
            % (world, args.gpus, args.gpus))
        if world == 1 and args.gpus > 1:
            sys.exit(2)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
